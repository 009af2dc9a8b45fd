mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.log 2>&1; echo "n2 exit $?"; tail -1 gpurun_out/bench_n2.log | cut -c1-700
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_n2_ref.log 2>&1; echo "n2 ref exit $?"; tail -1 gpurun_out/bench_n2_ref.log | cut -c1-400
timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1_same_box.log 2>&1; tail -1 gpurun_out/bench_n1_same_box.log | cut -c1-200
