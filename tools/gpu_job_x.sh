mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu28.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu28.log
grep -E "passed|failed|^FAILED|^E   .*(assert|Error)" gpurun_out/pytest_gpu28.log | cut -c1-300 | head -30
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke4.log 2>&1; tail -1 gpurun_out/smoke4.log
timeout -k 10 900 python bench.py > gpurun_out/bench25_default.log 2>&1; tail -1 gpurun_out/bench25_default.log | cut -c1-2800
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench25_n2.log 2>&1; echo "n2 exit $?"; tail -1 gpurun_out/bench25_n2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=2', d['ms_per_step'], d['value'], d['e2e']['value'], d['n_gpus'])"
