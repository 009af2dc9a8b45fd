mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu7.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/pytest_gpu7.log | head -20
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench5.log 2>&1; tail -2 gpurun_out/bench5.log
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-graph > gpurun_out/bench5_eager.log 2>&1; tail -1 gpurun_out/bench5_eager.log | cut -c1-400
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r1_engine3.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_ncu4.log 2>&1; echo "ncu-list exit $?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 180 -c 6 -f -o gpurun_out/conv_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_conv.log 2>&1; echo "ncu-conv exit $?"
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:render_fwd_tc -s 2 -c 1 -f -o gpurun_out/render_tc_full2 python tools/profile_render.py 1 tc > gpurun_out/ncu_render_tc2.log 2>&1; echo "ncu-render exit $?"
