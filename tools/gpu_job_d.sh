mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu10.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu10.log
grep -E "passed|failed|^FAILED|^E   .*(assert|Error)" gpurun_out/pytest_gpu10.log | cut -c1-200 | head -30
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench8.log 2>&1; tail -1 gpurun_out/bench8.log | cut -c1-250
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r1_engine6.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_ncu7.log 2>&1; echo "ncu-list exit $?"
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:render_fwd_tc -s 2 -c 1 -f -o gpurun_out/render_tc_full4 python tools/profile_render.py 1 tc > gpurun_out/ncu_render_tc4.log 2>&1; echo "ncu-render exit $?"
