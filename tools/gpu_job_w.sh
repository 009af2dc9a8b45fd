mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_tcconv.py tests/test_gpu_engine.py tests/test_gpu_synthesis.py -m gpu -q --timeout 300 > gpurun_out/pytest_gpu27.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu27.log
grep -E "passed|failed|^FAILED|^E   .*(assert|Error)" gpurun_out/pytest_gpu27.log | cut -c1-300 | head -30
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench24.log 2>&1; tail -1 gpurun_out/bench24.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline_tensor']['achieved'])"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r1_engine14.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_ncu15.log 2>&1; echo "ncu-list exit $?"
