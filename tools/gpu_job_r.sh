mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_tcconv.py -m gpu -q --timeout 120 -k "cta_pair" > gpurun_out/pytest_gpu22a.log 2>&1; echo "pair test exit $?"
grep -E "passed|failed|^FAILED|^E   .*(assert|Error)" gpurun_out/pytest_gpu22a.log | cut -c1-300 | head -10
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu22.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu22.log
grep -E "passed|failed|^FAILED|^E   .*(assert|Error)" gpurun_out/pytest_gpu22.log | cut -c1-300 | head -30
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench19.log 2>&1; tail -1 gpurun_out/bench19.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline_tensor']['achieved'])"
P3D_CONV_PAIR=0 timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench19_nopair.log 2>&1; tail -1 gpurun_out/bench19_nopair.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('nopair', d['ms_per_step'], d['value'])"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r1_engine12.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_ncu13.log 2>&1; echo "ncu-list exit $?"
