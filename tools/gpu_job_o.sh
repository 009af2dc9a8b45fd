mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu20.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu20.log
grep -E "passed|failed|^FAILED|^E   .*(assert|Error)" gpurun_out/pytest_gpu20.log | cut -c1-300 | head -30
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench17.log 2>&1; tail -1 gpurun_out/bench17.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'])"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r1_engine11.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_ncu12.log 2>&1; echo "ncu-list exit $?"
