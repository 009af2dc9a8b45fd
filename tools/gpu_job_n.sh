mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 --durations=8 > gpurun_out/pytest_gpu19.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu19.log
grep -E "passed|failed|^FAILED|^E   .*(assert|Error)|s call" gpurun_out/pytest_gpu19.log | cut -c1-300 | head -30
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench16.log 2>&1; tail -1 gpurun_out/bench16.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['roofline_tensor']['achieved'])"
