mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_synthesis.py -m gpu -q --timeout 300 > gpurun_out/pytest_gpu16.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu16.log
grep -E "passed|failed|^FAILED|^E   .*(assert|Error)" gpurun_out/pytest_gpu16.log | cut -c1-200 | head -30
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench14.log 2>&1; tail -1 gpurun_out/bench14.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernel_ms'])"
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-graph > gpurun_out/bench14_eager.log 2>&1; tail -1 gpurun_out/bench14_eager.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('eager', d['ms_per_step'], d['value'])"
