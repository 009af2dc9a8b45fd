mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_renderer.py -m gpu -q --timeout 300 > gpurun_out/pytest_gpu26.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu26.log
grep -E "passed|failed|^FAILED|^E   .*(assert|Error)" gpurun_out/pytest_gpu26.log | cut -c1-300 | head -30
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench23.log 2>&1; tail -1 gpurun_out/bench23.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
