mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu13.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu13.log
grep -E "passed|failed|^FAILED|^E   .*(assert|Error)" gpurun_out/pytest_gpu13.log | cut -c1-200 | head -30
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench11.log 2>&1; tail -1 gpurun_out/bench11.log | cut -c1-200; tail -1 gpurun_out/bench11.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d.get('roofline_tensor'))"
