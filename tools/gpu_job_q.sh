mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu21.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu21.log
grep -E "passed|failed|^FAILED|^E   .*(assert|Error)" gpurun_out/pytest_gpu21.log | cut -c1-300 | head -30
timeout -k 10 900 python bench.py > gpurun_out/bench18_default.log 2>&1; tail -1 gpurun_out/bench18_default.log | cut -c1-3000
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke3.log 2>&1; tail -2 gpurun_out/smoke3.log
