mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_tcconv.py -m gpu -q --timeout 200 -k "strided" > gpurun_out/pytest_gpu29a.log 2>&1; rc=$?; echo "strided test exit $rc"
grep -E "passed|failed|^FAILED|^E   .*(assert|Error)" gpurun_out/pytest_gpu29a.log | cut -c1-300 | head -10
if [ $rc -ne 0 ]; then tail -30 gpurun_out/pytest_gpu29a.log | cut -c1-200; exit 0; fi
timeout -k 10 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_synthesis.py -m gpu -q --timeout 300 -k "encoder or mapping" > gpurun_out/pytest_gpu29b.log 2>&1; echo "encoder tests exit $?"
grep -E "passed|failed|^FAILED|^E   .*(assert|Error)" gpurun_out/pytest_gpu29b.log | cut -c1-300 | head -10
timeout 600 python tools/time_mapping.py > gpurun_out/time_mapping2.log 2>&1; head -3 gpurun_out/time_mapping2.log | cut -c1-200
