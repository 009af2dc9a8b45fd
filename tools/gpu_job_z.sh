mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_engine.py tests/test_gpu_synthesis.py -m gpu -q --timeout 300 -k "encoder or mapping" > gpurun_out/pytest_gpu30.log 2>&1; echo "encoder tests exit $?"
grep -E "passed|failed|^FAILED|^E   .*(assert|Error)" gpurun_out/pytest_gpu30.log | cut -c1-300 | head -10
timeout 600 python tools/time_mapping.py > gpurun_out/time_mapping3.log 2>&1; head -1 gpurun_out/time_mapping3.log | cut -c1-200; sed -n 6,20p gpurun_out/time_mapping3.log | cut -c1-70,150-230
