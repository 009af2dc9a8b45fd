mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests/test_gpu_engine.py tests/test_gpu_synthesis.py -m gpu -q --timeout 600 -k "full_size or training_step" > gpurun_out/pytest_gpu17.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu17.log
grep -E "passed|failed|^FAILED|^E   .*(assert|Error)" gpurun_out/pytest_gpu17.log | cut -c1-300 | head -30
