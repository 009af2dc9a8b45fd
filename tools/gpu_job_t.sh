mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu24.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu24.log
grep -E "passed|failed|^FAILED|^E   .*(assert|Error)" gpurun_out/pytest_gpu24.log | cut -c1-300 | head -30
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench21.log 2>&1; tail -1 gpurun_out/bench21.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline_tensor']['achieved'], d['roofline']['frac'])"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_pair -s 12 -c 2 -f -o gpurun_out/conv_pair_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_conv4.log 2>&1; echo "ncu-pair exit $?"
