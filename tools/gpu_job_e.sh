mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu11.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu11.log
grep -E "passed|failed|^FAILED|^E   .*(assert|Error)" gpurun_out/pytest_gpu11.log | cut -c1-200 | head -30
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench9.log 2>&1; tail -1 gpurun_out/bench9.log | cut -c1-250
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r1_engine7.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_ncu8.log 2>&1; echo "ncu-list exit $?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:fir_act_nhwc -s 38 -c 2 -f -o gpurun_out/fir_full2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_fir2.log 2>&1; echo "ncu-fir exit $?"
