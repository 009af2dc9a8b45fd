mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_gpu9.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu9.log
grep -E "passed|failed|FAILED|^E  " gpurun_out/pytest_gpu9.log | head -20
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench7.log 2>&1; tail -1 gpurun_out/bench7.log | cut -c1-250
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r1_engine5.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_ncu6.log 2>&1; echo "ncu-list exit $?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:fir_act_nhwc -s 36 -c 4 -f -o gpurun_out/fir_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_fir.log 2>&1; echo "ncu-fir exit $?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:modulate_weights -s 96 -c 3 -f -o gpurun_out/modw_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_modw.log 2>&1; echo "ncu-modw exit $?"
