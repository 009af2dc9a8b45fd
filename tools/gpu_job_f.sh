mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu12.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu12.log
grep -E "passed|failed|^FAILED|^E   .*(assert|Error)" gpurun_out/pytest_gpu12.log | cut -c1-200 | head -30
timeout -k 10 900 python bench.py > gpurun_out/bench10_default.log 2>&1; tail -1 gpurun_out/bench10_default.log | cut -c1-1800
timeout -k 10 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench10_ref.log 2>&1; tail -1 gpurun_out/bench10_ref.log | cut -c1-600
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r1_engine8.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_ncu9.log 2>&1; echo "ncu-list exit $?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke2.log 2>&1; tail -2 gpurun_out/smoke2.log
