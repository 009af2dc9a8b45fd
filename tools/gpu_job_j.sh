mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu15.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu15.log
grep -E "passed|failed|^FAILED|^E   .*(assert|Error)" gpurun_out/pytest_gpu15.log | cut -c1-200 | head -30
timeout -k 10 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench13.log 2>&1; tail -1 gpurun_out/bench13.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:render_fwd_tc -s 2 -c 1 -f -o gpurun_out/render_tc_full5 python tools/profile_render.py 1 tc > gpurun_out/ncu_render_tc5.log 2>&1; echo "ncu-render exit $?"
