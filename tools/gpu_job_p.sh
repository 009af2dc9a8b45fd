mkdir -p gpurun_out
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:fir_act_nhwc_sep -s 5 -c 2 -f -o gpurun_out/fir_sep_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_fir3.log 2>&1; echo "ncu-fir exit $?"
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_persist -s 70 -c 3 -f -o gpurun_out/conv_persist_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_conv3.log 2>&1; echo "ncu-conv exit $?"
