mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_ik_adam_hot -s 1 -c 1 -f -o gpurun_out/r2_ik_hot python tools/bench_ik.py > gpurun_out/r2_ik_ncu.log 2>&1
tail -2 gpurun_out/r2_ik_ncu.log
