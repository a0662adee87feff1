mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_objective_hot -s 28 -c 1 -f -o gpurun_out/r2_hot_tableshelf python tools/bench_envs.py > /dev/null 2>&1
ls -la gpurun_out/r2_hot_tableshelf.ncu-rep
