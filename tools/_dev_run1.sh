mkdir -p gpurun_out
python tools/bench_sphere_model.py 2>&1 | tail -6 | tee gpurun_out/r2_ext_v5.jsonl
TRK_B200_NO_HOT=1 python tools/bench_sphere_model.py 2>&1 | tail -6 | tee gpurun_out/r2_ext_v5_nohot.jsonl
