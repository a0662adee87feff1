mkdir -p gpurun_out
for i in 1 2; do for v in nobar bar; do
  TRK_B200_LIB=build_variants/libtrk_$v.so python tools/time_objective.py 2>&1 | tail -1 | tee -a gpurun_out/r2_variants_c.txt
done; done
TRK_B200_LIB=build_variants/libtrk_nobar.so python -m pytest tests/test_gpu_baseline_configs.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
