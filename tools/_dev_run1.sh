mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2_gpu_tests_b.log
python tools/bench_sphere_model.py 2>&1 | tail -2 | tee gpurun_out/r2_ext_v4_full.jsonl
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > gpurun_out/r02_san_memcheck.log 2>&1; tail -2 gpurun_out/r02_san_memcheck.log
