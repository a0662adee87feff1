mkdir -p gpurun_out
python tools/sanitize_smoke.py 2>&1 | tail -2
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > gpurun_out/r02_san_memcheck.log 2>&1; tail -3 gpurun_out/r02_san_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > gpurun_out/r02_san_racecheck.log 2>&1; tail -3 gpurun_out/r02_san_racecheck.log
