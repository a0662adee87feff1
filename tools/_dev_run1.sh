mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r2_gpu_tests_c.log
python tools/time_objective.py 2>&1 | tail -1 | tee gpurun_out/r2_time_hot_c.json
python tools/bench_envs.py 2>&1 | tail -6 | tee gpurun_out/r2_bench_envs.jsonl
