mkdir -p gpurun_out
python tools/time_objective.py 2>&1 | tail -1
python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q 2>&1 | tail -2
