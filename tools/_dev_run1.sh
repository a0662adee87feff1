python -m pytest tests/test_base_translation.py tests/test_gpu_dropin.py tests/test_environments.py tests/test_attached_points.py -m gpu -q 2>&1 | tail -8
