mkdir -p gpurun_out
python -m pytest tests/test_specialize.py -m gpu -q -x 2>&1 | tail -15
