mkdir -p gpurun_out
python -m pytest tests/test_gpu_dropin.py tests/test_reference_examples.py -m gpu -q -x -s -k "inverse or ik" 2>&1 | grep -v "^$" | tail -12
