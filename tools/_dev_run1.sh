# scratch script of the development loop: `gpurun -- bash tools/_dev_run1.sh`
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4
