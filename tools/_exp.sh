python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
python tools/_exp_steps.py 2>&1 | grep -A13 "timing mean"
