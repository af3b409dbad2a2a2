python -m pytest tests/test_gpu_parity.py -x -q -k "device_demand or baseline or config1" 2>&1 | tail -3
python tools/_exp_k4.py 2>&1 | tail -5
