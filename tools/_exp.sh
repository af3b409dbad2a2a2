python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5
python tools/_exp_steps.py 2>&1 | grep -A13 "timing mean"
ncu --set full --import-source on --clock-control none -k regex:env_alloc --launch-skip 140 --launch-count 1 -f -o gpurun_out/alloc_v3 python tools/_exp_steps.py > gpurun_out/ncu_v3.log 2>&1
