ncu --set full --import-source on --clock-control none -k regex:env_feature --launch-skip 140 --launch-count 1 -f -o gpurun_out/feat_v1 python tools/_exp_steps.py > gpurun_out/ncu_f1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:env_place --launch-skip 140 --launch-count 1 -f -o gpurun_out/place_v1 python tools/_exp_steps.py > gpurun_out/ncu_p1.log 2>&1
