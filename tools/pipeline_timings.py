"""All-device pipeline of the bench workload (K4 -> K1 with the in-step base-stock heuristic), per-kernel CUDA-event times.

    python tools/pipeline_timings.py      # on a B200; E=..., STEPS=...
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, numpy as np
import bench
import marlsc_b200
from marlsc_b200.config import environment_config_from_dict
from marlsc_b200.envs import BatchedInventoryEnv
from marlsc_b200.rollout import base_stock_levels
E = int(os.environ.get("E", 65536))
STEPS = int(os.environ.get("STEPS", 40))
env_dict, _ = bench.workload("large")
cfg = environment_config_from_dict(dict(env_dict, allow_region_mismatch=True))
dev = torch.device("cuda", 0)
env = BatchedInventoryEnv(cfg, E, device=dev, host_samplers=False, device_demand=True, demand_seed=5)
lvl = torch.from_numpy(base_stock_levels(env, 2.0, serve="cheapest")).float().to(dev)
obs = torch.empty_like(env.obs); rew = torch.empty((E, cfg.n_warehouses), device=dev)
env.reset(obs_out=obs)
env._dd["overlap"] = True
ts = []
for t in range(STEPS):
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    env.step(None, obs_out=obs, rewards_out=rew, base_stock_level=lvl)
    c.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(c))
print("K4 (overlapped on a side stream) + step: mean of steps 20..", np.round(np.mean(ts[20:]), 3), "ms; overflow", env.demand_overflowed())
env._dd["overlap"] = False
ts = []
for t in range(STEPS):
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    env.step(None, obs_out=obs, rewards_out=rew, base_stock_level=lvl)
    c.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(c))
print("K4 then step on one stream:", np.round(np.mean(ts[20:]), 3), "ms")
