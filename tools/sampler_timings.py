"""Timings of the on-device demand sampler (K4), the base-stock policy kernel (K5) and a step that draws its own
demand, at the bench workload (large network, 65,536 envs).

    python tools/sampler_timings.py       # on a B200
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import bench
from marlsc_b200.config import environment_config_from_dict
from marlsc_b200.envs import BatchedInventoryEnv
from marlsc_b200.rollout import base_stock_levels
E = 65536
env_dict, _ = bench.workload("large")
d = dict(env_dict); d["allow_region_mismatch"] = True
cfg = environment_config_from_dict(d)
dev = torch.device("cuda", 0)
env = BatchedInventoryEnv(cfg, E, device=dev, host_samplers=False)
env.enable_device_demand(seed=5)
lvl = torch.from_numpy(base_stock_levels(env, 2.0, serve="cheapest")).float().to(dev)
act = torch.empty((E, 10, 100), device=dev)
env.reset()
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
for _ in range(30):
    env.base_stock_actions(lvl, out=act); env.step(act)
print("K4 sample ms", timeit(env.sample_device_demand))
print("K5 policy ms", timeit(lambda: env.base_stock_actions(lvl, out=act)))
print("K4+K1 step ms", timeit(lambda: env.step(act)))
print("mean orders", float(env._dd["counts"].float().mean()), "overflow", env.demand_overflowed())
