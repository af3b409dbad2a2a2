"""Per-step timings of the env step over one episode of the bench workload (large network, 65,536 envs, base-stock
replay): whole step by CUDA events, then the library's per-launch events. Also the command the ncu captures of one
steady-state launch use (STEPS=40 and --launch-skip on the step kernel's name).

    python tools/step_timings.py          # on a B200; LAYOUT=wide|compact, E=..., STEPS=...
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import bench
import marlsc_b200
from marlsc_b200.config import environment_config_from_dict
from marlsc_b200.envs import BatchedInventoryEnv
E = int(os.environ.get("E", 65536))
STEPS = int(os.environ.get("STEPS", 100))
env_dict, cfg_desc = bench.workload("large")
d = dict(env_dict); d["allow_region_mismatch"] = True
cfg = environment_config_from_dict(d)
dev = torch.device("cuda", 0)
env = BatchedInventoryEnv(cfg, E, device=dev, host_samplers=False, layout=os.environ.get("LAYOUT") or None)
demand = bench.synth_demand(env_dict, E, STEPS, dev, 99)
if env.layout == "compact":
    demand = [env.lines_from_orders(dm) for dm in demand]
    print("rounds per env-step", np.mean([dm.n_rounds for dm in demand]) / E, "lines", np.mean([dm.n_lines for dm in demand]) / E)
actions = bench.record_base_stock_actions(env, env_dict, demand, 2.0)
obs = torch.empty_like(env.obs); rew = torch.empty((E, cfg.n_warehouses), device=dev)
modes = ("plain",) if os.environ.get("PLAIN") else ("timing", "plain")
for mode in modes:
    env.reset(obs_out=obs)
    env.set_timing(mode == "timing")
    rows = []
    for t in range(STEPS):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        env.step(actions[t], orders=demand[t], obs_out=obs, rewards_out=rew)
        b.record()
        torch.cuda.synchronize()
        rows.append([a.elapsed_time(b)] + (env.last_step_timing() if mode == "timing" else []))
    env.set_timing(False)
    r = np.array(rows)
    print(mode, env.layout, "mean", np.round(r.mean(0), 3), "steady (t>=20)", np.round(r[20:].mean(0), 3))
    for t in (0, 1, 2, 5, 10, 15, 20, 30, 50, 70, 99):
        if t < STEPS:
            print("  t", t, np.round(r[t], 3))
print("inv mean", float(env.inventory.float().mean()))
