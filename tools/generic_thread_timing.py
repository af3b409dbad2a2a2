"""Step time of the reference's shipped 3 WH x 5 SKU config (demand_centered actions: the generic kernel instantiation,
a thread per environment) next to the same network with direct actions (lean instantiation).   E=..., STEPS=..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import marlsc_b200
from golden.scenarios import small_default
from marlsc_b200.config import environment_config_from_dict
from marlsc_b200.envs import BatchedInventoryEnv

E, STEPS = int(os.environ.get("E", 65536)), int(os.environ.get("STEPS", 60))


def cfg(kind):
    d = small_default()
    S = 5
    d.update(n_skus=S, initial_inventory=dict(type="custom", params=dict(values=[[60] * S] * 3)))
    d["action_space"] = (dict(type="demand_centered", params=dict(max_quantity_adjustment=[20] * S)) if kind == "demand_centered"
                         else dict(type="direct", params=dict(max_order_quantities=[40] * S)))
    cs = d["cost_structure"]
    cs["penalty_cost"], cs["sku_weights"] = [5] * S, [1.0] * S
    cs["shipment_cost"]["inbound_fixed"], cs["shipment_cost"]["inbound_variable"] = [[0] * S] * 3, [[1.0] * S] * 3
    d["components"]["demand_sampler"]["params"]["lambda_quantity"] = [[5] * S] * 3
    d["components"]["lead_time_sampler"]["params"]["expected_lead_times"] = [[3] * S] * 3
    return environment_config_from_dict(d)


for kind in ("direct", "demand_centered"):
    env = BatchedInventoryEnv(cfg(kind), E, device="cuda:0", host_samplers=False, device_demand=True, demand_seed=1)
    env.reset()
    gen = torch.Generator(device="cuda:0").manual_seed(0)
    acts = [torch.rand((E, 3, 5), device="cuda:0", generator=gen) * 2 - 1 for _ in range(8)]
    ts = []
    for t in range(STEPS):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        env.step(acts[t % 8])
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(kind, "team", env.team_size, "layout", env.layout, "ms/step (K4 + K1)", round(float(np.mean(ts[10:])), 4),
          "M agent-steps/s", round(E * 3 / np.mean(ts[10:]) / 1e3, 1))
    env.close()
