"""Shared pieces of the trajectory-parity tests (host emulation on CPU, CUDA library on the GPU)."""
from __future__ import annotations

from typing import Dict

import numpy as np

import marlsc_b200  # noqa: F401
from golden_io import Golden
from marlsc_b200.config import environment_config_from_dict
from marlsc_b200.demand import OrderBatch, pack_orders
from marlsc_b200.spec import build_env_spec

INT_OUT = ("inventory", "ordered", "unfulfilled", "ship_counts", "ship_by_sku", "lost_orders", "trunc")
# float32 results against the reference's float64: rtol from BASELINE.json's north_star
RTOL, ATOL = 1e-5, 1e-6


def spec_for(g: Golden, region_map=None):
    env = dict(g.env)
    if g.meta["allow_region_mismatch"]:
        env["allow_region_mismatch"] = True
    cfg = environment_config_from_dict(env)
    return cfg, build_env_spec(cfg, obs_normalization=g.meta["obs_normalization"], obs_stats=g.obs_stats,
                               include_warehouse_id=g.meta["include_warehouse_id"], region_map=region_map)


def step_orders(g: Golden, t: int, region_shift=None) -> OrderBatch:
    """All environments' orders of step t in the kernel's CSR format."""
    per_env = []
    for i in range(g.N):
        orders = g.orders(i, t)
        if region_shift is not None:
            orders = [(r + region_shift(i, t, j), q) for j, (r, q) in enumerate(orders)]
        per_env.append(orders)
    return pack_orders(per_env, g.S)


def compare_step(g: Golden, t: int, out: Dict[str, np.ndarray], what: str = ""):
    """``out`` holds [N, ...] arrays for step t produced by the implementation under test."""
    exp_int = {k: g[k][:, t] for k in ("inventory", "ordered", "unfulfilled", "ship_counts", "ship_by_sku", "lost_orders")
               if k in g.z}                              # "lite" goldens do not carry the [W,R,S] shipment cube
    for k, exp in exp_int.items():
        if k in out:
            assert np.array_equal(np.asarray(out[k]).astype(np.int64), exp.astype(np.int64)), \
                f"{what}{g.name} step {t}: integer mismatch in {k}"
    if "trunc" in out:
        assert np.array_equal(np.asarray(out["trunc"]).astype(bool), np.broadcast_to(g["trunc"][:, t], (g.N,)))
    cb = out.get("cost_breakdown")
    exp_f = dict(rewards=g["rewards"][:, t], obs=g["obs_local"][:, t], lost_sales=g["lost_sales"][:, t])
    if cb is not None:
        exp_f.update(cost_hold=g["cost_hold"][:, t], cost_pen=g["cost_pen"][:, t], cost_out=g["cost_out"][:, t],
                     cost_in=g["cost_in"][:, t])
        out = dict(out, cost_hold=cb[..., 0], cost_pen=cb[..., 1], cost_out=cb[..., 2], cost_in=cb[..., 3])
    for k, exp in exp_f.items():
        if k in out:
            np.testing.assert_allclose(np.asarray(out[k], dtype=np.float64), exp, rtol=RTOL, atol=ATOL,
                                       err_msg=f"{what}{g.name} step {t}: {k}")


def _lean_env_dict(rng, W, S, R, lost, pen_uniform, lead_hi, scope="agent", demand_home=False, qmax_hi=30):
    out_var = np.stack([rng.permutation(W) for _ in range(R)], 1) * 0.05 + 0.05          # tie-free -> static priority
    pen = 4.0 if pen_uniform else [float(x) for x in rng.integers(1, 9, S)]
    lead = rng.integers(1, lead_hi + 1, (W, S))
    lead[0, 0] = lead_hi                                                                  # some cell has the longest lead
    return dict(
        action_space=dict(type="direct", params=dict(max_order_quantities=[int(x) for x in rng.integers(5, qmax_hi, S)])),
        n_warehouses=W, n_skus=S, n_regions=R, episode_length=9, max_wh_capacities=[1e7] * W,
        initial_inventory=dict(type="custom", params=dict(values=6)),
        cost_structure=dict(
            holding_cost=0.5, penalty_cost=pen,
            shipment_cost=dict(outbound_fixed=np.zeros((W, R)).tolist(), outbound_variable=out_var.tolist(),
                               inbound_fixed=np.full((W, S), 0.5).tolist(), inbound_variable=np.full((W, S), 0.25).tolist()),
            sku_weights=[1.0] * S, distances=(rng.integers(10, 500, (W, R)) * 1.0).tolist()),
        components=dict(
            demand_sampler=dict(type="poisson", params=dict(lambda_orders=1.0, probability_skus=0.3, lambda_quantity=5.0)),
            demand_allocator=dict(type="greedy", params=dict(max_splits=W - 1)),
            lead_time_sampler=dict(type="fixed", params=dict(expected_lead_times=lead.tolist())),
            lost_sales_handler=dict(type=lost, params=None),
            reward_calculator=dict(type="cost", params=dict(scope=scope, scale_factor=0.1, cost_weights=[0.25] * 4))),
        data_source=dict(type="custom"),
        features=dict(inventory=True, pipeline=True, incoming_demand_home=demand_home, units_shipped_home=False, units_shipped_away=False,
                      stockout=False, rolling_demand_mean=True, demand_forecast=False, days_of_supply=False,
                      net_inventory_position=False, demand_variability=False, demand_history=False, inventory_aggregate=True,
                      pipeline_aggregate=False, incoming_demand_home_aggregate=False, units_shipped_away_aggregate=False,
                      rolling_demand_mean_aggregate=False, demand_forecast_aggregate=False))
