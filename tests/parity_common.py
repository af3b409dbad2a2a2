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
