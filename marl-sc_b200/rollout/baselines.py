"""Batched heuristic baselines on the device (reference: src/experiments/run_baselines.py:133-207, 394-455).

``base_stock_levels`` restates the reference's newsvendor base-stock level
``S[w,k] = L*E[D] + z*sqrt(L*E[D])`` with ``E[D] = lambda_orders * probability_skus * lambda_quantity`` of the
warehouse's home region; ``baseline_rollout`` runs whole episodes of it for E environments with the
policy kernel (K5) and the fused step kernel (K1) - no host round trip per step.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from ..envs import BatchedInventoryEnv


def base_stock_levels(env: BatchedInventoryEnv, z: float, serve: str = "home") -> np.ndarray:
    """[W,S] base-stock levels. ``serve="home"`` is the reference's rule (demand of the home region);
    ``serve="cheapest"`` sums the demand of every region the warehouse serves first, which is the
    sensible variant when there are more regions than warehouses."""
    smp = env.spec.components["demand_sampler"]
    lam_o, prob, lam_q = smp.dense_params()
    W, S, R = env.n_warehouses, env.n_skus, env.n_regions
    lead = env.expected_lead_times.astype(float)
    ed = np.zeros((W, S))
    if serve == "home":
        for w in range(W):
            h = int(env.home_regions[w])
            ed[w] = lam_o[h] * prob[h] * lam_q[h]
    else:
        first = np.argsort(env.spec.tables["out_var"], axis=0, kind="stable")[0]
        for r in range(R):
            ed[first[r]] += lam_o[r] * prob[r] * lam_q[r]
    return lead * ed + z * np.sqrt(lead * ed)


def baseline_rollout(env: BatchedInventoryEnv, level: np.ndarray, num_episodes: int = 1,
                     orders_fn=None) -> torch.Tensor:
    """Episode returns [num_episodes, E, W] of the base-stock policy. Demand comes from ``orders_fn(t)``,
    the device sampler or the host samplers, whichever the environment is set up with."""
    lvl = torch.as_tensor(level, dtype=torch.float32, device=env.device)
    act = torch.empty((env.num_envs, env.n_warehouses, env.n_skus), device=env.device)
    returns: List[torch.Tensor] = []
    for _ in range(num_episodes):
        env.reset()
        total = torch.zeros((env.num_envs, env.n_warehouses), device=env.device, dtype=torch.float64)
        done = False
        while not done:
            env.base_stock_actions(lvl, out=act)
            orders = orders_fn(env.timestep) if orders_fn is not None else None
            _, rew, done = env.step(act, orders=orders)
            total += rew
        returns.append(total)
    return torch.stack(returns)
