"""Batched heuristic baselines on the device (reference: src/experiments/run_baselines.py:73-293, 394-455).

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


def random_actions(env: BatchedInventoryEnv, generator: Optional[torch.Generator] = None,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Uniform actions in [-1, 1) for every (environment, warehouse, SKU) - run_baselines.py:73-94."""
    shape = (env.num_envs, env.n_warehouses, env.n_skus)
    a = torch.rand(shape, device=env.device, generator=generator) * 2.0 - 1.0
    if out is not None:
        out.copy_(a)
        return out
    return a


def constant_actions(env: BatchedInventoryEnv, quantities: np.ndarray) -> torch.Tensor:
    """The action that orders fixed ``quantities [W,S]`` every step (direct action space): clip to the order
    maxima, ``2 q / max - 1`` in float32 - run_baselines.py:96-130. Returned as ``[E,W,S]`` (a broadcast view)."""
    if env.env_config.action_space.type != "direct":
        raise ValueError("the constant-order baseline assumes the direct action space")
    mx = np.asarray(env.env_config.action_space.params.max_order_quantities, dtype=float)
    q = np.clip(np.asarray(quantities, dtype=float), 0.0, mx)
    if q.shape != (env.n_warehouses, env.n_skus):
        raise ValueError(f"quantities must have shape {(env.n_warehouses, env.n_skus)}")
    a = torch.from_numpy((2.0 * q / mx - 1.0).astype(np.float32)).to(env.device)
    return a.unsqueeze(0).expand(env.num_envs, -1, -1).contiguous()


class AdaptiveBaseStock:
    """Rolling-mean base-stock heuristic ("BS-Adaptive", run_baselines.py:209-293) for every environment: each step
    record the home-region demand of the step that just ended, take its mean and population variance over the last
    ``H`` observations (variance := mean while there is a single one), set ``S = L mean + z sqrt(L var)`` per
    (environment, warehouse, SKU) and order up to it through the policy kernel (K5, per-environment levels). Orders
    nothing at t = 0. The demand comes from the environment's history planes, so the configuration must keep them
    (rolling-mean feature on)."""

    def __init__(self, env: BatchedInventoryEnv, z: float, H: int):
        if env.demand_hist is None:
            raise ValueError("the adaptive base-stock baseline needs the home-demand history (rolling_demand_mean feature)")
        if H < 1:
            raise ValueError("H must be positive")
        self.env, self.z, self.H = env, float(z), int(H)
        self.lead = torch.from_numpy(np.asarray(env.expected_lead_times, dtype=np.float32)).to(env.device)
        self.buf = torch.zeros((self.H, env.num_envs, env.n_warehouses, env.n_skus), device=env.device)
        self.count = 0
        self._prev_t = -1

    def actions(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        env, t = self.env, self.env.timestep
        if t <= self._prev_t or t == 0:
            self.count = 0
        self._prev_t = t
        if t > 0:                                   # demand of step t - 1 sits in history plane (t - 1) mod 5
            self.buf[self.count % self.H] = env.demand_hist[:, (t - 1) % env.rolling_window].to(torch.float32)
            self.count += 1
        shape = (env.num_envs, env.n_warehouses, env.n_skus)
        if self.count == 0:
            a = torch.full(shape, -1.0, device=env.device) if out is None else out.fill_(-1.0)
            return a
        n = min(self.count, self.H)
        win = self.buf[:n]
        mean = win.mean(0)
        var = win.var(0, unbiased=False) if n > 1 else mean
        level = self.lead * mean + self.z * torch.sqrt(self.lead * var)
        return env.base_stock_actions(level, out=out)


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
