"""Observation statistics from a random policy, on the device.

Replaces the reference's ``compute_obs_statistics`` (reference: src/utils/obs_stats.py:11-90, grouped variant
:93-168): it steps ONE CPU environment through ``n_episodes`` episodes under uniform random actions and takes the
mean / population std of every agent's local observation (at reset and after every step). Here the ``n_episodes``
episodes run side by side as one batched environment (K4 demand, K1 step) and the per-column sums are accumulated
in float64 on the device, so the sample count - ``n_episodes * (episode_length + 1) * n_warehouses`` - and the
estimator are the reference's; the random streams are not (Philox demand, torch actions), so the numbers agree
statistically, not bit for bit.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from ..config.schema import EnvironmentConfig
from ..envs import BatchedInventoryEnv


def _groups(env: BatchedInventoryEnv):
    """(per-SKU column count, has aggregate column) per feature group, in observation order - the groups the
    reference pools (obs_stats.py:120-145; the four extended features are left at mean 0 / std 1 there too)."""
    f, S, L = env.feature_config, env.n_skus, env.max_expected_lead_time
    out = []
    if f.inventory:
        out.append((S, f.inventory_aggregate))
    if f.pipeline:
        out.append((L * S, f.pipeline_aggregate))
    if f.incoming_demand_home:
        out.append((S, f.incoming_demand_home_aggregate))
    if f.units_shipped_home:
        out.append((S, False))
    if f.units_shipped_away:
        out.append((S, f.units_shipped_away_aggregate))
    if f.stockout:
        out.append((S, False))
    if f.rolling_demand_mean:
        out.append((S, f.rolling_demand_mean_aggregate))
    if f.demand_forecast:
        out.append((S, f.demand_forecast_aggregate))
    return out


def compute_obs_statistics(env_config: EnvironmentConfig, seed_manager, mode: str = "meanstd_custom", n_episodes: int = 10,
                           device: Optional[torch.device] = None) -> Tuple[np.ndarray, np.ndarray]:
    """``(obs_mean, obs_std)``, float32 ``[local_obs_dim]``; columns with std < 1e-8 get std 1 (obs_stats.py:78).
    ``mode``: ``"meanstd_custom"`` (per column) or ``"meanstd_grouped"`` (per-SKU columns of a feature share one pair)."""
    if mode not in ("meanstd_custom", "meanstd_grouped"):
        raise ValueError("mode must be 'meanstd_custom' or 'meanstd_grouped'")
    if n_episodes < 1:
        raise ValueError("n_episodes must be positive")
    env_seed, action_seed = seed_manager.spawn_child_seeds("obs_stats", 2)
    env = BatchedInventoryEnv(env_config, n_episodes, device=device, host_samplers=False)
    env.enable_device_demand(seed=0 if env_seed is None else int(env_seed))
    if env.stochastic_lead:
        env.enable_device_leads(seed=1 if env_seed is None else int(env_seed) + 1)
    gen = torch.Generator(device=env.device)
    gen.manual_seed(0 if action_seed is None else int(action_seed))
    E, W, S, D = env.num_envs, env.n_warehouses, env.n_skus, env.obs_dim
    s1 = torch.zeros(D, dtype=torch.float64, device=env.device)
    s2 = torch.zeros(D, dtype=torch.float64, device=env.device)

    def add(obs: torch.Tensor) -> None:
        x = obs.reshape(E * W, D).to(torch.float64)
        s1.add_(x.sum(0))
        s2.add_((x * x).sum(0))

    add(env.reset())
    for _ in range(env.episode_length):
        act = torch.rand((E, W, S), device=env.device, generator=gen) * 2.0 - 1.0
        obs, _, _ = env.step(act)
        add(obs)
    if env.demand_overflowed():
        raise RuntimeError("device demand buffer overflowed while estimating observation statistics")
    n = float(E * W * (env.episode_length + 1))
    c1, c2 = s1.cpu().numpy(), s2.cpu().numpy()
    env.close()
    if mode == "meanstd_custom":
        mean = c1 / n
        std = np.sqrt(np.maximum(c2 / n - mean * mean, 0.0))
    else:
        mean, std = np.zeros(D), np.ones(D)
        i = 0
        for cols, has_agg in _groups(env):
            m = c1[i:i + cols].sum() / (n * cols)
            mean[i:i + cols] = m
            std[i:i + cols] = np.sqrt(max(c2[i:i + cols].sum() / (n * cols) - m * m, 0.0))
            i += cols
            if has_agg:
                mean[i] = c1[i] / n
                std[i] = np.sqrt(max(c2[i] / n - mean[i] * mean[i], 0.0))
                i += 1
    std = np.where(std < 1e-8, 1.0, std)
    return mean.astype(np.float32), std.astype(np.float32)
