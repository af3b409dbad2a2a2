"""Drop-in single-environment adapter with the reference's dict API
(reference: src/environment/envs/multi_env.py:38-431): ``reset() -> (obs, infos)``,
``step(actions) -> (obs, rewards, terminations, truncations, infos)`` keyed by ``warehouse_i``,
plus the attributes the reference's callers read (``inventory``, ``_incoming_demand_home``,
``_compute_pending_matrix()``, ``timestep`` ... see SURVEY.md section 8b).

It is a ``num_envs == 1`` view of :class:`BatchedInventoryEnv`: every step still runs the fused CUDA
kernel; demand and lead times are drawn on the host from the same NumPy streams as the reference, so
a seeded instance reproduces the reference's trajectories.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch

from ..config.schema import EnvironmentConfig
from .batched_env import BatchedInventoryEnv


class _Box:
    """Minimal stand-in for ``gymnasium.spaces.Box`` (gymnasium is optional for this adapter)."""

    def __init__(self, low, high, shape, dtype):
        self.low, self.high, self.shape, self.dtype = low, high, shape, dtype


def _box(low, high, shape, dtype):
    try:
        from gymnasium.spaces import Box
        return Box(low=low, high=high, shape=shape, dtype=dtype)
    except Exception:
        return _Box(low, high, shape, dtype)


class InventoryEnvironment:
    metadata = {"render_modes": ["human"], "name": "multi_env"}

    def __init__(self, env_config: EnvironmentConfig, seed: Optional[int] = None,
                 env_meta: Optional[Dict[str, Any]] = None, device=None):
        self.env_config = env_config
        self._batch = BatchedInventoryEnv(env_config, 1, device=device, env_meta=env_meta,
                                          env_seeds=[seed], diagnostics=True)
        b = self._batch
        self.n_warehouses, self.n_skus, self.n_regions = b.n_warehouses, b.n_skus, b.n_regions
        self.episode_length = b.episode_length
        self.max_wh_capacities = env_config.max_wh_capacities
        self.feature_config = b.feature_config
        self.action_space_type = env_config.action_space.type
        self.expected_lead_times = b.expected_lead_times
        self.max_expected_lead_time = b.max_expected_lead_time
        self.home_regions = b.home_regions
        self.rolling_window = 5
        self.ema_alpha = 0.3
        self.obs_normalization = b.obs_normalization
        self.obs_stats = b.obs_stats
        self.include_warehouse_id = b.include_warehouse_id
        self._num_eval_episodes = b._num_eval_episodes
        self.seed_manager = b.seed_managers[0]
        self.demand_sampler = b.demand_samplers[0]
        self.lead_time_sampler = b.lead_time_samplers[0]
        self.agents = list(b.agents)
        self.possible_agents = list(b.agents)
        self.collect_step_info = False
        self._incoming_demand_home = np.zeros((self.n_warehouses, self.n_skus), dtype=np.float32)

    # ------------------------------------------------------------------ reference attributes
    @property
    def timestep(self) -> int:
        return self._batch.timestep

    @property
    def inventory(self) -> np.ndarray:
        return self._batch.inventory[0].to(torch.float64).cpu().numpy()

    def _compute_pending_matrix(self) -> np.ndarray:
        return self._batch.pending_matrix()[0].cpu().numpy().astype(np.float32)

    def _compute_local_obs_dim(self) -> int:
        return self._batch.obs_dim

    def observation_space(self, agent: str):
        d = self._batch.obs_dim
        return _box(-np.inf, np.inf, (d + self.n_warehouses * d,), np.float32)

    def global_observation_space(self):
        return _box(-np.inf, np.inf, (self.n_warehouses * self._batch.obs_dim,), np.float32)

    def action_space(self, agent: str):
        return _box(-1.0, 1.0, (self.n_skus,), np.float32)

    def render(self):
        pass

    # ------------------------------------------------------------------ API
    def _obs_dict(self, obs: torch.Tensor) -> Dict[str, np.ndarray]:
        full = self._batch.agent_observations(obs)[0].cpu().numpy()
        return {a: full[i] for i, a in enumerate(self.agents)}

    def reset(self, seed: Optional[int] = None, options: Optional[Dict] = None) -> Tuple[Dict, Dict]:
        obs = self._batch.reset(seed=seed)
        self._incoming_demand_home = np.zeros((self.n_warehouses, self.n_skus), dtype=np.float32)
        return self._obs_dict(obs), {a: {} for a in self.agents}

    def step(self, actions: Dict[str, np.ndarray]):
        b = self._batch
        inventory_before = self.inventory if self.collect_step_info else None
        pending_total = self._compute_pending_matrix() if self.collect_step_info else None
        act = np.stack([np.asarray(actions[a], dtype=np.float32) for a in self.agents])[None]
        orders, leads = b.sample_host_demand()
        obs, rew, truncated = b.step(torch.from_numpy(act).to(b.device), orders=orders, actual_lead=leads)
        n = orders.n_orders
        dem = np.zeros((self.n_regions, self.n_skus), dtype=np.float32)
        np.add.at(dem, orders.region[:n].astype(np.int64), orders.qty[:n].astype(np.float32))
        self._incoming_demand_home = dem[self.home_regions, :]
        r = rew[0].cpu().numpy()
        rewards = {a: float(r[i]) for i, a in enumerate(self.agents)}
        terminations = {a: False for a in self.agents}
        truncations = {a: bool(truncated) for a in self.agents}
        if self.collect_step_info:
            d = {k: v[0].cpu().numpy() for k, v in b.diag.items()}
            ship = d["ship_by_sku"].astype(np.float64)
            uniq = (orders.qty[:n] > 0).sum(axis=1)
            info = dict(
                inventory=inventory_before, pending_total=pending_total,
                order_quantities=d["ordered"].astype(np.float64), demand_per_region=dem.astype(np.float64),
                fulfilled_per_warehouse=ship.sum(axis=1), unfulfilled_demands=d["unfulfilled"].astype(np.float64),
                shipment_counts=d["ship_counts"].astype(np.int64), shipment_quantities=ship.sum(axis=2),
                shipment_quantities_by_sku=ship, lost_order_counts=d["lost_orders"].astype(np.int64),
                lost_sales=d["lost_sales"].astype(np.float64), n_orders=n,
                mean_unique_skus_per_order=float(uniq.mean()) if n else 0.0,
                holding_cost=d["cost_breakdown"][:, 0].astype(np.float64), penalty_cost=d["cost_breakdown"][:, 1].astype(np.float64),
                outbound_shipment_cost=d["cost_breakdown"][:, 2].astype(np.float64),
                inbound_shipment_cost=d["cost_breakdown"][:, 3].astype(np.float64))
            infos = {a: info for a in self.agents}
        else:
            infos = {a: {} for a in self.agents}
        return self._obs_dict(obs), rewards, terminations, truncations, infos
