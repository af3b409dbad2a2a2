"""Single-agent view of the inventory environment for the centralised (CPPO) setting.

Mirrors the reference's ``CentralizedEnvWrapper`` (reference: src/environment/envs/single_env.py:25-270): the
observation is the global state (all warehouses' local vectors concatenated, ``n_warehouses * local_obs_dim``), the
action the flat ``n_warehouses * n_skus`` vector in [-1, 1], the reward the sum of the per-warehouse rewards. On the
device the global state is simply ``obs.reshape(W * D)`` of the step kernel's output - nothing is copied twice.
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Tuple

import numpy as np

from ..config.schema import EnvironmentConfig
from .multi_env import InventoryEnvironment, _box


class CentralizedEnvWrapper:
    metadata = {"render_modes": ["human"], "name": "single_env"}

    def __init__(self, env_config: EnvironmentConfig, seed: Optional[int] = None, env_meta: Optional[Dict[str, Any]] = None,
                 device=None):
        self.env = InventoryEnvironment(env_config, seed=seed, env_meta=env_meta, device=device)
        self.env_config = env_config
        self.n_warehouses, self.n_skus = self.env.n_warehouses, self.env.n_skus
        self._local_obs_dim = self.env._compute_local_obs_dim()
        self._global_obs_dim = self.n_warehouses * self._local_obs_dim
        self.observation_space = _box(-np.inf, np.inf, (self._global_obs_dim,), np.float32)
        self.action_space = _box(-1.0, 1.0, (self.n_warehouses * self.n_skus,), np.float32)

    def reset(self, *, seed: Optional[int] = None, options: Optional[Dict[str, Any]] = None) -> Tuple[np.ndarray, Dict[str, Any]]:
        obs_dict, info_dict = self.env.reset(seed=seed, options=options)
        return self._extract_global_obs(obs_dict), info_dict.get(self.env.agents[0], {})

    def step(self, action: np.ndarray) -> Tuple[np.ndarray, float, bool, bool, Dict[str, Any]]:
        obs_dict, rewards, terminations, truncations, info_dict = self.env.step(self._split_action(action))
        return (self._extract_global_obs(obs_dict), sum(rewards.values()), all(terminations.values()),
                all(truncations.values()), info_dict.get(self.env.agents[0], {}))

    def render(self):
        self.env.render()

    def close(self):
        pass

    # ---- forwarded from the inner environment (single_env.py:172-218)
    @property
    def collect_step_info(self) -> bool:
        return self.env.collect_step_info

    @collect_step_info.setter
    def collect_step_info(self, value: bool):
        self.env.collect_step_info = value

    @property
    def agents(self):
        return self.env.agents

    @property
    def episode_length(self) -> int:
        return self.env.episode_length

    @property
    def max_expected_lead_time(self) -> int:
        return self.env.max_expected_lead_time

    @property
    def feature_config(self):
        return self.env.feature_config

    @property
    def include_warehouse_id(self) -> bool:
        return self.env.include_warehouse_id

    @property
    def rolling_window(self) -> int:
        return self.env.rolling_window

    @property
    def obs_normalization(self):
        return self.env.obs_normalization

    @property
    def obs_stats(self):
        return self.env.obs_stats

    # ---- helpers (single_env.py:225-270)
    def _extract_global_obs(self, obs_dict: Dict[str, np.ndarray]) -> np.ndarray:
        """Every agent's observation is ``[local_i | global]``; the global part is the same for all of them."""
        return obs_dict[self.env.agents[0]][self._local_obs_dim:]

    def _split_action(self, action: np.ndarray) -> Dict[str, np.ndarray]:
        a = np.asarray(action, dtype=np.float32).reshape(self.n_warehouses, self.n_skus)
        return {agent: a[i] for i, agent in enumerate(self.env.agents)}
