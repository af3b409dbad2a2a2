"""Seed bookkeeping with the reference's derivation rules, so host-side sampling (initial inventory,
Poisson demand, lead-time deviations) replays the reference's NumPy streams draw for draw.

Reference: src/utils/seed_manager.py - registries :8-31, spawn :203-224, per-episode derivation
:100-120, per-env derivation :166-186.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
from numpy.random import SeedSequence

EXPERIMENT_SEEDS: Tuple[str, ...] = ("data_weights", "data_distances", "data_costs", "train", "eval", "obs_stats")
ENVIRONMENT_SEEDS: Tuple[str, ...] = ("preprocessing", "inventory", "demand_sampler", "lead_time_sampler")
STOCHASTIC_SEEDS: Tuple[str, ...] = ("demand_sampler", "lead_time_sampler")


def _first_word(ss: SeedSequence) -> int:
    return int(ss.generate_state(1, dtype=np.uint32)[0])


class SeedManager:
    def __init__(self, root_seed: Optional[int] = None, seed_registry: Tuple[str, ...] = EXPERIMENT_SEEDS):
        self.root_seed = root_seed
        self._original_root_seed = root_seed
        self._episode_counter = 0
        self._seed_registry = seed_registry
        self._seed_sequences: Dict[str, Optional[SeedSequence]] = {}
        self._spawn_seeds()

    def _spawn_seeds(self) -> None:
        if self.root_seed is None:
            self._seed_sequences = {n: None for n in self._seed_registry}
        else:
            kids = SeedSequence(self.root_seed).spawn(len(self._seed_registry))
            self._seed_sequences = dict(zip(self._seed_registry, kids))

    def _get_seed_sequence(self, name: str) -> Optional[SeedSequence]:
        if name not in self._seed_registry:
            raise ValueError(f"Seed '{name}' not in registry {self._seed_registry}")
        return self._seed_sequences[name]

    def get_rng(self, name: str) -> np.random.Generator:
        return np.random.default_rng(self._get_seed_sequence(name))

    def get_seed_int(self, name: str) -> Optional[int]:
        ss = self._get_seed_sequence(name)
        return None if ss is None else _first_word(ss)

    def advance_episode(self) -> None:
        if self._original_root_seed is None:
            return
        self.root_seed = _first_word(SeedSequence([self._original_root_seed, self._episode_counter]))
        self._spawn_seeds()
        self._episode_counter += 1

    def update_root_seed(self, root_seed: Optional[int]) -> None:
        self.root_seed = self._original_root_seed = root_seed
        self._episode_counter = 0
        self._spawn_seeds()

    def spawn_child_seeds(self, name: str, n: int) -> List[Optional[int]]:
        ss = self._get_seed_sequence(name)
        return [None] * n if ss is None else [_first_word(c) for c in ss.spawn(n)]

    @staticmethod
    def derive_env_seed(base_seed: int, worker_index: int, env_index: int) -> int:
        return _first_word(SeedSequence([base_seed, worker_index, env_index]))
