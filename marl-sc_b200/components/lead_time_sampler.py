"""Lead-time samplers (reference: src/environment/components/lead_time_sampler.py:34-62, 97-131, 169-223)."""
from __future__ import annotations

from typing import Any, Dict, Optional

import numpy as np

from ..config.schema import LeadTimeSamplerConfig
from ..context import EnvironmentContext
from .base import StochasticComponent


class BaseLeadTimeSampler(StochasticComponent):
    kind = 0

    def __init__(self, context: EnvironmentContext, component_config: LeadTimeSamplerConfig):
        self.n_skus = context.n_skus
        self.n_warehouses = context.n_warehouses
        self._rng = np.random.default_rng()
        self.expected_lead_times = np.array(component_config.params.expected_lead_times, dtype=int)

    def sample(self) -> np.ndarray:
        raise NotImplementedError

    def get_expected(self) -> np.ndarray:
        return self.expected_lead_times.copy()

    def get_max_expected(self) -> int:
        return int(self.expected_lead_times.max())

    def max_actual(self) -> int:
        return self.get_max_expected()

    def reset(self, rng: Optional[np.random.Generator] = None):
        self._rng = rng if rng is not None else np.random.default_rng()

    def spec_fields(self) -> Dict[str, Any]:
        return dict(lead_mode=self.kind, expected_lead=self.expected_lead_times.astype(np.int32),
                    max_expected_lead=self.get_max_expected(), ring_depth=max(self.max_actual(), 1))


class FixedLeadTimeSampler(BaseLeadTimeSampler):
    kind = 0

    def sample(self) -> np.ndarray:
        return self.expected_lead_times.copy()


class StochasticLeadTimeSampler(BaseLeadTimeSampler):
    """expected + U{-d..+d} per cell, clipped to >= 1; per-SKU ``d`` lists draw one column at a time
    exactly like the reference (:179-195)."""
    kind = 1

    def __init__(self, context: EnvironmentContext, component_config: LeadTimeSamplerConfig):
        super().__init__(context, component_config)
        md = component_config.params.deviation.max_deviation
        self.max_deviation = np.array(md, dtype=int) if isinstance(md, list) else int(md)

    def max_actual(self) -> int:
        return self.get_max_expected() + int(np.max(self.max_deviation))

    def sample(self) -> np.ndarray:
        W, S = self.n_warehouses, self.n_skus
        if isinstance(self.max_deviation, np.ndarray):
            dev = np.column_stack([self._rng.integers(-self.max_deviation[s], self.max_deviation[s] + 1, size=W)
                                   for s in range(S)])
        else:
            dev = self._rng.integers(-self.max_deviation, self.max_deviation + 1, size=(W, S))
        return np.maximum(1, self.expected_lead_times + dev)
