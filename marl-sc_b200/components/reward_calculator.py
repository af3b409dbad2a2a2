"""Reward calculators (reference: src/environment/components/reward_calculator.py:10-56, 59-190).
Costs and rewards are computed in the fused step kernel (env_core.cuh, phases 3-4)."""
from __future__ import annotations

from typing import Any, Dict

import numpy as np

from ..config.schema import RewardCalculatorConfig
from ..context import EnvironmentContext
from .base import DeviceComponent


class BaseRewardCalculator(DeviceComponent):
    def __init__(self, context: EnvironmentContext, component_config: RewardCalculatorConfig):
        self.n_warehouses = context.n_warehouses
        self.n_skus = context.n_skus
        self.n_regions = context.n_regions
        self.holding_cost = context.holding_cost
        self.penalty_cost = context.penalty_cost
        self.sku_weights = context.sku_weights


class CostRewardCalculator(BaseRewardCalculator):
    def __init__(self, context: EnvironmentContext, component_config: RewardCalculatorConfig):
        super().__init__(context, component_config)
        p = component_config.params
        get = (lambda k: getattr(p, k)) if hasattr(p, "scope") else (lambda k: p[k])
        self.scope = get("scope")
        self.scale_factor = float(get("scale_factor"))
        self.cost_weights = np.array(get("cost_weights"), dtype=float)   # validated, unused (reference :154-160)

    def _rate(self, cost) -> np.ndarray:
        # list -> per-SKU rate as is; scalar -> rate * sku weight (reference :128-137)
        return np.asarray(cost, dtype=float) if isinstance(cost, np.ndarray) else self.sku_weights * float(cost)

    def spec_fields(self) -> Dict[str, Any]:
        return dict(reward_scope=1 if self.scope == "team" else 0, scale_factor=self.scale_factor,
                    hold_rate=self._rate(self.holding_cost), pen_rate=self._rate(self.penalty_cost))
