"""Demand allocators (reference: src/environment/components/demand_allocator.py:12-38, 41-88, 91-217).

The greedy cheapest-first allocation with order splitting runs inside the fused step kernel
(csrc/env_core.cuh, phase 2); the Python class only validates and forwards its parameter.
"""
from __future__ import annotations

from typing import Any, Dict

from ..config.schema import DemandAllocatorConfig
from ..context import EnvironmentContext
from .base import DeviceComponent


class BaseDemandAllocator(DeviceComponent):
    def __init__(self, context: EnvironmentContext, component_config: DemandAllocatorConfig):
        self.n_warehouses = context.n_warehouses
        self.n_skus = context.n_skus
        self.n_regions = context.n_regions
        self.fixed_cost_per_order = context.shipment_cost.outbound_fixed
        self.variable_cost_per_weight = context.shipment_cost.outbound_variable
        self.sku_weights = context.sku_weights


class GreedyDemandAllocator(BaseDemandAllocator):
    def __init__(self, context: EnvironmentContext, component_config: DemandAllocatorConfig):
        super().__init__(context, component_config)
        ms = component_config.params["max_splits"]
        self.max_splits = self.n_warehouses - 1 if ms == "default" else int(ms)

    def spec_fields(self) -> Dict[str, Any]:
        return dict(max_splits=self.max_splits)
