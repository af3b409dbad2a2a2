"""Lost-sales handlers (reference: src/environment/components/lost_sales_handler.py:10-50, 53-92,
95-148, 151-210). The attribution itself happens in the fused step kernel (env_core.cuh, phase 3)."""
from __future__ import annotations

from typing import Any, Dict

import numpy as np

from ..config.schema import LostSalesHandlerConfig
from ..context import EnvironmentContext
from .base import DeviceComponent


class BaseLostSalesHandler(DeviceComponent):
    kind = 0

    def __init__(self, context: EnvironmentContext, component_config: LostSalesHandlerConfig):
        self.n_warehouses = context.n_warehouses
        self.n_skus = context.n_skus
        self.n_regions = context.n_regions
        self.distances = context.distances
        self.closest_warehouses = np.argmin(self.distances, axis=0)   # first minimum, like the reference
        self.alpha = 1.0

    def spec_fields(self) -> Dict[str, Any]:
        return dict(lost_sales_type=self.kind, lost_alpha=float(self.alpha),
                    closest_wh=self.closest_warehouses.astype(np.int32))


class ClosestLostSalesHandler(BaseLostSalesHandler):
    kind = 0


class ShipmentLostSalesHandler(BaseLostSalesHandler):
    kind = 1


class CostLostSalesHandler(BaseLostSalesHandler):
    kind = 2

    def __init__(self, context: EnvironmentContext, component_config: LostSalesHandlerConfig):
        super().__init__(context, component_config)
        self.alpha = float(component_config.params["alpha"])
        if not self.alpha > 0.0:
            raise ValueError("lost_sales_handler 'cost' needs alpha > 0")
