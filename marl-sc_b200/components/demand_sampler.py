"""Demand samplers (reference: src/environment/components/demand_sampler.py:12-24, 27-71, 74-163).

``PoissonDemandSampler.sample`` draws from its ``numpy`` generator in exactly the reference's order
(order count per region, SKU mask, masked quantities), so a sampler seeded through ``SeedManager``
emits the same orders as the reference's. The orders are packed into the CSR order tensors the
kernel consumes (``marlsc_b200.demand.pack_orders``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from ..config.schema import DemandSamplerConfig
from ..context import EnvironmentContext
from .base import StochasticComponent


@dataclass
class Order:
    region_id: int
    sku_demands: np.ndarray   # [S]


class BaseDemandSampler(StochasticComponent):
    def __init__(self, context: EnvironmentContext, component_config: DemandSamplerConfig):
        self.n_regions = context.n_regions
        self.n_skus = context.n_skus
        self._rng = np.random.default_rng()

    def sample(self, timestep: int) -> List[Order]:
        raise NotImplementedError

    def reset(self, rng: Optional[np.random.Generator] = None):
        self._rng = rng if rng is not None else np.random.default_rng()


class PoissonDemandSampler(BaseDemandSampler):
    def __init__(self, context: EnvironmentContext, component_config: DemandSamplerConfig):
        super().__init__(context, component_config)
        p = component_config.params
        self.per_region = isinstance(p["lambda_orders"], list)
        if self.per_region:
            self.lambda_orders = np.array(p["lambda_orders"], dtype=float)
            self.probability_skus = np.array(p["probability_skus"], dtype=float)
            self.lambda_quantity = np.array(p["lambda_quantity"], dtype=float)
        else:
            self.lambda_orders = float(p["lambda_orders"])
            self.probability_skus = float(p["probability_skus"])
            self.lambda_quantity = float(p["lambda_quantity"])

    def dense_params(self):
        """(lambda_orders [R], probability_skus [R], lambda_quantity [R,S]) broadcast to full shape."""
        R, S = self.n_regions, self.n_skus
        return (np.broadcast_to(np.asarray(self.lambda_orders, dtype=float), (R,)).copy(),
                np.broadcast_to(np.asarray(self.probability_skus, dtype=float), (R,)).copy(),
                np.broadcast_to(np.asarray(self.lambda_quantity, dtype=float), (R, S)).copy())

    def sample(self, timestep: int) -> List[Order]:
        out: List[Order] = []
        rng, S = self._rng, self.n_skus
        for r in range(self.n_regions):
            lam_o = self.lambda_orders[r] if self.per_region else self.lambda_orders
            prob = self.probability_skus[r] if self.per_region else self.probability_skus
            for _ in range(rng.poisson(lam_o)):
                picked = np.where(rng.random(S) < prob)[0]
                qty = np.zeros(S, dtype=float)
                if len(picked) > 0:
                    if self.per_region:
                        draws = rng.poisson(self.lambda_quantity[r, picked])
                    else:
                        draws = rng.poisson(self.lambda_quantity, size=len(picked))
                    qty[picked] = np.maximum(1, draws)
                out.append(Order(region_id=r, sku_demands=qty))
        return out


class EmpiricalDemandSampler(BaseDemandSampler):
    """Replays a random contiguous window of a preprocessed demand frame (reference :166-271): the window start is one
    ``rng.integers(0, n_timesteps - episode_length + 1)`` draw per episode (:226-229), the orders of a step are the frame's
    rows of that timestep grouped by ``(region_id, order_id)`` (:247-258). The frame is grouped once at construction
    (``marlsc_b200.data.pack_demand_frame``); ``sample`` only slices. The reference's raw data set is not shipped, so the
    frame comes from the caller (``env_meta["preprocessed_data"]``); without one construction fails like the reference's."""

    def __init__(self, context: EnvironmentContext, component_config: DemandSamplerConfig):
        super().__init__(context, component_config)
        self.episode_length = context.episode_length
        pre = context.preprocessed_data
        if pre is None:
            raise ValueError("EmpiricalDemandSampler requires preprocessed_data. "
                             "Ensure real_world data source is configured and preprocessing is enabled.")
        from ..data import pack_demand_frame
        use_val = context.data_mode == "val" and getattr(pre, "val_demand_data", None) is not None
        self.data = pre.val_demand_data if use_val else pre.demand_data
        cache = getattr(pre, "_packed", None)                      # E samplers of a batch share one packed frame
        if cache is None:
            cache = {}
            try:
                pre._packed = cache
            except AttributeError:
                pass
        rmap = getattr(context, "region_map", None)     # frame with raw region ids: mapped before the orders are grouped
        key = ("val" if use_val else "train", self.n_skus, None if rmap is None else tuple(int(x) for x in rmap))
        if key not in cache:
            cache[key] = pack_demand_frame(self.data, self.n_skus, rmap)
        self.frame = cache[key]
        self.available_timesteps = [int(t) for t in self.frame.timesteps]
        self.max_timestep = max(self.available_timesteps) if self.available_timesteps else 0
        if len(self.available_timesteps) < self.episode_length:
            raise ValueError(f"EmpiricalDemandSampler: episode_length ({self.episode_length}) > "
                             f"available timesteps ({len(self.available_timesteps)}). "
                             "Episode length must be <= number of available timesteps.")
        self._start_index: Optional[int] = None

    def start_index(self) -> int:
        """Index (into ``available_timesteps``) of this episode's window start; drawn on first use like the reference."""
        if self._start_index is None:
            max_start_idx = len(self.available_timesteps) - self.episode_length
            self._start_index = int(self._rng.integers(0, max_start_idx + 1))
        return self._start_index

    def sample(self, timestep: int) -> List[Order]:
        k = self.start_index() + timestep % self.episode_length
        region, qty = self.frame.orders(k)
        return [Order(region_id=int(r), sku_demands=q.astype(float)) for r, q in zip(region, qty)]

    def reset(self, rng: Optional[np.random.Generator] = None):
        super().reset(rng)
        self._start_index = None


class ReplayDemandSampler(BaseDemandSampler):
    """Extension: orders are supplied from outside as pre-sampled tensors."""

    def __init__(self, context: EnvironmentContext, component_config: DemandSamplerConfig):
        super().__init__(context, component_config)
        self.steps: List[List[Order]] = []

    def load(self, steps: List[List[Order]]):
        self.steps = steps

    def sample(self, timestep: int) -> List[Order]:
        return self.steps[timestep]
