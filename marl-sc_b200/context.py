"""Shared environment context handed to every component
(reference: src/environment/context.py:30-65, 118-209)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple, Union

import numpy as np

from .config.schema import CostStructureConfig, EnvironmentConfig


@dataclass
class ShipmentCosts:
    outbound_fixed: np.ndarray      # [W, R]
    outbound_variable: np.ndarray   # [W, R]
    inbound_fixed: np.ndarray       # [W, S]
    inbound_variable: np.ndarray    # [W, S]


@dataclass
class EnvironmentContext:
    n_warehouses: int
    n_skus: int
    n_regions: int
    episode_length: int
    holding_cost: Union[float, np.ndarray]
    penalty_cost: Union[float, np.ndarray]
    shipment_cost: ShipmentCosts
    sku_weights: np.ndarray
    distances: np.ndarray
    preprocessed_data: Optional[object] = None
    data_mode: str = "train"
    region_map: Optional[object] = None   # raw -> included region ids of the demand frame (marlsc_b200.data.build_region_map)


def convert_cost_structure(cs: CostStructureConfig) -> Tuple[Union[float, np.ndarray], Union[float, np.ndarray]]:
    """Lists stay per-SKU rates, scalars stay scalars (they are multiplied by the SKU weight later,
    reference reward_calculator.py:128-137)."""
    def conv(v):
        return np.array(v, dtype=float) if isinstance(v, list) else float(v)
    return conv(cs.holding_cost), conv(cs.penalty_cost)


def create_environment_context(env_config: EnvironmentConfig, seed_manager=None, data_mode: str = "train",
                               preprocessed_data=None, region_map=None) -> EnvironmentContext:
    """``preprocessed_data`` (``marlsc_b200.data.PreprocessedData``) carries the demand frame of the ``empirical`` demand
    sampler; the reference builds it from raw CSVs that are not part of its repository (context.py:66-113)."""
    if env_config.data_source.type != "custom":
        raise ValueError(
            f"data_source.type='{env_config.data_source.type}' needs the reference's raw data files, which this "
            "hot-path implementation does not ship; provide a 'custom' data source")
    cs = env_config.cost_structure
    hold, pen = convert_cost_structure(cs)
    sc = cs.shipment_cost
    return EnvironmentContext(
        n_warehouses=env_config.n_warehouses, n_skus=env_config.n_skus, n_regions=env_config.n_regions,
        episode_length=env_config.episode_length, holding_cost=hold, penalty_cost=pen,
        shipment_cost=ShipmentCosts(
            outbound_fixed=np.array(sc.outbound_fixed, dtype=float), outbound_variable=np.array(sc.outbound_variable, dtype=float),
            inbound_fixed=np.array(sc.inbound_fixed, dtype=float), inbound_variable=np.array(sc.inbound_variable, dtype=float)),
        sku_weights=np.array(cs.sku_weights, dtype=float), distances=np.array(cs.distances, dtype=float),
        preprocessed_data=preprocessed_data, data_mode=data_mode, region_map=region_map)
