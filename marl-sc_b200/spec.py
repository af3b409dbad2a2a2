"""EnvironmentConfig + registry components -> ``marlsc_env_spec_t`` (include/marlsc_b200.h).

This is the host half of what ``InventoryEnvironment.__init__`` does in the reference
(src/environment/envs/multi_env.py:58-190): build the context, instantiate the five components from
the registry, derive home regions / closest warehouses / lead-time horizon, and freeze everything the
step needs into plain tables.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _capi
from .config.schema import EnvironmentConfig
from .context import EnvironmentContext, create_environment_context
from .registry import (get_demand_allocator, get_demand_sampler, get_lead_time_sampler,
                       get_lost_sales_handler, get_reward_calculator)

ROLLING_WINDOW = 5


def feature_mask(features) -> int:
    m = 0
    for name, bit in _capi.FEATURE_BITS.items():
        if getattr(features, name):
            m |= bit
    return m


def local_obs_dim(features, n_skus: int, max_lead: int, n_warehouses: int, include_warehouse_id: bool) -> int:
    """Width of one warehouse's observation (reference multi_env.py:444-502)."""
    f, S = features, n_skus
    d = S + int(f.inventory_aggregate) + max_lead * S + int(f.pipeline_aggregate)
    if f.incoming_demand_home:
        d += S + int(f.incoming_demand_home_aggregate)
    if f.units_shipped_home:
        d += S
    if f.units_shipped_away:
        d += S + int(f.units_shipped_away_aggregate)
    if f.stockout:
        d += S
    if f.rolling_demand_mean:
        d += S + int(f.rolling_demand_mean_aggregate)
    if f.demand_forecast:
        d += S + int(f.demand_forecast_aggregate)
    d += S * (int(f.days_of_supply) + int(f.net_inventory_position) + int(f.demand_variability))
    if f.demand_history:
        d += ROLLING_WINDOW * S
    return d + (n_warehouses if include_warehouse_id else 0)


@dataclass
class EnvSpec:
    """Python-side holder of the spec tables (keeps the NumPy buffers alive for ctypes)."""
    scalars: Dict[str, Any]
    tables: Dict[str, Optional[np.ndarray]]
    components: Dict[str, Any] = field(default_factory=dict)
    context: Optional[EnvironmentContext] = None

    def to_c(self) -> _capi.EnvSpecC:
        c = _capi.EnvSpecC()
        c.abi_version = _capi.ABI_VERSION
        for k, v in self.scalars.items():
            setattr(c, k, v)
        ctype = {"float64": C.c_double, "int32": C.c_int32, "float32": C.c_float}
        for k, arr in self.tables.items():
            if arr is None:
                continue
            setattr(c, k, arr.ctypes.data_as(C.POINTER(ctype[str(arr.dtype)])))
        return c


def build_env_spec(env_config: EnvironmentConfig, obs_normalization: str = "off",
                   obs_stats: Optional[Tuple[np.ndarray, np.ndarray]] = None, include_warehouse_id: bool = False,
                   region_map: Optional[Sequence[int]] = None, context: Optional[EnvironmentContext] = None,
                   seed_manager=None, data_mode: str = "train", preprocessed_data=None) -> EnvSpec:
    if obs_normalization not in _capi.NORM:
        raise ValueError(f"Unknown obs_normalization: {obs_normalization}. Available: {list(_capi.NORM)}")
    if obs_normalization == "meanstd":
        # like the reference's env, this mode leaves the observations raw (multi_env.py:700-702 only knows the fixed
        # statistics); the running filter lives outside the env - RLlib's MeanStdFilter connector there (ippo.py:173-176),
        # marlsc_b200.rollout.MeanStdFilter (RolloutCollector(obs_filter=...)) here. Say so instead of silently doing nothing.
        import warnings
        warnings.warn("obs_normalization='meanstd': the env emits raw observations; apply the running filter in the rollout "
                      "(marlsc_b200.rollout.MeanStdFilter via RolloutCollector(obs_filter=...))", stacklevel=2)
    empirical = env_config.components.demand_sampler.type == "empirical"
    ctx = context or create_environment_context(env_config, seed_manager=seed_manager, data_mode=data_mode,
                                                preprocessed_data=preprocessed_data,
                                                region_map=region_map if empirical else None)
    if empirical:
        # the empirical sampler maps its frame's raw region ids itself, before it sequences the orders of a step
        # (reference preprocessor.py:650); the kernels then see included region ids only
        region_map = None
    comps = dict(
        demand_sampler=get_demand_sampler(env_config, context=ctx),
        demand_allocator=get_demand_allocator(env_config, context=ctx),
        lead_time_sampler=get_lead_time_sampler(env_config, context=ctx),
        lost_sales_handler=get_lost_sales_handler(env_config, context=ctx),
        reward_calculator=get_reward_calculator(env_config, context=ctx))
    W, S, R = env_config.n_warehouses, env_config.n_skus, env_config.n_regions
    a = env_config.action_space
    amax = getattr(a.params, {"direct": "max_order_quantities", "demand_centered": "max_quantity_adjustment",
                              "base_stock": "max_stock_level"}[a.type])
    fields: Dict[str, Any] = {}
    for comp in comps.values():
        fields.update(comp.spec_fields())

    f64 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.float64))  # noqa: E731
    i32 = lambda x: np.ascontiguousarray(np.asarray(x, dtype=np.int32))    # noqa: E731
    norm = _capi.NORM[obs_normalization]
    mean = std = None
    L = int(fields["max_expected_lead"])
    if norm == 2:
        if obs_stats is None:       # the reference silently skips normalisation without stats (multi_env.py:700)
            norm = 0
        else:
            dim = local_obs_dim(env_config.features, S, L, W, False)
            mean = np.ascontiguousarray(np.asarray(obs_stats[0], dtype=np.float32).reshape(-1))
            std = np.ascontiguousarray(np.asarray(obs_stats[1], dtype=np.float32).reshape(-1))
            if mean.shape != (dim,) or std.shape != (dim,):
                raise ValueError(f"obs_stats must have shape ({dim},), got {mean.shape} / {std.shape}")
    rmap = None
    n_raw = R
    if region_map is not None:
        rmap = i32(region_map).reshape(-1)
        n_raw = int(rmap.shape[0])
        if rmap.min() < 0 or rmap.max() >= R:
            raise ValueError("region_map entries must be included region ids in [0, n_regions)")
    scalars = dict(
        n_warehouses=W, n_skus=S, n_regions=R, n_regions_raw=n_raw, episode_length=env_config.episode_length,
        max_expected_lead=L, ring_depth=int(fields["ring_depth"]), action_type=_capi.ACTION[a.type],
        lead_mode=int(fields["lead_mode"]), lost_sales_type=int(fields["lost_sales_type"]),
        reward_scope=int(fields["reward_scope"]), max_splits=int(fields["max_splits"]), obs_norm=norm,
        include_warehouse_id=int(bool(include_warehouse_id)), feature_mask=feature_mask(env_config.features),
        scale_factor=float(fields["scale_factor"]), lost_alpha=float(fields["lost_alpha"]))
    sc = ctx.shipment_cost
    tables = dict(
        action_max=f64(amax), out_fixed=f64(sc.outbound_fixed).reshape(W, R), out_var=f64(sc.outbound_variable).reshape(W, R),
        in_fixed=f64(sc.inbound_fixed).reshape(W, S), in_var=f64(sc.inbound_variable).reshape(W, S),
        hold_rate=f64(fields["hold_rate"]), pen_rate=f64(fields["pen_rate"]), sku_weights=f64(ctx.sku_weights),
        expected_lead=i32(fields["expected_lead"]).reshape(W, S), home_region=i32(np.argmin(ctx.distances, axis=1)),
        closest_wh=i32(fields["closest_wh"]), region_map=rmap, obs_mean=mean, obs_std=std)
    return EnvSpec(scalars=scalars, tables=tables, components=comps, context=ctx)
