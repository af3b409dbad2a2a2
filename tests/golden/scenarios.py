"""Scenario definitions for the golden trajectories (shared by make_golden.py and the tests).

Each scenario is an ``environment:`` dict in the reference's YAML layout plus the ``env_meta``-style
options the reference env takes (``obs_normalization``, ``include_warehouse_id``, ``obs_stats``;
reference: src/environment/envs/multi_env.py:152-168). Together they cover every registered component
(registry.py:300-308), every action space (multi_env.py:824-846), every feature block and
normalisation mode of the observation builder, R != W, a binding split limit and the large network.
"""
from __future__ import annotations

import copy
from typing import Any, Dict

import numpy as np

ALL_FEATURES_ON = dict(
    inventory=True, pipeline=True, incoming_demand_home=True, units_shipped_home=True,
    units_shipped_away=True, stockout=True, rolling_demand_mean=True, demand_forecast=True,
    days_of_supply=True, net_inventory_position=True, demand_variability=True, demand_history=True,
    inventory_aggregate=True, pipeline_aggregate=True, incoming_demand_home_aggregate=True,
    units_shipped_away_aggregate=True, rolling_demand_mean_aggregate=True,
    demand_forecast_aggregate=True)

DEFAULT_FEATURES = dict(  # == reference config_files/features/feature_config.yaml
    inventory=True, inventory_aggregate=True, pipeline=True, rolling_demand_mean=True,
    stockout=False, incoming_demand_home=False, units_shipped_home=False, units_shipped_away=False,
    demand_forecast=False, days_of_supply=False, net_inventory_position=False, demand_history=False,
    demand_variability=False, pipeline_aggregate=False, incoming_demand_home_aggregate=False,
    units_shipped_away_aggregate=False, rolling_demand_mean_aggregate=False,
    demand_forecast_aggregate=False)


def small_default() -> Dict[str, Any]:
    """The reference's shipped 3 warehouse x 2 SKU x 3 region config
    (config_files/environments/env_symmetric_3WH2SKU.yaml), written out so it travels."""
    return dict(
        action_space=dict(type="direct", params=dict(max_order_quantities=[40, 40])),
        n_warehouses=3, n_skus=2, n_regions=3, episode_length=100,
        max_wh_capacities=[10000000, 10000000, 10000000],
        initial_inventory=dict(type="custom", params=dict(values=[[60, 60], [60, 60], [60, 60]])),
        cost_structure=dict(
            holding_cost=1.0, penalty_cost=[5, 5],
            shipment_cost=dict(
                outbound_fixed=[[0, 0, 0], [0, 0, 0], [0, 0, 0]],
                outbound_variable=[[0.05, 0.5, 0.5], [0.5, 0.05, 0.5], [0.5, 0.5, 0.05]],
                inbound_fixed=[[0, 0], [0, 0], [0, 0]],
                inbound_variable=[[1.0, 1.0], [1.0, 1.0], [1.0, 1.0]]),
            sku_weights=[1.0, 1.0],
            distances=[[50, 500, 500], [500, 50, 500], [500, 500, 50]]),
        components=dict(
            demand_sampler=dict(type="poisson", params=dict(
                lambda_orders=[4, 4, 4], probability_skus=[0.667, 0.667, 0.667],
                lambda_quantity=[[5, 5], [5, 5], [5, 5]])),
            demand_allocator=dict(type="greedy", params=dict(max_splits="default")),
            lead_time_sampler=dict(type="fixed", params=dict(expected_lead_times=[[3, 3], [3, 3], [3, 3]])),
            lost_sales_handler=dict(type="shipment", params=None),
            reward_calculator=dict(type="cost", params=dict(
                scope="agent", scale_factor=0.01, cost_weights=[0.25, 0.25, 0.25, 0.25]))),
        data_source=dict(type="custom"),
        features=dict(DEFAULT_FEATURES))


def allfeat_ratio_stochastic() -> Dict[str, Any]:
    """3 WH x 5 SKU: every feature block, ratio normalisation, warehouse id, stochastic per-SKU lead
    deviation, closest handler, team reward, demand_centered actions, uniform start inventory,
    weight-dependent warehouse priority (non-zero outbound_fixed) and a binding split limit."""
    W, S, R = 3, 5, 3
    return dict(
        action_space=dict(type="demand_centered", params=dict(max_quantity_adjustment=[6, 8, 5, 7, 9])),
        n_warehouses=W, n_skus=S, n_regions=R, episode_length=40,
        max_wh_capacities=[1e7] * W,
        initial_inventory=dict(type="uniform", params=dict(min=5, max=40)),
        cost_structure=dict(
            holding_cost=[0.5, 1.0, 0.25, 2.0, 1.5], penalty_cost=4.0,
            shipment_cost=dict(
                outbound_fixed=[[1.0, 4.0, 6.0], [5.0, 0.5, 3.0], [2.0, 6.0, 1.5]],
                outbound_variable=[[0.125, 0.25, 0.5], [0.0625, 0.375, 0.75], [0.5, 0.03125, 0.25]],
                inbound_fixed=[[1, 2, 0, 3, 1], [0, 1, 2, 1, 0], [2, 2, 1, 0, 1]],
                inbound_variable=[[1.0, 0.5, 0.75, 1.25, 1.0], [0.5, 1.0, 1.0, 0.25, 2.0], [1.5, 1.0, 0.5, 1.0, 0.75]]),
            sku_weights=[0.5, 1.0, 2.0, 1.5, 0.25],
            distances=[[40, 300, 450], [320, 60, 280], [500, 310, 35]]),
        components=dict(
            demand_sampler=dict(type="poisson", params=dict(
                lambda_orders=[3, 2, 4], probability_skus=[0.5, 0.6, 0.4],
                lambda_quantity=[[4, 3, 2, 5, 6], [2, 5, 3, 4, 1], [6, 2, 4, 3, 5]])),
            demand_allocator=dict(type="greedy", params=dict(max_splits=1)),
            lead_time_sampler=dict(type="stochastic", params=dict(
                expected_lead_times=[[2, 3, 1, 4, 2], [1, 2, 3, 2, 4], [3, 1, 2, 3, 1]],
                deviation=dict(type="uniform", max_deviation=[0, 1, 2, 1, 0]))),
            lost_sales_handler=dict(type="closest", params=None),
            reward_calculator=dict(type="cost", params=dict(
                scope="team", scale_factor=0.05, cost_weights=[0.25, 0.25, 0.25, 0.25]))),
        data_source=dict(type="custom"),
        features=dict(ALL_FEATURES_ON))


def basestock_cost_meanstd() -> Dict[str, Any]:
    """3 WH x 2 SKU: base_stock actions, cost (softmax) lost-sales handler, no order splitting,
    empty start inventory, scalar stochastic lead deviation, meanstd_custom normalisation."""
    d = small_default()
    d["action_space"] = dict(type="base_stock", params=dict(max_stock_level=[90, 70]))
    d["episode_length"] = 30
    d["initial_inventory"] = dict(type="zero", params=None)
    d["cost_structure"]["holding_cost"] = 0.75
    d["cost_structure"]["penalty_cost"] = [3.0, 6.5]
    d["cost_structure"]["sku_weights"] = [1.0, 2.0]
    d["cost_structure"]["shipment_cost"]["outbound_fixed"] = [[0.5, 2, 2], [2, 0.5, 2], [2, 2, 0.5]]
    d["components"]["demand_allocator"]["params"]["max_splits"] = 0
    d["components"]["lead_time_sampler"] = dict(type="stochastic", params=dict(
        expected_lead_times=[[2, 3], [3, 2], [1, 4]], deviation=dict(type="uniform", max_deviation=1)))
    d["components"]["lost_sales_handler"] = dict(type="cost", params=dict(alpha=2.0))
    feats = dict(DEFAULT_FEATURES)
    feats.update(incoming_demand_home=True, demand_forecast=True, demand_forecast_aggregate=True,
                 units_shipped_away=True, units_shipped_away_aggregate=True, stockout=True)
    d["features"] = feats
    return d


def regions_ne_warehouses() -> Dict[str, Any]:
    """2 WH x 3 SKU x 5 regions (the env code is shape-agnostic; schema.py:670-675 is bypassed)."""
    W, S, R = 2, 3, 5
    return dict(
        action_space=dict(type="direct", params=dict(max_order_quantities=[30, 25, 20])),
        n_warehouses=W, n_skus=S, n_regions=R, episode_length=25,
        max_wh_capacities=[1e7] * W,
        initial_inventory=dict(type="custom", params=dict(values=20)),
        cost_structure=dict(
            holding_cost=1.0, penalty_cost=[5, 4, 6],
            shipment_cost=dict(
                outbound_fixed=[[0] * R] * W,
                outbound_variable=[[0.05, 0.3, 0.2, 0.45, 0.15], [0.4, 0.1, 0.25, 0.05, 0.35]],
                inbound_fixed=[[0.5] * S] * W,
                inbound_variable=[[1.0] * S] * W),
            sku_weights=[1.0, 1.0, 1.0],
            distances=[[60, 300, 210, 480, 140], [410, 90, 260, 55, 330]]),
        components=dict(
            demand_sampler=dict(type="poisson", params=dict(
                lambda_orders=[2, 2, 1, 3, 2], probability_skus=[0.6] * R,
                lambda_quantity=[[4, 4, 4]] * R)),
            demand_allocator=dict(type="greedy", params=dict(max_splits="default")),
            lead_time_sampler=dict(type="fixed", params=dict(expected_lead_times=[[2, 1, 3], [3, 2, 1]])),
            lost_sales_handler=dict(type="shipment", params=None),
            reward_calculator=dict(type="cost", params=dict(
                scope="agent", scale_factor=0.01, cost_weights=[0.25, 0.25, 0.25, 0.25]))),
        data_source=dict(type="custom"),
        features=dict(DEFAULT_FEATURES))


def large_network(lambda_orders: float = 1.0, probability_skus: float = 0.2,
                  lambda_quantity: float = 5.0, episode_length: int = 100) -> Dict[str, Any]:
    """BASELINE config 3: 10 WH x 100 SKU x 50 regions, lead times 1..10,
    tie-free outbound cost columns (static warehouse priority), shipment lost-sales handler."""
    W, S, R = 10, 100, 50
    rng = np.random.default_rng(1)
    lead = (1 + rng.integers(0, 10, size=(W, S))).tolist()
    dist = np.round(50 + 450 * rng.random((W, R)), 3).tolist()
    base = 0.05 + 0.045 * np.arange(W)
    out_var = np.stack([rng.permutation(base) for _ in range(R)], axis=1)  # [W, R], each column a permutation
    return dict(
        action_space=dict(type="direct", params=dict(max_order_quantities=[40] * S)),
        n_warehouses=W, n_skus=S, n_regions=R, episode_length=episode_length,
        max_wh_capacities=[1e7] * W,
        initial_inventory=dict(type="custom", params=dict(values=60)),
        cost_structure=dict(
            holding_cost=1.0, penalty_cost=5.0,
            shipment_cost=dict(
                outbound_fixed=[[0.0] * R] * W,
                outbound_variable=np.round(out_var, 6).tolist(),
                inbound_fixed=[[0.0] * S] * W,
                inbound_variable=[[1.0] * S] * W),
            sku_weights=[1.0] * S,
            distances=dist),
        components=dict(
            demand_sampler=dict(type="poisson", params=dict(
                lambda_orders=[lambda_orders] * R, probability_skus=[probability_skus] * R,
                lambda_quantity=[[lambda_quantity] * S] * R)),
            demand_allocator=dict(type="greedy", params=dict(max_splits="default")),
            lead_time_sampler=dict(type="fixed", params=dict(expected_lead_times=lead)),
            lost_sales_handler=dict(type="shipment", params=None),
            reward_calculator=dict(type="cost", params=dict(
                scope="agent", scale_factor=0.01, cost_weights=[0.25, 0.25, 0.25, 0.25]))),
        data_source=dict(type="custom"),
        features=dict(DEFAULT_FEATURES))


def empirical_regionmap() -> Dict[str, Any]:
    """Empirical demand (reference components/demand_sampler.py:166-271) with the excluded-region mapping
    (src/data/preprocessor.py:382-441): the large network's shape with 10 included regions; the demand frame carries 50
    raw regions, 40 of which are mapped onto the included ones (tests/golden/make_golden.py builds the frame)."""
    d = large_network(episode_length=24)
    W, S, R = 10, 100, 10
    rng = np.random.default_rng(3)
    base = 0.05 + 0.045 * np.arange(W)
    out_var = np.stack([rng.permutation(base) for _ in range(R)], axis=1)
    d["n_regions"] = R
    d["cost_structure"]["shipment_cost"]["outbound_fixed"] = [[0.0] * R] * W
    d["cost_structure"]["shipment_cost"]["outbound_variable"] = np.round(out_var, 6).tolist()
    d["cost_structure"]["distances"] = np.round(50 + 450 * rng.random((W, R)), 3).tolist()
    d["components"]["demand_sampler"] = dict(type="empirical", params=None)
    d["initial_inventory"] = dict(type="custom", params=dict(values=25))
    return d


def empirical_frame(n_raw_regions: int = 50, n_included: int = 10, n_skus: int = 100, n_days: int = 40, seed: int = 11):
    """Synthetic stand-in for the reference's raw order data: (frame with RAW region ids, warehouse-to-region cost table,
    all region ids, selected region ids). Region ids are strings like the reference's CSVs; a few excluded regions have no
    warehouse pair at all and a few share no warehouse with an included region (both fallback rules)."""
    import pandas as pd
    rng = np.random.default_rng(seed)
    all_ids = [f"REG_{i:02d}" for i in range(n_raw_regions)]
    selected = [all_ids[i] for i in sorted(rng.choice(n_raw_regions, n_included, replace=False))]
    wh = [f"WH_{i}" for i in range(10)]
    rows = []
    for i, rid in enumerate(all_ids):
        if rid not in selected and i % 11 == 3:
            continue                                         # no warehouse pair: first included region
        if rid not in selected and i % 13 == 5:
            rows.append(dict(sourcenodeid="WH_X", destinationregionid=rid, fixed_costs=1.0))   # no shared warehouse
            continue
        for wname in rng.choice(wh, int(rng.integers(2, 6)), replace=False):
            rows.append(dict(sourcenodeid=wname, destinationregionid=rid, fixed_costs=float(np.round(rng.uniform(1, 20), 3))))
    wtr = pd.DataFrame(rows)
    recs = []
    oid = 0
    for day in range(3, 3 + n_days):                         # timesteps need not start at 0
        for _ in range(int(rng.integers(30, 60))):
            rid = all_ids[int(rng.integers(0, n_raw_regions))]
            oid += 1
            skus = rng.choice(n_skus + 3, int(rng.integers(1, 25)))          # repeats (summed) and ids >= n_skus (dropped)
            for sk in skus:
                recs.append(dict(timestep=day, region_raw=rid, order_id=f"SO{oid:06d}", sku_id=int(sk),
                                 quantity=float(rng.integers(1, 9))))
    return pd.DataFrame(recs), wtr, all_ids, selected


def _obs_stats(dim: int, seed: int):
    rng = np.random.default_rng(seed)
    return (rng.normal(5.0, 3.0, dim).astype(np.float32), rng.uniform(0.5, 4.0, dim).astype(np.float32))


# name -> (env dict builder, options)
SCENARIOS = {
    "small_default": dict(env=small_default, n_envs=16, steps=100, base_seed=1234, action_seed=0),
    "allfeat_ratio_stochastic": dict(env=allfeat_ratio_stochastic, n_envs=8, steps=40, base_seed=77,
                                     action_seed=1, obs_normalization="ratio", include_warehouse_id=True),
    "basestock_cost_meanstd": dict(env=basestock_cost_meanstd, n_envs=8, steps=30, base_seed=5,
                                   action_seed=2, obs_normalization="meanstd_custom", obs_stats_seed=3),
    "regions_ne_warehouses": dict(env=regions_ne_warehouses, n_envs=6, steps=25, base_seed=99,
                                  action_seed=3, allow_region_mismatch=True),
    "large_network": dict(env=large_network, n_envs=2, steps=12, base_seed=2024, action_seed=4,
                          allow_region_mismatch=True),
    # the flagship shape in steady state: ring depth 10 wraps three times; "lite" files keep the state, ordered
    # quantities, per-region results, costs, rewards and observations but not the [W,R,S] shipment cube
    "large_network_long": dict(env=large_network, n_envs=8, steps=32, base_seed=2025, action_seed=5,
                               allow_region_mismatch=True, lite=True),
    # ... and driven by the base-stock heuristic the benchmark replays (stock settles near 5 units per cell)
    "large_network_basestock": dict(env=large_network, n_envs=8, steps=36, base_seed=2026, action_seed=6,
                                    allow_region_mismatch=True, lite=True, policy="base_stock"),
    "empirical_regionmap": dict(env=empirical_regionmap, n_envs=4, steps=24, base_seed=31, action_seed=7,
                                allow_region_mismatch=True, lite=True, empirical=True),
}


def scenario(name: str) -> Dict[str, Any]:
    sc = copy.deepcopy({k: v for k, v in SCENARIOS[name].items() if k != "env"})
    sc["env"] = SCENARIOS[name]["env"]()
    sc.setdefault("obs_normalization", "off")
    sc.setdefault("include_warehouse_id", False)
    sc.setdefault("allow_region_mismatch", False)
    return sc


def obs_stats_for(sc: Dict[str, Any], dim: int):
    return _obs_stats(dim, sc["obs_stats_seed"]) if "obs_stats_seed" in sc else None
