"""Empirical demand on the hot path: the preprocessed demand frame, the excluded-region mapping and their device form.

What the reference does with its (unshipped) raw data set, restated for the two pieces the env step consumes:

* ``map_excluded_regions`` / ``build_region_map`` - reference ``DataProcessor.map_excluded_regions``
  (src/data/preprocessor.py:382-441): an order whose region is not among the selected regions is reassigned to the
  selected region that shares warehouses with it and has the smallest mean ``fixed_costs`` over those warehouses;
  without shared warehouses (or without any warehouse pair) to the first selected region. The result is a static table
  ``raw region -> included region index`` - ``marlsc_env_spec_t.region_map`` - that the kernels (wide layout) or the
  line packers (compact layout) apply to every order's region id.
* ``PreprocessedData`` / ``pack_demand_frame`` / ``EmpiricalDemandSampler`` data - reference ``PreprocessedData``
  (preprocessor.py:21-32), the frame ``(timestep, region_id, order_id, sku_id, quantity)`` built at
  preprocessor.py:682-694, and the replay of a random contiguous window of it
  (src/environment/components/demand_sampler.py:166-271). ``pack_demand_frame`` turns the frame into per-timestep CSR
  order tensors with the reference's ``(region_id, order_id)`` grouping order, once; the sampler then only slices.

The raw CSVs are not part of the reference repository, so a frame has to be supplied by the caller
(``env_meta={"preprocessed_data": PreprocessedData(frame)}``); everything downstream of the frame is the reference's.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, List, Optional, Sequence

import numpy as np


@dataclass
class PreprocessedData:
    """Reference ``PreprocessedData`` (preprocessor.py:21-32): pandas frames with columns
    ``timestep, region_id, order_id, sku_id, quantity``."""
    demand_data: Any
    val_demand_data: Optional[Any] = None


def nearest_included_region(excluded_region, warehouse_to_region_df, selected_region_ids: Sequence):
    """The included region one excluded region is mapped to (preprocessor.py:407-436)."""
    selected_str = [str(r) for r in selected_region_ids]
    wtr = warehouse_to_region_df
    dest = wtr["destinationregionid"].astype(str)
    excluded_pairs = wtr[dest == str(excluded_region)]
    if len(excluded_pairs) == 0:
        return selected_region_ids[0]
    warehouse_ids = excluded_pairs["sourcenodeid"].unique()
    included_pairs = wtr[dest.isin(selected_str) & wtr["sourcenodeid"].isin(warehouse_ids)]
    if len(included_pairs) == 0:
        return selected_region_ids[0]
    region_costs = included_pairs.groupby("destinationregionid")["fixed_costs"].mean()
    nearest = str(region_costs.idxmin())
    return next((r for r in selected_region_ids if str(r) == nearest), selected_region_ids[0])


def map_excluded_regions(order_region_ids, warehouse_to_region_df, selected_region_ids: Sequence):
    """Reference ``DataProcessor.map_excluded_regions`` (preprocessor.py:382-441) on a pandas Series of region ids."""
    mapped = order_region_ids.copy()
    selected = set(str(r) for r in selected_region_ids)
    as_str = order_region_ids.astype(str)
    for excluded in order_region_ids[~as_str.isin(selected)].unique():
        mapped[as_str == str(excluded)] = nearest_included_region(excluded, warehouse_to_region_df, selected_region_ids)
    return mapped


def build_region_map(all_region_ids: Sequence, warehouse_to_region_df, selected_region_ids: Sequence) -> List[int]:
    """``region_map[i]`` = index (into ``selected_region_ids``) of the region that orders of raw region
    ``all_region_ids[i]`` are allocated to: itself when selected, else by the rule of ``map_excluded_regions``.
    This is the table ``BatchedInventoryEnv(region_map=...)`` / ``marlsc_env_spec_t.region_map`` take."""
    index = {str(r): i for i, r in enumerate(selected_region_ids)}
    out = []
    for r in all_region_ids:
        if str(r) in index:
            out.append(index[str(r)])
        else:
            out.append(index[str(nearest_included_region(r, warehouse_to_region_df, selected_region_ids))])
    return out


@dataclass
class DemandFrame:
    """A demand frame packed per timestep: orders of timestep ``timesteps[k]`` are rows
    ``[step_offsets[k], step_offsets[k+1])`` of ``order_region`` / ``order_qty``, in the reference sampler's order
    (``groupby(['region_id', 'order_id'])`` sorts by region, then order id; demand_sampler.py:247)."""
    timesteps: np.ndarray      # [n_t] sorted unique timesteps of the frame
    step_offsets: np.ndarray   # [n_t + 1] int64
    order_region: np.ndarray   # [n] int16 region ids as they stand in the frame
    order_qty: np.ndarray      # [n, S] uint8 | uint16 summed quantities per SKU

    @property
    def n_timesteps(self) -> int:
        return int(self.timesteps.shape[0])

    def orders(self, k: int):
        a, b = int(self.step_offsets[k]), int(self.step_offsets[k + 1])
        return self.order_region[a:b], self.order_qty[a:b]


def pack_demand_frame(df, n_skus: int, region_map: Optional[Sequence[int]] = None) -> DemandFrame:
    """Group the frame's rows into orders once: ``(timestep, region_id, order_id)`` groups in sorted order, quantities
    of the same SKU summed, SKU ids outside ``[0, n_skus)`` dropped (demand_sampler.py:247-258). With ``region_map`` the
    frame carries RAW region ids and is mapped first, as the reference's preprocessing does before it stores the frame
    (preprocessor.py:650) - the orders of a step are then sequenced by their MAPPED region, which is the order the
    greedy allocation sees them in."""
    import pandas as pd
    if len(df) == 0:
        return DemandFrame(np.zeros(0, np.int64), np.zeros(1, np.int64), np.zeros(0, np.int16), np.zeros((0, n_skus), np.uint8))
    if region_map is not None:
        df = df.assign(region_id=np.asarray(region_map, dtype=np.int64)[df["region_id"].to_numpy(dtype=np.int64)])
    keys = df[["timestep", "region_id", "order_id"]]
    # group ids in the order pandas' groupby (sort=True) yields them: timestep, then region_id, then order_id
    order_idx, uniques = pd.factorize(pd.MultiIndex.from_frame(keys), sort=True)
    n = len(uniques)
    ts = np.asarray(uniques.get_level_values(0), dtype=np.int64)
    region = np.asarray(uniques.get_level_values(1), dtype=np.int64)
    sku = df["sku_id"].to_numpy()
    qty = df["quantity"].to_numpy(dtype=np.float64)
    ok = (sku >= 0) & (sku < n_skus)
    dense = np.zeros((n, n_skus), dtype=np.float64)
    np.add.at(dense, (order_idx[ok], sku[ok].astype(np.int64)), qty[ok])
    if not np.array_equal(dense, np.round(dense)) or dense.min(initial=0) < 0:
        raise ValueError("order quantities must be non-negative whole numbers")
    if region.min(initial=0) < 0 or region.max(initial=0) > 32767:
        raise ValueError("region ids must be in [0, 32767]")
    dtype = np.uint8 if dense.max(initial=0) <= 255 else np.uint16
    if dense.max(initial=0) > 65535:
        raise ValueError("order quantities above 65535 are not supported")
    timesteps, first = np.unique(ts, return_index=True)
    offsets = np.concatenate([first, [n]]).astype(np.int64)
    return DemandFrame(timesteps=timesteps, step_offsets=offsets, order_region=region.astype(np.int16), order_qty=dense.astype(dtype))


class DeviceDemandFrame:
    """A packed demand frame resident on the device, sliced per step for a whole batch of environments: environment e
    replays the window starting at timestep index ``start[e]`` (reference demand_sampler.py:226-238), so step t of the
    batch is the concatenation of the frames' timestep slices ``start[e] + t % episode_length``."""

    def __init__(self, frame: DemandFrame, episode_length: int, device):
        import torch
        if frame.n_timesteps < episode_length:
            raise ValueError(f"EmpiricalDemandSampler: episode_length ({episode_length}) > available timesteps "
                             f"({frame.n_timesteps}). Episode length must be <= number of available timesteps.")
        self.frame, self.episode_length, self.device = frame, int(episode_length), device
        self.step_offsets = torch.from_numpy(frame.step_offsets).to(device)
        self.region = torch.from_numpy(frame.order_region).to(device)
        self.qty = torch.from_numpy(frame.order_qty.view(np.uint8)).to(device)
        self.qty_bytes = frame.order_qty.dtype.itemsize
        self.n_skus = frame.order_qty.shape[1]

    def max_start(self) -> int:
        return self.frame.n_timesteps - self.episode_length

    def step_orders(self, start, t: int):
        """Orders of step ``t`` for environments whose windows start at ``start`` (int64 [E] on the device) as the CSR
        tensors ``DeviceOrders`` wraps: (offsets int32 [E+1], region int16 [n], qty bytes [n*S*qty_bytes + pad], n)."""
        import torch
        k = start + (t % self.episode_length)
        a, b = self.step_offsets[k], self.step_offsets[k + 1]
        counts = b - a
        E = start.shape[0]
        offsets = torch.zeros(E + 1, dtype=torch.int32, device=self.device)
        offsets[1:] = counts.cumsum(0).to(torch.int32)
        n = int(offsets[-1].item())
        env = torch.repeat_interleave(torch.arange(E, device=self.device), counts)
        rows = a[env] + (torch.arange(n, device=self.device) - offsets[:-1].to(torch.int64)[env])
        region = self.region[rows] if n else torch.zeros(1, dtype=torch.int16, device=self.device)
        row_bytes = self.n_skus * self.qty_bytes
        qty = self.qty.view(-1, row_bytes)[rows].reshape(-1)
        pad = (-(n * row_bytes)) % 16 + (16 if n == 0 else 0)
        if pad:
            qty = torch.cat([qty, torch.zeros(pad, dtype=torch.uint8, device=self.device)])
        return offsets, region, qty, n
