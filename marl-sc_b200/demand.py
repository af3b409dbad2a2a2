"""Order wire format shared by the host samplers, the replay path and the kernel.

One step of demand for E environments is a CSR list (include/marlsc_b200.h ``marlsc_step_io``):
``offsets[E+1]`` int32, ``region[n]`` int16 (raw region ids), ``qty[n, S]`` uint8 or uint16, rows in the
order the reference sampler emits them (region-major for Poisson, reference demand_sampler.py:128-163).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np


@dataclass
class OrderBatch:
    offsets: np.ndarray   # [E+1] int32
    region: np.ndarray    # [n] int16
    qty: np.ndarray       # [n_pad, S] uint8 | uint16 (rows >= n are padding so the buffer is 16-byte sized)

    @property
    def n_orders(self) -> int:
        return int(self.offsets[-1])

    @property
    def qty_bytes(self) -> int:
        return self.qty.dtype.itemsize


def _pad_rows(n: int, S: int, itemsize: int) -> int:
    """Rows to allocate so that the byte size is a multiple of 16 and at least 16."""
    rows = max(n, 1)
    while (rows * S * itemsize) % 16:
        rows += 1
    return rows


def pack_orders(per_env: Sequence[Sequence], n_skus: int, qty_dtype=None) -> OrderBatch:
    """``per_env[e]`` is that environment's list of orders for this step; an order is anything with
    ``region_id`` / ``sku_demands`` or a ``(region, quantities)`` pair."""
    regions: List[int] = []
    rows: List[np.ndarray] = []
    offsets = np.zeros(len(per_env) + 1, dtype=np.int32)
    for e, orders in enumerate(per_env):
        for o in orders:
            r, q = (o.region_id, o.sku_demands) if hasattr(o, "region_id") else o
            regions.append(int(r))
            rows.append(np.asarray(q))
        offsets[e + 1] = len(regions)
    n = len(regions)
    q = np.stack(rows).astype(np.int64) if n else np.zeros((0, n_skus), dtype=np.int64)
    if q.min(initial=0) < 0:
        raise ValueError("order quantities must be non-negative")
    if qty_dtype is None:
        qty_dtype = np.uint8 if q.max(initial=0) <= 255 else np.uint16
    if q.max(initial=0) > np.iinfo(qty_dtype).max:
        raise ValueError(f"order quantity {q.max()} does not fit {np.dtype(qty_dtype).name}")
    qty = np.zeros((_pad_rows(n, n_skus, np.dtype(qty_dtype).itemsize), n_skus), dtype=qty_dtype)
    qty[:n] = q
    return OrderBatch(offsets=offsets, region=np.asarray(regions, dtype=np.int16), qty=qty)
