"""Order wire format shared by the host samplers, the replay path and the kernel.

One step of demand for E environments is a CSR list (include/marlsc_b200.h ``marlsc_step_io``):
``offsets[E+1]`` int32, ``region[n]`` int16 (raw region ids), ``qty[n, S]`` uint8 or uint16, rows in the
order the reference sampler emits them (region-major for Poisson, reference demand_sampler.py:128-163).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

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


# ---------------------------------------------------------------------------------------------------------
# Sparse demand ("lines", include/marlsc_b200.h ``marlsc_step_io.lines``): the native input of the compact layout.
# ---------------------------------------------------------------------------------------------------------
LINE_LANES = 32


@dataclass
class LineBatch:
    offsets: np.ndarray   # [E+1] int32, in rounds
    lines: np.ndarray     # [n_rounds_pad, 32] uint16 storage: qty | region << 8 | (sku // 32) << 14, 0 = padding; the entries
                          # 2i, 2i+1 of lane l are the two halves of 32-bit word l of round pair i (see marlsc_b200.h)
    n_lines: int          # non-zero (order, SKU) cells

    @property
    def n_rounds(self) -> int:
        return int(self.offsets[-1])


def pack_lines(batch: OrderBatch, region_map: Optional[Sequence[int]] = None) -> LineBatch:
    """Regroup one step of orders into 32 line streams per environment (stream l = SKUs with ``s % 32 == l``, cells in
    the order the reference allocator meets them: order index, then SKU; demand_allocator.py:150-208). Region ids are
    mapped through ``region_map`` (raw -> included, reference preprocessor.py:382-441) when given."""
    if batch.qty_bytes != 1:
        raise ValueError("lines carry one-byte quantities")
    E = batch.offsets.shape[0] - 1
    n = batch.n_orders
    S = batch.qty.shape[1]
    if S > 4 * LINE_LANES:
        raise ValueError("lines address at most 128 SKUs")
    region = batch.region[:n].astype(np.int64)
    if region_map is not None:
        region = np.asarray(region_map, dtype=np.int64)[region]
    if n and (region.min() < 0 or region.max() >= 64):
        raise ValueError("lines address at most 64 regions")
    oj, s = np.nonzero(batch.qty[:n])                      # row-major: order index ascending, then SKU
    q = batch.qty[:n][oj, s].astype(np.int64)
    env = np.searchsorted(batch.offsets, oj, side="right") - 1
    stream = env * LINE_LANES + (s % LINE_LANES)
    order = np.argsort(stream, kind="stable")             # keeps (order, SKU) sequence inside a stream
    stream_sorted = stream[order]
    counts = np.bincount(stream, minlength=E * LINE_LANES)
    starts = np.cumsum(counts) - counts
    pos = np.arange(stream_sorted.shape[0]) - starts[stream_sorted]      # position inside the stream
    rounds = counts.reshape(E, LINE_LANES).max(axis=1) if E else np.zeros(0, np.int64)
    rounds = (rounds + 1) & ~1                            # whole round pairs: a lane's entries 2i, 2i+1 share a 32-bit word
    offsets = np.zeros(E + 1, dtype=np.int32)
    offsets[1:] = np.cumsum(rounds)
    total = int(offsets[-1])
    lines = np.zeros((max(total, 2), LINE_LANES), dtype=np.uint16)
    entry = q[order] | (region[oj[order]] << 8) | ((s[order] // LINE_LANES) << 14)
    # entry p of lane l of an environment starting at round r0 sits at uint16 index (r0 + (p & ~1)) * 32 + 2 l + (p & 1)
    flat = lines.reshape(-1)
    flat[(offsets[env[order]] + (pos & ~1)) * LINE_LANES + 2 * (s[order] % LINE_LANES) + (pos & 1)] = entry.astype(np.uint16)
    return LineBatch(offsets=offsets, lines=lines, n_lines=int(q.shape[0]))
