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
    lines: np.ndarray     # [n_rounds_pad, 32] uint16 storage: qty | region << 8 | slot << 14, 0 = padding; the entries
                          # 2i, 2i+1 of lane l are the two halves of 32-bit word l of round pair i; entries 0, 1 of a lane are
                          # its SKU map (byte k = SKU of slot k, 255 = none) (see marlsc_b200.h)
    n_lines: int          # non-zero (order, SKU) cells

    @property
    def n_rounds(self) -> int:
        return int(self.offsets[-1])


def pack_lines(batch: OrderBatch, region_map: Optional[Sequence[int]] = None, balance: bool = True) -> LineBatch:
    """Regroup one step of orders into 32 line streams per environment. A stream carries the lines of up to four SKUs,
    named by the stream's first two entries (its SKU map: byte k = the SKU of slot k, 255 = none), SKU after SKU, each
    SKU's lines in order sequence (the greedy allocation of one SKU never looks at another SKU, demand_allocator.py:150-208,
    so only the sequence inside a SKU matters). With ``balance`` the SKUs of an environment are ranked by line count
    (descending, ties by SKU id) and dealt to the streams in a snake, so the streams end within a few entries of each
    other - the allocation kernel runs as long as the longest, and a step's block is as many rounds long; without it
    SKU s goes to stream s % 32, slot s // 32. ``marlsc_lines_from_orders`` builds the same bytes on the device.
    Region ids are mapped through ``region_map`` (raw -> included, reference preprocessor.py:382-441) when given."""
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
    per_sku = np.bincount(env * S + s, minlength=E * S).reshape(E, S)        # lines of every (environment, SKU)
    sku_ids = np.arange(S)
    if balance:
        order_by_load = np.argsort(-per_sku * 256 + sku_ids[None, :], axis=1, kind="stable")
        rank = np.empty_like(order_by_load)
        np.put_along_axis(rank, order_by_load, np.broadcast_to(sku_ids, (E, S)), axis=1)
        slot, col = rank // LINE_LANES, rank % LINE_LANES
        lane = np.where(slot % 2 == 1, LINE_LANES - 1 - col, col)
    else:
        lane, slot = np.broadcast_to(sku_ids % LINE_LANES, (E, S)), np.broadcast_to(sku_ids // LINE_LANES, (E, S))
    # entries of a stream: two map entries, then slot 0's lines, slot 1's, ...
    by_dest = np.zeros((E, LINE_LANES, 4), dtype=np.int64)
    sku_map = np.full((E, LINE_LANES, 4), 255, dtype=np.int64)
    ee = np.broadcast_to(np.arange(E)[:, None], (E, S))
    by_dest[ee, lane, slot] = per_sku
    sku_map[ee, lane, slot] = np.broadcast_to(sku_ids, (E, S))
    first = np.cumsum(by_dest, axis=2) - by_dest           # first entry of a slot's lines inside its stream (after the map)
    base = first[ee, lane, slot]                           # [E, S]
    stream_len = by_dest.sum(axis=2)                       # [E, 32]
    longest = stream_len.max(axis=1) if E else np.zeros(0, np.int64)
    rounds = np.where(longest > 0, (longest + 2 + 1) & ~1, 0)       # whole round pairs: a lane's entries 2i, 2i+1 share a 32-bit word
    offsets = np.zeros(E + 1, dtype=np.int32)
    offsets[1:] = np.cumsum(rounds)
    total = int(offsets[-1])
    lines = np.zeros((max(total, 2), LINE_LANES), dtype=np.uint16)
    flat = lines.reshape(-1)
    # position of a line inside its SKU's run: lines of one (environment, SKU) in order sequence
    key = env * S + s
    order = np.argsort(key, kind="stable")
    key_sorted = key[order]
    starts = np.cumsum(per_sku.reshape(-1)) - per_sku.reshape(-1)
    within = np.arange(key_sorted.shape[0]) - starts[key_sorted]
    e_o, s_o = env[order], s[order]
    pos = 2 + base[e_o, s_o] + within
    entry = q[order] | (region[oj[order]] << 8) | (slot[e_o, s_o] << 14)
    # entry p of lane l of an environment starting at round r0 sits at uint16 index (r0 + (p & ~1)) * 32 + 2 l + (p & 1)
    flat[(offsets[e_o] + (pos & ~1)) * LINE_LANES + 2 * lane[e_o, s_o] + (pos & 1)] = entry.astype(np.uint16)
    live = np.nonzero(longest > 0)[0]
    if live.size:
        words = (sku_map[live, :, 0] | (sku_map[live, :, 1] << 8)), (sku_map[live, :, 2] | (sku_map[live, :, 3] << 8))
        idx = offsets[live].astype(np.int64)[:, None] * LINE_LANES + 2 * np.arange(LINE_LANES)[None, :]
        flat[idx] = words[0].astype(np.uint16)
        flat[idx + 1] = words[1].astype(np.uint16)
    return LineBatch(offsets=offsets, lines=lines, n_lines=int(q.shape[0]))


def unpack_lines(batch: LineBatch, env: int):
    """Decode one environment's block back into ``{sku: [(region, qty), ...]}`` in stream sequence - the inverse of
    ``pack_lines`` up to the interleaving of different SKUs, which the allocation does not depend on."""
    r0, r1 = int(batch.offsets[env]), int(batch.offsets[env + 1])
    out = {}
    if r1 == r0:
        return out
    block = batch.lines[r0:r1].reshape((r1 - r0) // 2, LINE_LANES, 2).transpose(1, 0, 2).reshape(LINE_LANES, -1)   # [lane, entry]
    for lane in range(LINE_LANES):
        sku_map = [int(block[lane, 0]) & 0xff, int(block[lane, 0]) >> 8, int(block[lane, 1]) & 0xff, int(block[lane, 1]) >> 8]
        for v in block[lane, 2:]:
            v = int(v)
            if v == 0:
                break
            out.setdefault(sku_map[v >> 14], []).append(((v >> 8) & 0x3f, v & 0xff))
    return out
