"""TEST-ONLY: builds tests/emu/emu.cpp (the kernel's per-environment step logic compiled as plain C++,
one thread per environment) and drives it with NumPy buffers through the same C structs as the CUDA
library. Lets the container without a GPU check the step logic against the golden trajectories."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional

import numpy as np

from marlsc_b200 import _capi
from marlsc_b200.demand import OrderBatch
from marlsc_b200.spec import EnvSpec

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "emu", "emu.cpp")
OUT = os.path.join(HERE, "emu", "_build", "libmarlsc_emu.so")
OUT_LANES = os.path.join(HERE, "emu", "_build", "libmarlsc_emu_lanes.so")   # wide-team (per-lane chains) allocation
_DEPS = [SRC, os.path.join(ROOT, "marl-sc_b200", "csrc", "env_core.cuh"),
         os.path.join(ROOT, "marl-sc_b200", "csrc", "spec_build.h"), os.path.join(ROOT, "include", "marlsc_b200.h")]
_libs = {}


def emu_lib(lanes: bool = False):
    if lanes in _libs:
        return _libs[lanes]
    out = OUT_LANES if lanes else OUT
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in _DEPS):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                               f"-DMARLSC_FORCE_LANE_ALLOC={int(lanes)}", "-o", out, SRC])
    L = C.CDLL(out)
    L.emu_env_create.argtypes = [C.POINTER(_capi.EnvSpecC), C.POINTER(C.c_void_p)]
    L.emu_env_reset.argtypes = [C.c_void_p, C.POINTER(_capi.EnvStateC), C.c_void_p, C.c_int, C.c_void_p]
    L.emu_env_step.argtypes = [C.c_void_p, C.POINTER(_capi.EnvStateC), C.POINTER(_capi.StepIOC), C.c_int]
    L.emu_env_step_lean.argtypes = [C.c_void_p, C.POINTER(_capi.EnvStateC), C.POINTER(_capi.StepIOC), C.c_int]
    L.emu_alloc_avail.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.emu_env_destroy.argtypes = [C.c_void_p]
    L.emu_env_destroy.restype = None
    for fn in ("emu_env_obs_dim", "emu_env_needs_history", "emu_env_needs_forecast"):
        getattr(L, fn).argtypes = [C.c_void_p]
    L.emu_last_error.restype = C.c_char_p
    _libs[lanes] = L
    return L


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


class EmuBatch:
    def __init__(self, spec: EnvSpec, num_envs: int, lanes: bool = False):
        L = self.L = emu_lib(lanes)
        self.spec, self.E = spec, num_envs
        self._c = spec.to_c()
        h = C.c_void_p()
        rc = L.emu_env_create(C.byref(self._c), C.byref(h))
        if rc != 0:
            raise ValueError(L.emu_last_error().decode())
        self.h = h
        s = spec.scalars
        self.W, self.S, self.R, self.D = s["n_warehouses"], s["n_skus"], s["n_regions"], s["ring_depth"]
        self.obs_dim = L.emu_env_obs_dim(h)
        E, W, S, D = num_envs, self.W, self.S, self.D
        self.inv = np.zeros((E, W, S), np.int32)
        self.ring_q = np.full((E, D, W, S), -7, np.int32)           # garbage: reset must clear it
        self.ring_l = np.full((E, D, W, S), 9, np.uint8) if s["lead_mode"] == 1 else None
        self.hist = np.full((E, 5, W, S), -3, np.int32) if L.emu_env_needs_history(h) else None
        self.fcst = np.full((E, W, S), 1.5, np.float32) if L.emu_env_needs_forecast(h) else None
        self.state = _capi.EnvStateC(E, _p(self.inv), _p(self.ring_q), _p(self.ring_l), _p(self.hist), _p(self.fcst))

    def reset(self, init_inventory: np.ndarray) -> np.ndarray:
        init = np.ascontiguousarray(init_inventory, dtype=np.int32)
        obs = np.zeros((self.E, self.W, self.obs_dim), np.float32)
        per_env = int(init.ndim == 3)
        assert self.L.emu_env_reset(self.h, C.byref(self.state), _p(init), per_env, _p(obs)) == 0
        return obs

    def step_lean(self, t: int, actions: np.ndarray, orders: OrderBatch) -> Dict[str, np.ndarray]:
        """The lean kernel instantiation (no diagnostics): inventory, rewards, observations, truncation only."""
        E, W = self.E, self.W
        act = np.ascontiguousarray(actions, dtype=np.float32)
        o = dict(rewards=np.zeros((E, W), np.float32), obs=np.zeros((E, W, self.obs_dim), np.float32), trunc=np.zeros(E, np.uint8))
        io = _capi.StepIOC(_p(act), _p(orders.offsets), _p(orders.region), _p(orders.qty), orders.qty_bytes, None,
                           _p(o["rewards"]), _p(o["obs"]), _p(o["trunc"]), None, None, None, None, None, None, None)
        assert self.L.emu_env_step_lean(self.h, C.byref(self.state), C.byref(io), t) == 0
        o["inventory"] = self.inv.copy()
        return o

    def step(self, t: int, actions: np.ndarray, orders: OrderBatch, leads: Optional[np.ndarray]) -> Dict[str, np.ndarray]:
        E, W, S, R = self.E, self.W, self.S, self.R
        act = np.ascontiguousarray(actions, dtype=np.float32)
        lead = None if leads is None else np.ascontiguousarray(leads, dtype=np.uint8)
        o = dict(rewards=np.zeros((E, W), np.float32), obs=np.zeros((E, W, self.obs_dim), np.float32),
                 trunc=np.zeros(E, np.uint8), cost_breakdown=np.zeros((E, W, 4), np.float32),
                 ordered=np.zeros((E, W, S), np.int32), ship_by_sku=np.zeros((E, W, R, S), np.int32),
                 ship_counts=np.zeros((E, W, R), np.int32), unfulfilled=np.zeros((E, R, S), np.int32),
                 lost_orders=np.zeros((E, R), np.int32), lost_sales=np.zeros((E, W, S), np.float32))
        io = _capi.StepIOC(_p(act), _p(orders.offsets), _p(orders.region), _p(orders.qty), orders.qty_bytes, _p(lead),
                           _p(o["rewards"]), _p(o["obs"]), _p(o["trunc"]), _p(o["cost_breakdown"]), _p(o["ordered"]),
                           _p(o["ship_by_sku"]), _p(o["ship_counts"]), _p(o["unfulfilled"]), _p(o["lost_orders"]),
                           _p(o["lost_sales"]))
        assert self.L.emu_env_step(self.h, C.byref(self.state), C.byref(io), t) == 0
        o["inventory"] = self.inv.copy()
        return o

    def alloc_avail(self, inv: np.ndarray, region: np.ndarray, qty: np.ndarray):
        """One environment's allocation the way env_alloc.cuh does it (availability masks + the library's permutation
        table): returns (inventory after, shipped units [W,R], lost units [R])."""
        inv = np.ascontiguousarray(inv, dtype=np.int32).copy()
        region = np.ascontiguousarray(region, dtype=np.int16)
        qty = np.ascontiguousarray(qty, dtype=np.uint8)
        shipq = np.zeros((self.W, self.R), np.int32)
        lost = np.zeros(self.R, np.int32)
        rc = self.L.emu_alloc_avail(self.h, _p(inv), len(region), _p(region), _p(qty), _p(shipq), _p(lost))
        if rc != 0:
            raise ValueError(self.L.emu_last_error().decode())
        return inv, shipq, lost

    def close(self):
        self.L.emu_env_destroy(self.h)
