"""CPU oracle for the GAE / value-target scan and advantage standardisation (TEST INFRASTRUCTURE).

PARITY UNPINNED: the reference delegates this arithmetic to ray.rllib 2.52.1 (requirements.txt:81;
configured at src/algorithms/ippo.py:145-160, mappo.py:142-157, cppo.py:145-160), which is neither
vendored in /root/reference nor installed, and no reference test pins its numbers
(tests/test_algorithms.py:253-268 only checks that train() returns a dict). This file restates the
published semantics of RLlib's new-API-stack ``GeneralAdvantageEstimation`` connector /
``compute_value_targets``:

    target_t = r_t + gamma * ((1 - lambda) * V_{t+1} + lambda * target_{t+1}),   target_T := V_T
    adv_t    = target_t - V_t
    truncated episodes bootstrap from the value of their final observation, terminated ones from 0,
    and the recursion restarts at every episode boundary;
    per-module standardisation: (adv - mean) / max(1e-4, std).

It also checks itself against the textbook delta form of GAE(lambda) in tests/test_gae_oracle.py.
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def gae_targets(rewards: np.ndarray, values: np.ndarray, gamma: float, lam: float,
                cut: Optional[np.ndarray] = None, cut_values: Optional[np.ndarray] = None):
    """rewards [T,N], values [T+1,N] (last row = bootstrap), cut [T] bool, cut_values [T,N].
    Returns (adv, targets) float32 [T,N]; float32 arithmetic in the kernel's operation order."""
    r = np.asarray(rewards, np.float32)
    v = np.asarray(values, np.float32)
    T, N = r.shape
    g, l = np.float32(gamma), np.float32(lam)
    gl, g1 = np.float32(g * l), np.float32(g * (np.float32(1.0) - l))
    adv = np.zeros((T, N), np.float32)
    tgt = np.zeros((T, N), np.float32)
    v_next = v[T].copy()
    t_next = v_next.copy()
    for t in range(T - 1, -1, -1):
        if cut is not None and cut[t]:
            v_next = (np.asarray(cut_values[t], np.float32) if cut_values is not None else np.zeros(N, np.float32)).copy()
            t_next = v_next.copy()
        cur = (r[t] + (g1 * v_next + gl * t_next)).astype(np.float32)
        tgt[t] = cur
        adv[t] = cur - v[t]
        t_next = cur
        v_next = v[t]
    return adv, tgt


def gae_delta_form(rewards, values, gamma, lam, cut=None, cut_values=None):
    """Textbook GAE(lambda): delta_t = r_t + gamma V_{t+1} - V_t, A_t = delta_t + gamma lambda A_{t+1} (float64)."""
    r = np.asarray(rewards, np.float64)
    v = np.asarray(values, np.float64)
    T, N = r.shape
    adv = np.zeros((T, N))
    a_next = np.zeros(N)
    v_next = v[T]
    for t in range(T - 1, -1, -1):
        if cut is not None and cut[t]:
            v_next = np.asarray(cut_values[t], np.float64) if cut_values is not None else np.zeros(N)
            a_next = np.zeros(N)
        delta = r[t] + gamma * v_next - v[t]
        adv[t] = delta + gamma * lam * a_next
        a_next = adv[t]
        v_next = v[t]
    return adv, adv + v[:T]


def standardize(x: np.ndarray) -> np.ndarray:
    x64 = np.asarray(x, np.float64)
    return ((x64 - x64.mean()) / max(1e-4, x64.std())).astype(np.float32)
