"""Loader for the committed golden trajectories (tests/golden/*.npz, made by make_golden.py)."""
from __future__ import annotations

import json
import os
from typing import Any, Dict, List, Tuple

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["small_default", "allfeat_ratio_stochastic", "basestock_cost_meanstd",
         "regions_ne_warehouses", "large_network", "large_network_long", "large_network_basestock"]

INT_KEYS = ("inventory", "pending", "ordered", "fulfilled", "unfulfilled", "ship_counts", "ship_qty",
            "ship_by_sku", "lost_orders")
FLOAT_KEYS = ("lost_sales", "cost_hold", "cost_pen", "cost_out", "cost_in", "rewards", "obs_local")


class Golden:
    def __init__(self, name: str):
        z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
        self.name = name
        self.z = {k: z[k] for k in z.files}
        if "obs_whole" in self.z:                      # "lite" files: whole-number entries as int16 + the others as a list
            obs = self.z.pop("obs_whole").astype(np.float32)
            obs.reshape(-1)[self.z.pop("obs_frac_idx")] = self.z.pop("obs_frac_val")
            self.z["obs_local"] = obs.astype(np.float64)
        self.env: Dict[str, Any] = json.loads(str(self.z["env_json"]))
        self.meta: Dict[str, Any] = json.loads(str(self.z["meta_json"]))
        self.N, self.T = self.meta["n_envs"], self.meta["steps"]
        self.W, self.S, self.R = self.env["n_warehouses"], self.env["n_skus"], self.env["n_regions"]
        self.obs_stats = (self.z["obs_mean"], self.z["obs_std"]) if "obs_mean" in self.z else None

    def __getitem__(self, k):
        return self.z[k]

    def orders(self, env: int, t: int) -> List[Tuple[int, np.ndarray]]:
        ptr = self.z["order_ptr"]
        a, b = ptr[env * self.T + t], ptr[env * self.T + t + 1]
        return [(int(self.z["order_region"][j]), self.z["order_qty"][j].astype(np.float64)) for j in range(a, b)]

    def oracle_kwargs(self):
        return dict(obs_normalization=self.meta["obs_normalization"], obs_stats=self.obs_stats,
                    include_warehouse_id=self.meta["include_warehouse_id"])

    def leads(self, env: int, t: int):
        return self.z["lead_times"][env, t].astype(np.int64) if self.meta["stochastic_lead"] else None
