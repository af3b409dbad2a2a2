"""Generate tests/golden/*.npz from the UNMODIFIED reference env (run in the build container only).

    python tests/golden/make_golden.py [scenario ...]

For every scenario in scenarios.py this steps ``n_envs`` reference ``InventoryEnvironment`` instances
(/root/reference/src/environment/envs/multi_env.py) seeded ``derive_env_seed(base_seed, 0, i)``
(src/utils/seed_manager.py:166-186) with pre-sampled float32 actions, records the demand orders and
lead times its own samplers drew, and stores inputs + every per-step output (SURVEY.md section 8c
parity classes). The committed vectors are what pins ``oracle/inventory_oracle.py`` and, through it,
the CUDA path; ``/root/reference`` itself never travels to the GPU box.

numpy's argsort tie-break is SIMD-dispatch dependent, so the script re-execs itself with
NPY_DISABLE_CPU_FEATURES set (see oracle/ref_harness.py).
"""
from __future__ import annotations

import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import ref_harness as H  # noqa: E402

if os.environ.get("NPY_DISABLE_CPU_FEATURES") != H.STABLE_SORT_ENV["NPY_DISABLE_CPU_FEATURES"]:
    os.execvpe(sys.executable, [sys.executable] + sys.argv, {**os.environ, **H.STABLE_SORT_ENV})

import numpy as np  # noqa: E402

import scenarios as SC  # noqa: E402

GRID = np.array([-1.0, 1.0, 0.0, 0.5, -0.5, 0.25, -0.25, 0.3, 0.75, -0.75, 0.125, 0.9], dtype=np.float32)


def make_actions(seed, shape):
    rng = np.random.default_rng(seed)
    a = rng.uniform(-1, 1, shape).astype(np.float32)
    pick = rng.random(shape) < 0.15            # exact grid points exercise round-half-even and the clip
    a[pick] = GRID[rng.integers(0, len(GRID), size=int(pick.sum()))]
    return a


def check_small_default_matches_yaml():
    ref = H.load_env_config("config_files/environments/env_symmetric_3WH2SKU.yaml")
    mine = H.env_config_from_dict(SC.small_default())
    assert ref.model_dump() == mine.model_dump(), "small_default() drifted from the reference YAML"


def generate(name: str) -> str:
    sc = SC.scenario(name)
    env_dict = sc["env"]
    cfg = H.env_config_from_dict(env_dict, allow_region_mismatch=sc["allow_region_mismatch"])
    N, T = sc["n_envs"], sc["steps"]
    W, S, R = env_dict["n_warehouses"], env_dict["n_skus"], env_dict["n_regions"]
    actions = make_actions(sc["action_seed"], (N, T, W, S))
    stochastic = env_dict["components"]["lead_time_sampler"]["type"] == "stochastic"

    probe = H.make_env(cfg, seed=0, env_meta=dict(include_warehouse_id=False))
    stats = SC.obs_stats_for(sc, probe._compute_local_obs_dim())
    meta = dict(obs_normalization=sc["obs_normalization"], obs_stats=stats,
                include_warehouse_id=sc["include_warehouse_id"])

    keys = None
    stacks = {}
    ptr, regions, qtys, leads, seeds = [0], [], [], [], []
    for i in range(N):
        seed = H.derive_env_seed(sc["base_seed"], 0, i)
        seeds.append(seed)
        env = H.make_env(cfg, seed=seed, env_meta=meta)
        dem = H.DemandRecorder(env)
        ltr = H.LeadTimeRecorder(env)
        res = H.run_episode(env, actions[i])
        assert len(dem.steps) == T and len(ltr.steps) == T
        for step in dem.steps:
            for r, q in step:
                regions.append(r)
                qtys.append(q)
            ptr.append(len(regions))
        leads.append(np.stack(ltr.steps))
        keys = keys or list(res.keys())
        for k in keys:
            stacks.setdefault(k, []).append(res[k])

    out = {k: np.stack(v) for k, v in stacks.items()}
    int_keys = ("init_inventory", "inventory", "pending", "ordered", "fulfilled", "unfulfilled",
                "ship_counts", "ship_qty", "ship_by_sku", "lost_orders")
    for k in int_keys:
        assert np.array_equal(out[k], np.round(out[k])), k
        out[k] = out[k].astype(np.int32)
    if env_dict["components"]["lost_sales_handler"]["type"] == "closest":
        assert np.array_equal(out["lost_sales"], np.round(out["lost_sales"]))
    qty_arr = np.asarray(qtys, dtype=np.float64).reshape(-1, S)
    assert np.array_equal(qty_arr, np.round(qty_arr)) and qty_arr.max(initial=0) < 32767
    payload = dict(
        env_json=np.array(json.dumps(env_dict)),
        meta_json=np.array(json.dumps(dict(
            obs_normalization=sc["obs_normalization"], include_warehouse_id=sc["include_warehouse_id"],
            allow_region_mismatch=sc["allow_region_mismatch"], base_seed=sc["base_seed"],
            n_envs=N, steps=T, stochastic_lead=stochastic,
            numpy=np.__version__, reference="Jakoebly/marl-sc @ /root/reference"))),
        env_seeds=np.asarray(seeds, dtype=np.int64),
        actions=actions,
        order_ptr=np.asarray(ptr, dtype=np.int64),          # CSR over (env, step): row = i*T + t
        order_region=np.asarray(regions, dtype=np.int16),
        order_qty=qty_arr.astype(np.int16),
        lead_times=np.stack(leads).astype(np.int8),          # [N, T, W, S] actual lead times drawn
        **out)
    if stats is not None:
        payload["obs_mean"], payload["obs_std"] = stats
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **payload)
    print(f"{name}: {N} envs x {T} steps, {len(regions)} orders, {os.path.getsize(path) / 1024:.0f} KiB")
    return path


if __name__ == "__main__":
    check_small_default_matches_yaml()
    for nm in (sys.argv[1:] or list(SC.SCENARIOS)):
        generate(nm)
