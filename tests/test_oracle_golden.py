"""Pins oracle/inventory_oracle.py to the reference: replaying the recorded demand, lead times and
actions must reproduce the reference env's own trajectories (integers exactly, floats to 1e-12)."""
import numpy as np
import pytest

from golden_io import FLOAT_KEYS, INT_KEYS, NAMES, Golden
from oracle.inventory_oracle import OracleEnv, agent_observation, local_obs_dim


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_reference(name):
    g = Golden(name)
    for i in range(g.N):
        env = OracleEnv(g.env, **g.oracle_kwargs())
        obs0 = env.reset(g["init_inventory"][i])
        np.testing.assert_allclose(obs0, g["obs0_local"][i], rtol=1e-6 if g.meta.get("lite") else 1e-12, atol=0)
        for t in range(g.T):
            out = env.step(g["actions"][i, t], g.orders(i, t), g.leads(i, t))
            for k in INT_KEYS:
                if k in g.z:                             # "lite" goldens do not carry the [W,R,S] shipment cube
                    assert np.array_equal(out[k], g[k][i, t]), (name, i, t, k)
            for k in FLOAT_KEYS:
                # observations of "lite" goldens are stored as the float32 the reference emits
                tol = 1e-6 if (k == "obs_local" and g.meta.get("lite")) else 1e-12
                np.testing.assert_allclose(out[k], g[k][i, t], rtol=tol, atol=tol, err_msg=f"{name} env {i} step {t} {k}")
            assert bool(out["trunc"]) == bool(g["trunc"][i, t])


def test_obs_dim_matches_reference_probe():
    # all 12 features + 6 aggregates + id at W3 S2 L3 -> 45
    from golden.scenarios import ALL_FEATURES_ON
    assert local_obs_dim(2, 3, 3, ALL_FEATURES_ON, True) == 45


def test_agent_observation_layout():
    loc = np.arange(6, dtype=float).reshape(2, 3)
    full = agent_observation(loc)
    assert full.shape == (2, 9)
    assert np.array_equal(full[1], [3, 4, 5, 0, 1, 2, 3, 4, 5])
