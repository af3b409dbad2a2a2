"""The CUDA kernel's per-environment step logic (env_core.cuh), compiled for the host with one thread
per environment, replayed against the reference's golden trajectories. Catches indexing / arithmetic
mistakes without a GPU; the multi-lane paths are covered by the -m gpu tests."""
import numpy as np
import pytest

from emu_backend import EmuBatch
from golden_io import NAMES, Golden
from parity_common import compare_step, spec_for, step_orders


@pytest.mark.parametrize("name", NAMES)
def test_emu_matches_reference(name):
    g = Golden(name)
    _, spec = spec_for(g)
    b = EmuBatch(spec, g.N)
    obs0 = b.reset(g["init_inventory"])
    np.testing.assert_allclose(obs0, g["obs0_local"], rtol=1e-5, atol=1e-6)
    stochastic = g.meta["stochastic_lead"]
    for t in range(g.T):
        out = b.step(t, g["actions"][:, t], step_orders(g, t), g["lead_times"][:, t] if stochastic else None)
        compare_step(g, t, out, what="emu ")
    b.close()


def test_emu_region_map_equals_premapped_demand():
    """Raw region ids + an in-kernel raw->included map (reference preprocessor.py:382-441 does the remap
    offline) must give the same trajectory as demand whose regions were mapped beforehand."""
    g = Golden("small_default")
    R = g.R
    region_map = list(range(R)) * 2                      # raw ids r and r + R both mean region r
    _, spec = spec_for(g, region_map=region_map)
    b = EmuBatch(spec, g.N)
    b.reset(g["init_inventory"])
    for t in range(20):
        shifted = step_orders(g, t, region_shift=lambda i, tt, j: R * ((i + tt + j) % 2))
        compare_step(g, t, b.step(t, g["actions"][:, t], shifted, None), what="emu remap ")
    b.close()


# lanes=True: the wide-team allocation (independent per-lane chains) forced onto the one-lane emulation
@pytest.mark.parametrize("lanes", [False, True], ids=["dense", "lanes"])
@pytest.mark.parametrize("name", ["small_default", "regions_ne_warehouses", "large_network"])
def test_emu_lean_instantiation_matches_reference(name, lanes):
    """Configurations the lean kernel instantiation covers (independent per-SKU allocation chains, no
    diagnostics) replayed through it on the CPU."""
    g = Golden(name)
    _, spec = spec_for(g)
    b = EmuBatch(spec, g.N, lanes=lanes)
    b.reset(g["init_inventory"])
    for t in range(g.T):
        compare_step(g, t, b.step_lean(t, g["actions"][:, t], step_orders(g, t)), what="emu lean ")
    b.close()


@pytest.mark.parametrize("W,S,R", [(3, 5, 4), (10, 100, 50), (16, 33, 7)])
def test_availability_mask_allocation_equals_the_sequential_greedy(W, S, R):
    """The allocation kernel of one-warp teams (env_alloc.cuh) ships only from warehouses that hold the SKU, found through
    a host-built table that permutes an availability mask into a region's priority order. Restated on the CPU over
    the library's own tables (tests/emu/emu.cpp) it must reproduce the reference's sequential greedy allocation
    (demand_allocator.py:150-208, as restated in the oracle) unit for unit - scarce stock, so most lines split or lose."""
    from marlsc_b200.config import environment_config_from_dict
    from marlsc_b200.spec import build_env_spec
    rng = np.random.default_rng(W * 100 + S)
    out_var = np.stack([rng.permutation(W) for _ in range(R)], 1) * 0.05 + 0.05          # tie-free -> static priority
    env = dict(
        action_space=dict(type="direct", params=dict(max_order_quantities=[10] * S)),
        n_warehouses=W, n_skus=S, n_regions=R, episode_length=5, max_wh_capacities=[1e7] * W,
        initial_inventory=dict(type="custom", params=dict(values=5)),
        cost_structure=dict(holding_cost=1.0, penalty_cost=2.0,
                            shipment_cost=dict(outbound_fixed=np.zeros((W, R)).tolist(), outbound_variable=out_var.tolist(),
                                               inbound_fixed=np.zeros((W, S)).tolist(), inbound_variable=np.ones((W, S)).tolist()),
                            sku_weights=[1.0] * S, distances=(rng.integers(10, 500, (W, R)) * 1.0).tolist()),
        components=dict(demand_sampler=dict(type="poisson", params=dict(lambda_orders=1.0, probability_skus=0.3, lambda_quantity=5.0)),
                        demand_allocator=dict(type="greedy", params=dict(max_splits=W - 1)),
                        lead_time_sampler=dict(type="fixed", params=dict(expected_lead_times=np.ones((W, S), int).tolist())),
                        lost_sales_handler=dict(type="shipment", params=None),
                        reward_calculator=dict(type="cost", params=dict(scope="agent", scale_factor=1.0, cost_weights=[0.25] * 4))),
        data_source=dict(type="custom"),
        features=dict(inventory=True, pipeline=True, incoming_demand_home=False, units_shipped_home=False, units_shipped_away=False,
                      stockout=False, rolling_demand_mean=True, demand_forecast=False, days_of_supply=False,
                      net_inventory_position=False, demand_variability=False, demand_history=False, inventory_aggregate=True,
                      pipeline_aggregate=False, incoming_demand_home_aggregate=False, units_shipped_away_aggregate=False,
                      rolling_demand_mean_aggregate=False, demand_forecast_aggregate=False))
    cfg = environment_config_from_dict(dict(env, allow_region_mismatch=True))
    b = EmuBatch(build_env_spec(cfg), 1)
    for trial in range(20):
        inv = rng.integers(0, 9, (W, S)) * (rng.random((W, S)) < 0.5)
        n = int(rng.integers(0, 90))
        region = rng.integers(0, R, n)
        qty = np.where(rng.random((n, S)) < 0.3, rng.integers(1, 15, (n, S)), 0)
        # the reference's order-by-order greedy (oracle/inventory_oracle.py step 4./5.; ties cannot occur here)
        avail, ship, lost = inv.astype(np.int64).copy(), np.zeros((W, R), np.int64), np.zeros(R, np.int64)
        for j in range(n):
            rem = qty[j].astype(np.int64).copy()
            for w in np.argsort(out_var[:, region[j]], kind="stable"):
                f = np.minimum(rem, avail[w])
                ship[w, region[j]] += f.sum()
                rem -= f
                avail[w] -= f
            lost[region[j]] += rem.sum()
        got_inv, got_ship, got_lost = b.alloc_avail(inv, region, qty)
        assert np.array_equal(got_inv, avail), trial
        assert np.array_equal(got_ship, ship) and np.array_equal(got_lost, lost), trial
    b.close()
