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
