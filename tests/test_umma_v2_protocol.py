"""CPU model check of the barrier protocol of the opt-in second-generation tcgen05 GEMM (`umma_gemm_nt_v2_kernel`,
csrc/agp_umma.cu): random schedules of the CTA's roles must never read a stale tile, overwrite live data or deadlock,
and the checker itself must catch each protocol mutation (a dropped wait)."""
import os
import random
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
import umma_v2_protocol_sim as sim  # noqa: E402


def test_protocol_random_schedules():
    assert sim.run_trials(trials=60, seed=7) > 0


@pytest.mark.parametrize("units", [[4], [4, 4, 4, 4, 4, 4, 4, 4], [16, 12, 8, 4, 4, 8, 12, 16], [1, 1, 1, 2, 3], [256]])
@pytest.mark.parametrize("presplit", [True, False])
def test_protocol_unit_shapes(units, presplit):
    # C2 shapes: 4..16 k-blocks per unit (triangular operand), 256 per Gram split slice
    for seed, slow in enumerate([(), (2, 3, 4, 5), (6, 7, 8, 9), ("tensor",), (0,)]):
        sim.simulate(units, presplit=presplit, seed=seed, slow=slow)


@pytest.mark.parametrize("bug,slow", [("slot", ()), ("a_empty", (6, 7, 8, 9)), ("b_stage", ("tensor",)), ("tmem_empty", (2, 3, 4, 5))])
def test_checker_catches_dropped_waits(bug, slow):
    caught = 0
    for seed in range(10):
        try:
            sim.simulate([12, 8, 16, 4, 9], presplit=bool(seed & 1), seed=seed, bug=bug, slow=slow)
        except AssertionError:
            caught += 1
    assert caught >= 8
