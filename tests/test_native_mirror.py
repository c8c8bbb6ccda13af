"""The native-mode algorithm (as restated by the scalar mirror, oracle/native_mirror.c) against the reference
model (FP64 oracle): statistical agreement on CPU, so the native *specification* is validated without a GPU.
The GPU kernel is then held bit-exact to the mirror (tests/test_gpu_native.py)."""
import numpy as np
import pytest

import golden_cases as gc
from stats_util import assert_agree_two_stage


@pytest.mark.parametrize("case,n", [("bahrain_dry", 250000), ("sprint19", 400000), ("events", 150000)])
def test_mirror_statistics_match_reference(oracle, case, n):
    cfg, mc, seed, _ = gc.get_case(case)
    params = oracle.make_params(cfg, mc)
    for exact in (True, False):
        res = assert_agree_two_stage(
            lambda m, st: oracle.run_native(params, seed=17 + exact + 10 * st, n_sims=m, exact=exact, threads=8)["hist"],
            lambda m, st: oracle.run_monte_carlo(cfg, mc, m, 4321 + st, threads=8), n, n, f"{case} exact={exact}")
        print(case, exact, res)


def test_mirror_is_counter_based(oracle):
    cfg, mc, seed, _ = gc.get_case("attrition")
    p = oracle.make_params(cfg, mc)
    whole = oracle.run_native(p, 5, 3000, detail=True)
    a = oracle.run_native(p, 5, 1000, sim_begin=0, detail=True)
    b = oracle.run_native(p, 5, 2000, sim_begin=1000, detail=True)
    assert np.array_equal(whole["finish"], np.concatenate([a["finish"], b["finish"]]))
    assert np.array_equal(whole["hist"], a["hist"] + b["hist"])
    assert not np.array_equal(whole["hist"], oracle.run_native(p, 6, 3000)["hist"])
    assert not np.array_equal(whole["hist"], oracle.run_native(p, 5, 3000, stream=1)["hist"])


def test_native_draw_plan_is_uniform(oracle):
    """Lap-1 DNF frequency of the mirror matches 4 x team rate: checks thresholds + Philox word usage."""
    cfg, mc, seed, _ = gc.get_case("one_lap")
    p = oracle.make_params(cfg, mc)
    n = 400000
    out = oracle.run_native(p, 3, n, detail=True, threads=8)
    # with total_laps == 1 the only way to be classified behind a running car is a lap-1 retirement
    rates = np.array([4 * cfg["dnf_rates"][cfg["driver_teams"][d]] for d in mc["grid_probs"]])
    expect_dnf_per_race = rates.sum()
    # position p (0-based) is taken by a retired car iff at least 20-p cars retired; count retirements via times<0
    retired = (out["times"] < 0).sum(1)
    assert abs(retired.mean() - expect_dnf_per_race) < 4 * np.sqrt(expect_dnf_per_race / n)


# Random123's known-answer vectors (kat_vectors, philox4x32 with 7 and 10 rounds): counter, key -> output
_PHILOX_KAT = [
    (7, [0, 0, 0, 0], [0, 0], [0x5f6fb709, 0x0d893f64, 0x4f121f81, 0x4f730a48]),
    (7, [0xffffffff] * 4, [0xffffffff] * 2, [0x5207ddc2, 0x45165e59, 0x4d8ee751, 0x8c52f662]),
    (7, [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0x4dfccaba, 0x190a87f0, 0xc47362ba, 0xb6b5242a]),
    (10, [0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    (10, [0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    (10, [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


def test_philox_known_answers(oracle):
    """The mirror's block function is Philox4x32-R as published (the kernel is held bit-exact to the mirror)."""
    import ctypes as C
    L = oracle.lib()
    L.orc_philox4x32.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_philox4x32.restype = None
    for rounds, ctr, key, want in _PHILOX_KAT:
        c, k, o = np.array(ctr, np.uint32), np.array(key, np.uint32), np.zeros(4, np.uint32)
        L.orc_philox4x32(rounds, c.ctypes.data, k.ctypes.data, o.ctypes.data)
        assert o.tolist() == want, (rounds, [hex(x) for x in o])


def test_mirror_and_kernel_use_the_same_round_count(oracle):
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "monte-carlo-gp_b200", "csrc", "native_math.cuh")).read()
    kernel_rounds = int(re.search(r"#define MCGP_PHILOX_ROUNDS (\d+)", src).group(1))
    L = oracle.lib()
    assert L.orc_native_philox_rounds() == kernel_rounds == 7
