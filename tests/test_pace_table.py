"""Host logic, no GPU: the overtake pace table the library derives (include/mcgp.h: mcgp_pace_table) must give, for
EVERY (ahead driver, its tyre age, chasing driver, its tyre age, DRS) combination, the truth value of the reference's
FP64 pair test (src/simulation.py:514-521)

    pace = base_pace + tire_age * tire_deg;  pace_delta = pace_ahead - pace_behind;
    if car_behind.drs_enabled: pace_delta += drs_delta;  pace_delta > overtake_delta

evaluated here with NumPy float64 in the same operation order.  The BASELINE synthetic inputs are round numbers, so
thousands of combinations land exactly on the threshold, where only faithful FP64 rounding gives the reference's answer
(the finding behind the table: DESIGN.md, "Native mode").  Also checked: the float images are strictly increasing in
the FP64 pace (that is what makes ONE float compare exact) and stay within an ulp or two of pace * 2^15."""
import numpy as np
import pytest

import golden_cases as gc


@pytest.fixture(scope="module")
def mcgp():
    import mcgp_b200
    return mcgp_b200


def _params(mcgp, cfg, mc):
    sim = mcgp.simulation.RaceSimulator(mcgp.simulation.RaceConfig(**cfg), device=0, pop_no_medium="SOFT", pop_no_soft="MEDIUM")
    return sim._params(mc["grid_probs"], mc["base_pace"], mc["tire_deg"], mc["driver_variance"], mc.get("driver_dnf_rates"),
                       mc.get("track_condition", "dry"))


def _check(mcgp, cfg, mc, max_rows=None):
    p = _params(mcgp, cfg, mc)
    tab = mcgp.capi.pace_table(p)
    n, rows = p.n_drivers, p.total_laps + 5
    assert tab.shape == (rows, 20 if n <= 20 else 32, 4)
    base = np.array([p.base_pace[d] for d in range(n)], np.float64)
    deg = np.array([p.tire_deg[d] for d in range(n)], np.float64)
    ages = np.arange(rows, dtype=np.float64)
    P = base[None, :] + ages[:, None] * deg[None, :]            # [age, driver], two roundings like upstream
    f, thr0, thr1 = tab[:, :n, 0], tab[:, :n, 1], tab[:, :n, 2]
    # strictly increasing image: P < P'  =>  f < f',  P == P'  =>  f == f'
    order = np.argsort(P, axis=None, kind="stable")
    ps, fs = P.ravel()[order], f.ravel()[order]
    dp, df = np.diff(ps), np.diff(fs.astype(np.float64))
    assert np.all(df[dp > 0] > 0) and np.all(df[dp == 0] == 0)
    exact = (P * 32768.0).astype(np.float32)
    ulps = np.abs(f.astype(np.float64) - exact.astype(np.float64)) / np.spacing(np.abs(exact)).astype(np.float64)
    assert ulps.max() <= 64, "float images drifted from pace * 2^15"
    # every combination: table decision == FP64 decision
    R = rows if max_rows is None else min(rows, max_rows)
    n_tie = n_comb = 0
    for drs, thr in ((False, thr0), (True, thr1)):
        for b in range(n):                                       # chasing driver
            pb = P[:R, b]                                        # [A_b]
            delta = P[:R, :, None] - pb[None, None, :]           # [A_a, a, A_b]
            if drs:
                delta = delta + p.drs_delta
            near_tie = np.abs(delta - p.overtake_delta) < 1e-9    # equal in exact arithmetic: FP64 rounding decides
            want = delta > p.overtake_delta
            got = f[:R, :, None] >= thr[:R, b][None, None, :]
            bad = np.argwhere(want != got)
            assert bad.size == 0, f"drs={drs} chasing {b}: {len(bad)} combinations differ, first (A_a, a, A_b) = {bad[0]}"
            n_tie += int(near_tie.sum())
            n_comb += want.size
    return n_comb, n_tie


@pytest.mark.parametrize("name", ["bahrain", "monaco_sc", "sprint19", "season:7", "point:quali"])
def test_table_reproduces_the_fp64_pair_test_on_baseline_workloads(mcgp, name):
    cfg, mc = mcgp.workloads.workload(name)
    n_comb, n_tie = _check(mcgp, cfg, mc)
    print(name, n_comb, "combinations,", n_tie, "on the threshold up to FP64 rounding")
    if name == "bahrain":
        assert n_tie > 100, "the BASELINE round-number inputs are expected to produce exact ties (that is the point)"


@pytest.mark.parametrize("case", ["defaults", "tight", "small_grids", "single", "one_lap"])
def test_table_on_golden_cases(mcgp, case):
    cfg, mc, _, _ = gc.get_case(case)
    _check(mcgp, cfg, mc, max_rows=40)


def test_table_on_random_and_degenerate_inputs(mcgp):
    rng = np.random.default_rng(5)
    cfg, mc = mcgp.workloads.workload("bahrain")
    D = list(mc["base_pace"])
    for trial in range(6):
        c, m = dict(cfg), dict(mc)
        c["total_laps"] = int(rng.integers(1, 70))
        c["overtake_delta"] = float(rng.choice([0.0, 0.3, 0.6, 1.5, -0.2, rng.uniform(0, 2)]))
        c["drs_delta"] = float(rng.choice([0.0, 0.3, rng.uniform(0, 1)]))
        if trial % 2:
            m["base_pace"] = {d: float(rng.uniform(70, 110)) for d in D}
            m["tire_deg"] = {d: float(rng.uniform(0, 0.1)) for d in D}
        else:  # everybody equal / multiples of 2^-k: every pair sits on a lattice
            m["base_pace"] = {d: 90.0 + 0.125 * (k % 4) for k, d in enumerate(D)}
            m["tire_deg"] = {d: 0.0625 * (k % 3) for k, d in enumerate(D)}
        _check(mcgp, c, m, max_rows=48)
    # paces closer than a float ulp of pace * 2^15 must still map to strictly increasing floats
    m = dict(mc)
    m["base_pace"] = {d: 92.0 + 1e-9 * k for k, d in enumerate(D)}
    m["tire_deg"] = {d: 0.0 for d in D}
    _check(mcgp, dict(cfg, total_laps=5), m)


@pytest.mark.parametrize("name", ["bahrain", "monaco_sc", "sprint19", "season:11"])
def test_library_and_mirror_build_the_same_float_images(mcgp, name):
    """The kernel's host-built table (C++, libmcgp.so) and the scalar mirror's (C, oracle/) are written independently;
    the bit-exact kernel == mirror GPU tests rest on them being identical."""
    from oracle import pyoracle as po
    cfg, mc = mcgp.workloads.workload(name)
    tab = mcgp.capi.pace_table(_params(mcgp, cfg, mc))
    mir = po.native_op32_table(po.make_params(cfg, mc, "SOFT", "MEDIUM"))
    n = mir.shape[1]
    assert np.array_equal(tab[:, :n, 0].view(np.uint32), mir.view(np.uint32))
