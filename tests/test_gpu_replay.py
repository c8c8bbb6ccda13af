"""Replay-mode CUDA path (through the C ABI) against the reference fixtures and the CPU oracle.

Bar: bit-exact finishing orders, grids, DNF laps, draw consumption AND race times (north_star allows
1e-5 relative on times; the kernel is held to 0 ulp).
"""
import numpy as np
import pytest

import golden_cases as gc
from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu


def _case_of(name):
    return name.split("_h")[0] if name.rsplit("_h", 1)[-1].isdigit() else name


@pytest.fixture(scope="module")
def mcgp():
    import mcgp_b200
    return mcgp_b200


def _replay(mcgp, oracle, cfg, mc, seed, n_sims, pop_a, pop_b, serial_grid=False):
    oparams = oracle.make_params(cfg, mc, pop_a, pop_b)
    ref = oracle.run_streams(oparams, oracle.Rng(seed), n_sims, detail=True, tapes=True)
    sim = mcgp.simulation.RaceSimulator(mcgp.simulation.RaceConfig(**cfg), pop_no_medium=pop_a, pop_no_soft=pop_b)
    kw = {k: mc.get(k) for k in ("grid_probs", "base_pace", "tire_deg", "driver_variance", "driver_dnf_rates")}
    got = sim.replay(**kw, track_condition=mc.get("track_condition", "dry"), u_py=ref["tape_upy"], z=ref["tape_z"],
                     u_np=ref["tape_unp"], offsets=ref["tape_off"], serial_grid=serial_grid)
    return ref, got


@pytest.mark.parametrize("name", golden_names())
def test_replay_matches_reference_fixture(mcgp, oracle, name):
    g = load_golden(name)
    meta = g["meta"]
    cfg, mc, seed, _ = gc.get_case(_case_of(name))
    n_sims = min(meta["n_sims"], 3000)
    ref, got = _replay(mcgp, oracle, cfg, mc, seed, n_sims, meta["pop_no_medium"], meta["pop_no_soft"])
    k = min(g["finish"].shape[0], n_sims)
    # against the fixture recorded from the unmodified reference
    assert np.array_equal(got["grid"][:k], g["grid"][:k]), "sampled grids differ from the reference"
    assert np.array_equal(got["finish"][:k], g["finish"][:k]), "finishing orders differ from the reference"
    assert np.array_equal(got["dnf_lap"][:k], g["dnf_lap"][:k])
    assert np.array_equal(got["times"][:k].view(np.uint64), g["times"][:k].view(np.uint64)), "race times not bit-exact"
    # against the oracle on every sim, including how many draws each sim consumed
    assert np.array_equal(got["finish"], ref["finish"])
    assert np.array_equal(got["times"].view(np.uint64), ref["times"].view(np.uint64))
    assert np.array_equal(got["used"].cumsum(0), ref["draws"])
    assert np.array_equal(got["hist"].astype(np.int64), ref["hist"])
    if n_sims == meta["n_sims"]:
        assert np.array_equal(got["hist"].astype(np.int64), g["hist"])


@pytest.mark.parametrize("name", golden_names())
def test_replay_serial_grid_path_gives_the_same_races(mcgp, oracle, name):
    """_sample_grid: the kernel's certified parallel selection and the serial evaluation in the reference's operation
    order (which it falls back to where it cannot certify) are both bit-exact to the reference on every fixture."""
    g = load_golden(name)
    meta = g["meta"]
    cfg, mc, seed, _ = gc.get_case(_case_of(name))
    n_sims = min(meta["n_sims"], 1000)
    ref, got = _replay(mcgp, oracle, cfg, mc, seed, n_sims, meta["pop_no_medium"], meta["pop_no_soft"], serial_grid=True)
    k = min(g["finish"].shape[0], n_sims)
    assert np.array_equal(got["grid"][:k], g["grid"][:k]), "serial-path grids differ from the reference"
    assert np.array_equal(got["finish"], ref["finish"])
    assert np.array_equal(got["times"].view(np.uint64), ref["times"].view(np.uint64))
    assert np.array_equal(got["used"].cumsum(0), ref["draws"])


def _boundary_tapes(mc, laps, n_sims, seed=7):
    """Worst-case-sized synthetic tapes whose FIRST grid draw of every sim sits on (or one ulp either side of) a
    boundary of the reference's cdf = cumsum(p / sum(p)) / cumsum(...)[-1] of grid position 0."""
    drivers = list(mc["grid_probs"])
    n = len(drivers)
    n_py, n_z = n + (laps - 1) * (4 + n + 3 * (n - 1)), 2 * n + (laps - 1) * n
    rng = np.random.default_rng(seed)
    upy, z, unp = rng.random(n_sims * n_py), rng.standard_normal(n_sims * n_z), rng.random(n_sims * n)
    off = np.arange(n_sims + 1, dtype=np.int64)[:, None] * np.array([n_py, n_z, n], np.int64)
    p = [float(mc["grid_probs"][d][0]) for d in drivers]
    tot = sum(p)                                  # builtin sum() as upstream (:123)
    cdf = np.array([x / tot for x in p]).cumsum()
    cdf = cdf / cdf[-1]
    inner = np.flatnonzero(np.diff(np.concatenate([[0.0], cdf])) > 0)[:-1]
    first = np.zeros(n_sims, np.int64)            # searchsorted(cdf, u, side='right') of the edited draw
    for s in range(n_sims):
        b = cdf[rng.choice(inner)]
        unp[off[s, 2]] = [np.nextafter(b, 0.0), b, np.nextafter(b, 1.0)][s % 3]
        first[s] = int((cdf <= unp[off[s, 2]]).sum())
    return upy, z, unp, off, first


def test_replay_grid_draw_on_a_boundary_takes_the_serial_path(mcgp, oracle):
    """Uniforms placed ON (and one ulp either side of) the cumulative-probability boundaries of a grid row: the
    parallel selection cannot certify these; the serial path decides them exactly as the oracle does."""
    cfg, mc, _, _ = gc.get_case("bahrain_dry")
    oparams = oracle.make_params(cfg, mc, "SOFT", "MEDIUM")
    upy, z, unp, off, first = _boundary_tapes(mc, cfg["total_laps"], 900)
    want = oracle.run_tapes(oparams, upy, z, unp, off, detail=True)
    assert np.array_equal(want["grid"][:, 0], first)   # the draw one ulp below a boundary selects the driver before it
    sim = mcgp.simulation.RaceSimulator(mcgp.simulation.RaceConfig(**cfg), pop_no_medium="SOFT", pop_no_soft="MEDIUM")
    kw = {k: mc.get(k) for k in ("grid_probs", "base_pace", "tire_deg", "driver_variance", "driver_dnf_rates")}
    for serial in (False, True):
        got = sim.replay(**kw, u_py=upy, z=z, u_np=unp, offsets=off, serial_grid=serial)
        assert np.array_equal(got["grid"], want["grid"]), f"grids differ from the oracle (serial_grid={serial})"
        assert np.array_equal(got["finish"], want["finish"])
        assert np.array_equal(got["times"].view(np.uint64), want["times"].view(np.uint64))


@pytest.mark.parametrize("block", range(4))
def test_replay_random_configurations(mcgp, oracle, block):
    """Randomised sweep of the parameter space (1-32 cars: both kernel instantiations; 1-90 laps; event storms; certain /
    impossible retirements; zero variance: exact ties; one-hot, flat and sparse grids with all-zero rows: the serial
    grid path; all track conditions; both pop knobs): every race bit-exact against the oracle, with the certified scan
    and with the serial grid path forced."""
    import random
    from test_gpu_native import _random_case
    rnd = random.Random(4200 + block)
    for trial in range(6):
        cfg, mc = _random_case(rnd)
        pop = (rnd.choice(["SOFT", "HARD"]), rnd.choice(["MEDIUM", "HARD"]))
        n_sims = 400
        what = f"block {block} trial {trial}: n={len(mc['grid_probs'])} laps={cfg['total_laps']} {mc['track_condition']}"
        for serial in (False, True):
            ref, got = _replay(mcgp, oracle, cfg, mc, rnd.getrandbits(32) if not serial else 99, n_sims, *pop, serial_grid=serial)
            assert np.array_equal(got["grid"], ref["grid"]), what
            bad = np.nonzero((got["finish"] != ref["finish"]).any(1))[0]
            assert bad.size == 0, f"{what}: {bad.size} of {n_sims} races differ from the oracle, first: sim {bad[:5]}"
            assert np.array_equal(got["times"].view(np.uint64), ref["times"].view(np.uint64)), what
            assert np.array_equal(got["dnf_lap"], ref["dnf_lap"]), what
            assert np.array_equal(got["used"].cumsum(0), ref["draws"]), what


@pytest.mark.parametrize("n,laps", [(20, 505), (32, 314), (21, 200), (1, 505), (20, 2)])
def test_replay_extreme_sizes(mcgp, oracle, n, laps):
    """The limits of the boundary: the longest races the pace table admits (505 laps up to 20 cars, 314 beyond), both
    kernel instantiations, a one-car race, a two-lap race -- with events on ~20 % of the laps.  Bit-exact vs the oracle."""
    wl = mcgp.workloads
    D = [f"R{i:02d}" for i in range(n)]
    cfg, _ = wl.workload("bahrain")
    cfg.update(total_laps=laps, driver_teams={d: "Unknown" for d in D}, sc_probability=0.08, vsc_probability=0.1,
               red_flag_probability=0.02)
    mc = dict(grid_probs=wl.gaussian_grid_probs(D, spread=2.0), base_pace={d: 90 + 0.05 * i for i, d in enumerate(D)},
              tire_deg={d: 0.03 for d in D}, driver_variance={d: 0.2 for d in D}, driver_dnf_rates={d: 0.0005 for d in D},
              track_condition="dry")
    ref, got = _replay(mcgp, oracle, cfg, mc, 11, 300, "SOFT", "MEDIUM")
    assert np.array_equal(got["finish"], ref["finish"])
    assert np.array_equal(got["times"].view(np.uint64), ref["times"].view(np.uint64))
    assert np.array_equal(got["used"].cumsum(0), ref["draws"])


def test_replay_tape_overrun_is_an_error(mcgp, oracle):
    cfg, mc, seed, _ = gc.get_case("small_grids")
    oparams = oracle.make_params(cfg, mc)
    ref = oracle.run_streams(oparams, oracle.Rng(seed), 10, detail=True, tapes=True)
    off = ref["tape_off"].copy()
    off[1:, 0] -= 3  # every sim's U_py tape is 3 draws short
    off[1:, 0] = np.maximum(off[1:, 0], off[:-1, 0])
    sim = mcgp.simulation.RaceSimulator(mcgp.simulation.RaceConfig(**cfg))
    kw = {k: mc.get(k) for k in ("grid_probs", "base_pace", "tire_deg", "driver_variance", "driver_dnf_rates")}
    with pytest.raises(mcgp.capi.McgpError) as e:
        sim.replay(**kw, u_py=ref["tape_upy"], z=ref["tape_z"], u_np=ref["tape_unp"], offsets=off)
    assert e.value.code == mcgp.capi.ETAPE


def test_config2_chunked_replay_is_bit_exact(mcgp, oracle):
    """BASELINE config 2 at reduced size (the full 10 M sims: tests/replay_config2.py, result in profiles/): chunk c is
    the reference's run_monte_carlo(10 000, seed=42+c); chunk 0 is config 1, whose count table is a golden fixture."""
    import replay_config2
    r = replay_config2.run(chunks=12, chunk_sims=10000, threads=8)
    assert r["sims"] == 120000
    assert r["order_mismatches"] == 0 and r["time_mismatches"] == 0 and r["draws_mismatches"] == 0, r
