"""TEST INFRASTRUCTURE -- not part of the product path.

Tape recorder around the *unmodified* reference simulator imported in place from
/root/reference (never copied).  Only usable in the build container (the reference tree does
not exist on the GPU box); its outputs are committed as fixtures under tests/golden/ by
oracle/gen_golden.py.

What it records per simulated race (reference src/simulation.py:83-94 loop body):
  * the sampled grid, the finishing order, every car's final cumulative_time / dnf / dnf lap;
  * the cumulative number of `random.random()`, `np.random.normal`, `np.random.choice` calls;
  * optionally the raw draws themselves ("tapes"): U_py (random.random values), Z (standard
    normals: np.random.normal(loc, s) is recorded as loc + s*standard_normal(), which is
    bit-identical including NumPy's cached second Gaussian -- asserted in `self_check`) and
    U_np (the single random_sample() each np.random.choice consumes -- choice(p) is restated
    as searchsorted(cumsum(p)/cumsum(p)[-1], u, 'right'), also asserted in `self_check`).
"""
from __future__ import annotations

import random
import sys

import numpy as np

REFERENCE_ROOT = "/root/reference"


def import_reference():
    sys.dont_write_bytecode = True  # the reference tree is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import src.simulation as ref_sim  # noqa: E402  (reference src/simulation.py)
    return ref_sim


def reference_pop_choices() -> tuple[str, str]:
    """What `available.pop()` (reference src/simulation.py:486,488) yields in THIS process.

    The outcome depends on the collision layout of the reference's own `dry_compounds` set constant,
    which varies with PYTHONHASHSEED *and* with how CPython built the code object (fresh compile vs
    .pyc, constant merging) -- so evaluating a look-alike expression elsewhere is not reliable.
    We therefore probe the reference's own `_handle_pit_stops` with a crafted car on each path."""
    ref_sim = import_reference()
    cfg = ref_sim.RaceConfig(
        total_laps=60, pit_loss=20.0, overtake_delta=1.0, sc_probability=0.0, vsc_probability=0.0,
        red_flag_probability=0.0, dnf_rates={}, drs_zones=1, drs_delta=0.3,
        tire_compounds={c: {"pace_delta": 0.0, "deg_rate": 0.05, "optimal_laps": 10}
                        for c in ("SOFT", "MEDIUM", "HARD")}, driver_teams={})
    sim = ref_sim.RaceSimulator(cfg)

    def probe(compound: str, lap: int) -> str:
        car = ref_sim.CarState(driver="P", team="T", position=1, lap=lap, tire_compound=compound,
                               tire_age=50, fuel_load=50.0, time_behind_leader=0.0, pit_stops=0)
        sim._handle_pit_stops([car], lap, "dry", {})
        return car.tire_compound

    no_medium = probe("MEDIUM", 35)   # remaining 25: wants MEDIUM, must differ, {SOFT,HARD}.pop()  (path A)
    no_soft = probe("SOFT", 50)       # remaining 10: wants SOFT, must differ, {MEDIUM,HARD}.pop()  (path B)
    assert no_medium in ("SOFT", "HARD") and no_soft in ("MEDIUM", "HARD")
    return no_medium, no_soft


class Recorder:
    """Monkey-patches the three global draw functions the reference uses and logs every draw."""

    def __init__(self, keep_tapes: bool):
        self.keep = keep_tapes
        self.n_py = self.n_z = self.n_np = 0
        self.u_py: list[float] = []
        self.z: list[float] = []
        self.u_np: list[float] = []
        self._orig = None

    def __enter__(self):
        self._orig = (random.random, np.random.normal, np.random.choice)
        orig_random = random.random
        rec = self

        def rec_random():
            v = orig_random()
            rec.n_py += 1
            if rec.keep:
                rec.u_py.append(v)
            return v

        def rec_normal(loc=0.0, scale=1.0, size=None):
            assert size is None
            z = np.random.standard_normal()
            rec.n_z += 1
            if rec.keep:
                rec.z.append(float(z))
            return loc + scale * z

        def rec_choice(a, size=None, replace=True, p=None):
            assert size is None and replace and p is not None
            pa = np.array(p, dtype=np.float64)
            # same validation order as numpy/random/mtrand.pyx RandomState.choice
            if np.isnan(pa.sum()):
                raise ValueError("probabilities contain NaN")
            if (pa < 0).any():
                raise ValueError("probabilities are not non-negative")
            cdf = pa.cumsum()
            cdf /= cdf[-1]
            u = np.random.random_sample()
            rec.n_np += 1
            if rec.keep:
                rec.u_np.append(float(u))
            return np.asarray(a)[int(cdf.searchsorted(u, side="right"))]

        random.random, np.random.normal, np.random.choice = rec_random, rec_normal, rec_choice
        return self

    def __exit__(self, *exc):
        random.random, np.random.normal, np.random.choice = self._orig
        return False


def self_check(seed: int = 123, n: int = 4000) -> None:
    """The recorder's restatements of normal()/choice() are bit-identical to NumPy's own."""
    rng = np.random.RandomState(99)
    scales = rng.uniform(0.05, 1.5, n)
    ps = rng.dirichlet(np.ones(20), n)
    ps[:, 3] = 0.0
    ps /= ps.sum(1, keepdims=True)
    random.seed(seed); np.random.seed(seed)
    want = []
    for i in range(n):
        want.append(float(np.random.normal(0, scales[i])))
        if i % 3 == 0:
            want.append(int(np.random.choice(20, p=ps[i])))
        if i % 5 == 0:
            want.append(random.random())
    random.seed(seed); np.random.seed(seed)
    got = []
    with Recorder(False):
        for i in range(n):
            got.append(float(np.random.normal(0, scales[i])))
            if i % 3 == 0:
                got.append(int(np.random.choice(np.arange(20), p=ps[i])))
            if i % 5 == 0:
                got.append(random.random())
    assert want == got, "recorder restatement diverges from NumPy"


def record(cfg_kwargs: dict, mc_kwargs: dict, seed, n_sims: int, tape_sims: int = 0,
           reseed: bool = True) -> dict:
    """Run the reference's run_monte_carlo loop body (src/simulation.py:76-94) n_sims times.

    Returns arrays indexed by *driver index* (= position of the driver in grid_probs' key order).
    """
    ref_sim = import_reference()

    class Capturing(ref_sim.RaceSimulator):
        # a recording hook only: delegates to the unmodified method and remembers the car list,
        # whose objects the reference mutates in place (so after simulate_race it holds final state)
        def _update_positions(self, cars, lap=3, drs_disabled=False):
            self._cars = cars
            return super()._update_positions(cars, lap=lap, drs_disabled=drs_disabled)

    sim = Capturing(ref_sim.RaceConfig(**cfg_kwargs))
    gp = mc_kwargs["grid_probs"]
    drivers = list(gp.keys())
    idx = {d: i for i, d in enumerate(drivers)}
    n = len(drivers)
    args = (mc_kwargs["base_pace"], mc_kwargs["tire_deg"], mc_kwargs["driver_variance"],
            mc_kwargs.get("driver_dnf_rates") or {}, mc_kwargs.get("track_condition", "dry"))

    grid = np.zeros((n_sims, n), np.uint8)
    finish = np.zeros((n_sims, n), np.uint8)
    times = np.zeros((n_sims, n), np.float64)
    dnf_lap = np.zeros((n_sims, n), np.int16)      # 0 = classified finisher
    draws = np.zeros((n_sims, 3), np.int64)
    hist = np.zeros((n, n), np.int64)
    tapes = []

    if reseed and seed is not None:                # src/simulation.py:76-78
        random.seed(seed)
        np.random.seed(seed)
    with Recorder(tape_sims > 0) as rec:
        for s in range(n_sims):
            if s == tape_sims:
                rec.keep = False
            m0 = (len(rec.u_py), len(rec.z), len(rec.u_np))
            g = sim._sample_grid(gp)                                   # :85
            res = sim.simulate_race(g, *args)                          # :88
            grid[s] = [idx[str(d)] for d in g]
            for d, pos in res:                                          # :93-94
                finish[s, pos - 1] = idx[str(d)]
                hist[idx[str(d)], pos - 1] += 1
            for car in sim._cars:
                i = idx[str(car.driver)]
                times[s, i] = car.cumulative_time
                dnf_lap[s, i] = car.lap if car.dnf else 0
            draws[s] = (rec.n_py, rec.n_z, rec.n_np)
            if s < tape_sims:
                tapes.append((np.array(rec.u_py[m0[0]:]), np.array(rec.z[m0[1]:]), np.array(rec.u_np[m0[2]:])))
    out = dict(grid=grid, finish=finish, times=times, dnf_lap=dnf_lap, draws=draws, hist=hist)
    if tapes:
        out["tape_upy"] = np.concatenate([t[0] for t in tapes])
        out["tape_z"] = np.concatenate([t[1] for t in tapes])
        out["tape_unp"] = np.concatenate([t[2] for t in tapes])
    return out
