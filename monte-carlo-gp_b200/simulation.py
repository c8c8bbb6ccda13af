"""Drop-in for the reference's ``src/simulation.py``: same ``RaceConfig`` / ``RaceSimulator`` entry points
(reference src/simulation.py:37-52, :55-100, :147-155), executed by the sm_100a kernels in libmcgp.so.

``src/predictor.py:8`` does ``from src.simulation import RaceSimulator, RaceConfig`` and then
``RaceSimulator(race_config).run_monte_carlo(n_simulations=..., grid_probs=..., base_pace=..., tire_deg=...,
driver_variance=..., driver_dnf_rates=..., track_condition=...)`` (:264, :283-291); the return value is
``{driver: {position(1-based): probability}}`` holding only non-zero cells (:97-100).  This module keeps exactly
that surface.  Host code here only marshals dicts into the dense ``mcgp_race_params`` block (applying the
reference's ``.get`` defaults) and turns count tables back into dicts; all simulation happens on the
GPU.  There is no CPU fallback.
"""
from __future__ import annotations

import os
import random
from dataclasses import dataclass, field
from itertools import chain

import numpy as np

from . import capi

__all__ = ["CarState", "RaceConfig", "RaceSimulator", "build_race_params", "counts_to_probabilities",
           "default_pop_choices", "run_batch"]


@dataclass
class CarState:
    """Mirror of the reference's per-car record (src/simulation.py:9-34); on the GPU this state lives in the
    registers of one lane.  Kept for API compatibility (nothing in the product path instantiates it)."""
    driver: str
    team: str
    position: int
    lap: int
    tire_compound: str
    tire_age: int
    fuel_load: float
    time_behind_leader: float
    pit_stops: int
    cumulative_time: float = 0.0
    drs_enabled: bool = False
    dnf: bool = False
    used_compounds: set = field(default_factory=set)
    laps_completed: int = 0
    last_lap_time: float = 0.0

    def __post_init__(self):
        self.used_compounds.add(self.tire_compound)


@dataclass
class RaceConfig:
    """Field-for-field the reference dataclass (src/simulation.py:37-52)."""
    total_laps: int
    pit_loss: float
    overtake_delta: float
    sc_probability: float
    vsc_probability: float
    red_flag_probability: float
    dnf_rates: dict[str, float]
    drs_zones: int
    drs_delta: float
    tire_compounds: dict[str, dict]
    driver_teams: dict[str, str]
    dirty_air_threshold: float = 2.0
    dirty_air_penalty: float = 0.5


def default_pop_choices() -> tuple[str, str]:
    """Which compound ``available.pop()`` yields on the two hash-dependent paths of the reference's
    two-compound rule (src/simulation.py:486,488).  The reference's outcome depends on
    PYTHONHASHSEED (and on how CPython built its set constant), i.e. it is a *model input*; we evaluate the
    reference's own expressions in this interpreter, and let MCGP_POP_NO_MEDIUM / MCGP_POP_NO_SOFT or the
    RaceSimulator arguments override them."""
    dry_compounds = {'SOFT', 'MEDIUM', 'HARD'}
    no_medium = (dry_compounds - {'MEDIUM'}).pop()
    no_soft = (dry_compounds - {'SOFT'}).pop()
    return os.environ.get("MCGP_POP_NO_MEDIUM", no_medium), os.environ.get("MCGP_POP_NO_SOFT", no_soft)


def _item_kind(x) -> int:
    """How CPython's builtin sum() treats this grid_probs item (src/simulation.py:123,133)."""
    if type(x) is float:
        return capi.ITEM_FLOAT
    if isinstance(x, (bool, int, np.integer)):
        return capi.ITEM_INT0 if x == 0 else capi.ITEM_FLOAT
    return capi.ITEM_NPFLOAT


def build_race_params(config: RaceConfig, grid_probs: dict, base_pace: dict, tire_deg: dict, driver_variance: dict,
                      driver_dnf_rates: dict | None = None, track_condition: str = 'dry',
                      pop_no_medium: str | None = None, pop_no_soft: str | None = None, stream: int = 0,
                      drivers: list | None = None) -> capi.McgpRaceParams:
    """Flatten the reference's call arguments into one ``mcgp_race_params`` block (include/mcgp.h).

    This runs on every product call (``run_monte_carlo(10 000)`` is 0.13 ms of GPU time), so the 20 x 20 grid rows go
    through numpy views of the struct instead of a Python loop per cell; rows of plain ``float`` / ``np.float64`` items
    (what callers pass: src/predictor.py:189-205, :367-372) take the fast path, anything else the per-item one."""
    D = list(grid_probs.keys()) if drivers is None else list(drivers)  # driver universe, src/simulation.py:107
    n = len(D)
    if not 1 <= n <= capi.MAX_DRIVERS:
        raise ValueError(f"this engine maps one driver per warp lane: 1..{capi.MAX_DRIVERS} drivers, got {n}")
    dflt_a, dflt_b = default_pop_choices()
    p = capi.McgpRaceParams()
    p.n_drivers, p.total_laps = n, int(config.total_laps)
    p.track_condition = capi.TRACK_CONDITIONS.get(track_condition, 0)  # anything else behaves as dry (:251-258)
    p.pop_no_medium = capi.COMPOUNDS.index(pop_no_medium or dflt_a)
    p.pop_no_soft = capi.COMPOUNDS.index(pop_no_soft or dflt_b)
    p.stream = int(stream)
    p.pit_loss, p.overtake_delta = config.pit_loss, config.overtake_delta
    p.sc_probability, p.vsc_probability = config.sc_probability, config.vsc_probability
    p.red_flag_probability, p.drs_delta = config.red_flag_probability, config.drs_delta
    p.dirty_air_threshold, p.dirty_air_penalty = config.dirty_air_threshold, config.dirty_air_penalty
    for k, name in enumerate(capi.COMPOUNDS):
        info = config.tire_compounds.get(name, {})           # :317, :454
        p.compound_pace_delta[k] = info.get('pace_delta', 0)  # :325
        p.compound_deg_rate[k] = info.get('deg_rate', 0.05)   # :320
        p.compound_optimal_laps[k] = info.get('optimal_laps', 30)  # :455
    driver_dnf_rates = driver_dnf_rates or {}                # :81, :161
    teams, rates = config.driver_teams, config.dnf_rates
    team_rate = [rates.get(teams.get(d, 'Unknown'), 0.002) for d in D]   # :263, :192, :286
    p.team_dnf_rate[:n] = team_rate
    p.base_pace[:n] = [base_pace.get(d, 90.0) for d in D]               # :202
    p.tire_deg[:n] = [tire_deg.get(d, 0.05) for d in D]                 # :203, :514
    p.tire_deg_pit[:n] = [tire_deg.get(d, 0.0) for d in D]              # :458
    p.driver_variance[:n] = [driver_variance.get(d, 0.15) for d in D]   # :204
    p.dnf_rate[:n] = [driver_dnf_rates.get(d, r) for d, r in zip(D, team_rate)]  # :190-193
    gp = np.frombuffer(p, np.float64, capi.MAX_DRIVERS ** 2, capi.McgpRaceParams.grid_probs.offset).reshape(capi.MAX_DRIVERS, -1)
    gk = np.frombuffer(p, np.uint8, capi.MAX_DRIVERS ** 2, capi.McgpRaceParams.grid_kind.offset).reshape(capi.MAX_DRIVERS, -1)
    rows = [grid_probs.get(d, ()) for d in D]                 # (views of the struct's memory above)
    kinds = set(map(type, chain.from_iterable(rows))) if all(len(r) == n for r in rows) else None
    if kinds == {float} or kinds == {np.float64}:             # the whole table at once
        gp[:n, :n] = rows
        gk[:n, :n] = capi.ITEM_FLOAT if kinds == {float} else capi.ITEM_NPFLOAT
    else:
        for i, row in enumerate(rows):
            m = min(len(row), n)                              # bounds check :120 (missing cells are int 0)
            if m:
                cells = row[:m]
                gp[i, :m] = [float(v) for v in cells]
                gk[i, :m] = [_item_kind(v) for v in cells]
    block = gp[:n, :n]
    if not block.min() >= 0:                                   # (a NaN anywhere makes the minimum NaN)
        if np.isnan(block).any():
            raise ValueError("probabilities contain NaN")          # np.random.choice (:137) would raise
        raise ValueError("probabilities are not non-negative")      # idem
    return p


def counts_to_probabilities(hist: np.ndarray, drivers: list, n_simulations: int) -> dict:
    """hist[driver, pos] -> {driver: {pos+1: count / n}} with only non-zero cells (src/simulation.py:97-100, Q9)."""
    n = len(drivers)
    rows, cols = np.nonzero(hist[:n])                          # row-major: grouped by driver, positions ascending
    probs = (hist[rows, cols] / n_simulations).tolist()        # uint64 / int in float64 == Python's int / int here
    pos1 = (cols + 1).tolist()
    ends = np.cumsum(np.bincount(rows, minlength=n)).tolist()
    out, a = {}, 0
    for d, b in zip(drivers, ends):
        if b > a:
            out[np.str_(d)] = dict(zip(pos1[a:b], probs[a:b]))
        a = b
    return out


class RaceSimulator:
    """Same constructor and public methods as the reference class (src/simulation.py:55-57, :59-69, :147-155).

    Extra keyword-only knobs (all optional): ``device`` (CUDA ordinal, default ``LOCAL_RANK`` or 0),
    ``pop_no_medium`` / ``pop_no_soft``, ``exact_normal`` (bit-reproducible normal generator).
    """

    def __init__(self, config: RaceConfig, *, device: int | None = None, pop_no_medium: str | None = None,
                 pop_no_soft: str | None = None, exact_normal: bool = False):
        self.config = config
        self.device = int(os.environ.get("LOCAL_RANK", "0")) if device is None else int(device)
        self.pop_no_medium, self.pop_no_soft = pop_no_medium, pop_no_soft
        self.flags = capi.F_EXACT_NORMAL if exact_normal else 0
        self.last_seed: int | None = None

    # -- helpers ------------------------------------------------------------------------------------
    def _engine(self) -> capi.Engine:
        return capi.get_engine(self.device)  # raises if libmcgp.so is missing or no B200 is visible

    def _params(self, grid_probs, base_pace, tire_deg, driver_variance, driver_dnf_rates, track_condition,
                drivers=None, stream=0):
        return build_race_params(self.config, grid_probs, base_pace, tire_deg, driver_variance, driver_dnf_rates,
                                 track_condition, self.pop_no_medium, self.pop_no_soft, stream, drivers)

    @staticmethod
    def _resolve_seed(seed) -> int:
        # seed=None continues the global `random` stream, so backtest_model's one-time random.seed(seed)
        # (src/validation.py:172-174) still makes a whole backtest reproducible.
        return random.getrandbits(64) if seed is None else int(seed) & (2 ** 64 - 1)

    # -- reference API ------------------------------------------------------------------------------
    def run_monte_carlo(self, n_simulations: int, grid_probs: dict[str, list[float]], base_pace: dict[str, float],
                        tire_deg: dict[str, float], driver_variance: dict[str, float],
                        driver_dnf_rates: dict[str, float] | None = None, seed: int | None = None,
                        track_condition: str = 'dry') -> dict[str, dict[int, float]]:
        """Run n simulations and return position probability distributions (src/simulation.py:59-100)."""
        hist = self.run_monte_carlo_counts(n_simulations, grid_probs, base_pace, tire_deg, driver_variance,
                                           driver_dnf_rates, seed, track_condition)
        if hist is None:
            return {}
        return counts_to_probabilities(hist, list(grid_probs.keys()), n_simulations)

    def run_monte_carlo_counts(self, n_simulations, grid_probs, base_pace, tire_deg, driver_variance,
                               driver_dnf_rates=None, seed=None, track_condition='dry', sim_begin: int = 0,
                               stream: int = 0):
        """The integer table behind run_monte_carlo: hist[driver, pos] (uint64), or None for an empty problem.
        `stream` selects an independent family of draws (races of one batch use their index)."""
        if n_simulations <= 0 or not grid_probs:  # reference: the loop body never runs / _sample_grid returns []
            return None
        params = self._params(grid_probs, base_pace, tire_deg, driver_variance, driver_dnf_rates, track_condition,
                              stream=stream)
        self.last_seed = self._resolve_seed(seed)
        hist = self._engine().run_native([params], int(n_simulations), sim_begin, self.last_seed, self.flags)
        return hist[0]

    def run_monte_carlo_by_lap(self, n_simulations: int, grid_probs, base_pace, tire_deg, driver_variance,
                               driver_dnf_rates=None, seed=None, track_condition='dry'):
        """Extension (absent upstream): run_monte_carlo plus the running-position distribution after EVERY lap, reduced
        on the GPU (no per-sim trace leaves the chip).  Returns ``(probabilities, by_lap)`` where ``probabilities`` is
        run_monte_carlo's dict and ``by_lap`` a float64 array [total_laps, n_drivers, n_drivers]:
        by_lap[lap-1, d, pos] = P(driver d runs in position pos+1 after lap `lap`); a row sums to P(d still running)."""
        if n_simulations <= 0 or not grid_probs:
            return {}, None
        params = self._params(grid_probs, base_pace, tire_deg, driver_variance, driver_dnf_rates, track_condition)
        self.last_seed = self._resolve_seed(seed)
        hist, laphist = self._engine().run_native_laphist([params], int(n_simulations), 0, self.last_seed, self.flags)
        return (counts_to_probabilities(hist[0], list(grid_probs.keys()), n_simulations),
                laphist[0].astype(np.float64) / n_simulations)

    def simulate_race(self, grid: list[str], base_pace: dict[str, float], tire_deg: dict[str, float],
                      driver_variance: dict[str, float], driver_dnf_rates: dict[str, float] | None = None,
                      track_condition: str = 'dry') -> list[tuple[str, int]]:
        """Simulate a single race from a given grid, returns [(driver, position)] (src/simulation.py:147-242)."""
        if not grid:
            return []
        n = len(grid)
        onehot = {d: [1.0 if pos == k else 0.0 for pos in range(n)] for k, d in enumerate(grid)}
        params = self._params(onehot, base_pace, tire_deg, driver_variance, driver_dnf_rates, track_condition,
                              drivers=list(grid))
        self.last_seed = self._resolve_seed(None)
        _, finish = self._engine().run_native([params], 1, 0, self.last_seed, self.flags, want_finish=True)
        return [(grid[int(d)], pos + 1) for pos, d in enumerate(finish[0, 0])]

    # -- replay mode (verification): consume the reference's own draws, bit-exact -----------------------
    def replay(self, grid_probs, base_pace, tire_deg, driver_variance, driver_dnf_rates=None, track_condition='dry',
               *, u_py, z, u_np, offsets, serial_grid: bool = False) -> dict:
        """FP64 replay of explicit draw tapes (see include/mcgp.h mcgp_run_replay).  Returns per-sim arrays
        (finish, times, dnf_lap, grid, used) and the count table.  serial_grid=True evaluates every _sample_grid
        position in the reference's serial operation order (include/mcgp.h: mcgp_replay_serial_grid); the default
        does so only where the kernel's parallel evaluation cannot certify the same selection -- same results."""
        params = self._params(grid_probs, base_pace, tire_deg, driver_variance, driver_dnf_rates, track_condition)
        eng = self._engine()
        eng.replay_serial_grid(serial_grid)
        try:
            return eng.run_replay(params, u_py, z, u_np, offsets)
        finally:
            eng.replay_serial_grid(False)


def run_batch(simulators_and_inputs: list[tuple[RaceSimulator, dict]], n_simulations: int, seed: int | None = None,
              device: int | None = None) -> list[dict]:
    """Several races in ONE launch (BASELINE config 4: a 24-race season).  Each item is
    ``(RaceSimulator, run_monte_carlo kwargs without n_simulations/seed)``; race r draws from stream r."""
    if not simulators_and_inputs:
        return []
    sims = [s for s, _ in simulators_and_inputs]
    dev = sims[0].device if device is None else device
    params = [s._params(kw['grid_probs'], kw['base_pace'], kw['tire_deg'], kw['driver_variance'],
                        kw.get('driver_dnf_rates'), kw.get('track_condition', 'dry'), stream=r)
              for r, (s, kw) in enumerate(simulators_and_inputs)]
    sd = RaceSimulator._resolve_seed(seed)
    hist = capi.get_engine(dev).run_native(params, int(n_simulations), 0, sd, sims[0].flags)
    return [counts_to_probabilities(hist[r], list(kw['grid_probs'].keys()), n_simulations)
            for r, (_, kw) in enumerate(simulators_and_inputs)]
