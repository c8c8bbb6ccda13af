"""Golden parity cases: named (RaceConfig kwargs, run_monte_carlo kwargs, seed, n_sims) tuples.

Used by ``oracle/gen_golden.py`` (which runs the *unmodified reference* from /root/reference on
them, in this container) and by the tests (which run the C oracle and the CUDA path on the same
inputs and compare against the committed ``tests/golden/<case>.npz``).  The reference has no tests
or golden vectors of its own, so these are the pin.

Every case is a deterministic function of constants -- nothing here reads /root/reference.
"""
from __future__ import annotations

import copy
import importlib
import math

wl = importlib.import_module("monte-carlo-gp_b200.workloads")


def _bahrain():
    return wl.workload("bahrain")


def case_bahrain_dry():
    """BASELINE config 1: Bahrain 57 laps, 20 drivers, seed 42, 10 000 sims."""
    cfg, mc = _bahrain()
    return cfg, mc, 42, 10000


def case_monaco_sc():
    """BASELINE config 3 inputs: Monaco 78 laps, high SC rate (fuel clamp Q11, long stints)."""
    cfg, mc = wl.workload("monaco_sc")
    return cfg, mc, 42, 1500


def case_sprint19():
    """19-lap race: exercises the `{'MEDIUM','HARD'}.pop()` path B."""
    cfg, mc = wl.workload("sprint19")
    return cfg, mc, 42, 2000


def case_damp():
    cfg, mc = _bahrain()
    mc["track_condition"] = "damp"
    return cfg, mc, 7, 800


def case_wet():
    cfg, mc = _bahrain()
    mc["track_condition"] = "wet"
    return cfg, mc, 8, 800


def case_onehot():
    """quali/sprint prediction point: one-hot grid (reference src/predictor.py:189-205), variance x0.9."""
    cfg, mc = wl.workload("point:quali")
    return cfg, mc, 11, 1000


def case_npfloat_grid():
    """grid_probs rows made of np.float64 items as src/predictor.py:367-372 produces them:
    CPython's sum() then takes the plain left-to-right path instead of Neumaier."""
    import numpy as np
    cfg, mc = _bahrain()
    D = list(mc["grid_probs"])
    n = len(D)
    gp = {}
    for k, d in enumerate(D):
        expected = (1 - min(0.999, max(0.001, 0.3 * math.exp(-0.35 * k)))) * n
        probs = [np.exp(-((pos - expected) ** 2) / (2 * 5.0 ** 2)) for pos in range(n)]
        total = sum(probs)
        gp[d] = [p / total for p in probs]
    mc["grid_probs"] = gp
    return cfg, mc, 5, 1000


def case_defaults():
    """.get() defaults: 22 drivers, two of them unknown to every dict, missing compound
    table entries, missing per-driver entries, driver_dnf_rates=None, short grid_probs rows."""
    cfg, mc = _bahrain()
    D = list(mc["grid_probs"]) + ["XXA", "XXB"]
    mc["grid_probs"] = wl.gaussian_grid_probs(D, spread=3.0)
    mc["grid_probs"]["XXB"] = mc["grid_probs"]["XXB"][:15]       # short row -> bounds check :120
    mc["grid_probs"]["HAM"] = [0.0] * len(D)                      # never chosen until the uniform fallback
    for d in ("NOR", "XXA", "BEA"):
        mc["base_pace"].pop(d, None)
    for d in ("VER", "XXB", "HUL", "SAI"):
        mc["tire_deg"].pop(d, None)
    for d in ("PIA", "XXA"):
        mc["driver_variance"].pop(d, None)
    mc["driver_dnf_rates"] = None
    mc["tire_deg"]["LEC"] = 0.0        # driver_factor falls back to 1.0 (:321); pit class '<0.02'
    mc["tire_deg"]["ALO"] = -0.01
    del cfg["tire_compounds"]["HARD"]                               # -> deg 0.05 / delta 0 / optimal 30
    cfg["tire_compounds"]["MEDIUM"] = {"pace_delta": 0.1}           # partial entry
    del cfg["dnf_rates"]["Haas"]                                    # team default 0.002
    return cfg, mc, 3, 1000


def case_attrition():
    """Heavy attrition: lap-1 DNF 4x team rate up to 20 %, per-lap driver rate 1.5 %: exercises DNF
    classification (:231-242), DNF cars blocking overtake pairs (Q5), and all-cars-out races."""
    cfg, mc = _bahrain()
    cfg["dnf_rates"] = {t: 0.01 + 0.004 * i for i, t in enumerate(cfg["dnf_rates"])}
    mc["driver_dnf_rates"] = {d: 0.015 for d in mc["driver_dnf_rates"]}
    cfg["total_laps"] = 40
    return cfg, mc, 9, 1500


def case_wipeout():
    """Everybody retires quickly: handlers' `if not active: return` early exits (:348,:381,:407)."""
    cfg, mc = _bahrain()
    cfg["total_laps"] = 30
    cfg["sc_probability"], cfg["vsc_probability"], cfg["red_flag_probability"] = 0.2, 0.5, 0.1
    mc["driver_dnf_rates"] = {d: 0.35 for d in mc["driver_dnf_rates"]}
    return cfg, mc, 10, 400


def case_events():
    """Event storm: red 10 %, SC 30 %, VSC 30 % per lap; DRS-disable windows overlap."""
    cfg, mc = _bahrain()
    cfg["sc_probability"], cfg["vsc_probability"], cfg["red_flag_probability"] = 0.3, 0.3, 0.1
    return cfg, mc, 12, 1000


def case_tight():
    """Tight field and low overtake delta: long multi-pass overtake chains, frequent dirty air."""
    cfg, mc = _bahrain()
    cfg["overtake_delta"] = 0.05
    cfg["drs_delta"] = 0.4
    D = list(mc["base_pace"])
    mc["base_pace"] = {d: 90.0 + 0.01 * ((k * 7) % 20) for k, d in enumerate(D)}
    mc["tire_deg"] = {d: 0.03 + 0.004 * ((k * 3) % 10) for k, d in enumerate(D)}
    return cfg, mc, 13, 1000


def case_small_grids():
    """n = 3 drivers, 6 laps."""
    cfg, mc = _bahrain()
    D = ["VER", "NOR", "LEC"]
    cfg["total_laps"] = 6
    mc = {k: ({d: v[d] for d in D} if isinstance(v, dict) else v) for k, v in mc.items()}
    mc["grid_probs"] = wl.gaussian_grid_probs(D, spread=1.0)
    return cfg, mc, 14, 500


def case_single():
    """n = 1 driver, 3 laps, and total_laps edge."""
    cfg, mc = _bahrain()
    D = ["VER"]
    cfg["total_laps"] = 3
    mc = {k: ({d: v[d] for d in D} if isinstance(v, dict) else v) for k, v in mc.items()}
    mc["grid_probs"] = {"VER": [1.0]}
    return cfg, mc, 15, 50


def case_one_lap():
    """total_laps = 1: only the lap-1 path and the final classification run."""
    cfg, mc = _bahrain()
    cfg["total_laps"] = 1
    return cfg, mc, 16, 300


def case_canada70():
    cfg, mc = wl.workload("season:7")
    return cfg, mc, 1007, 600


CASES = {
    "bahrain_dry": case_bahrain_dry,
    "monaco_sc": case_monaco_sc,
    "sprint19": case_sprint19,
    "damp": case_damp,
    "wet": case_wet,
    "onehot": case_onehot,
    "npfloat_grid": case_npfloat_grid,
    "defaults": case_defaults,
    "attrition": case_attrition,
    "wipeout": case_wipeout,
    "events": case_events,
    "tight": case_tight,
    "small_grids": case_small_grids,
    "single": case_single,
    "one_lap": case_one_lap,
    "canada70": case_canada70,
}

# how many leading sims of each case keep their full per-sim record in the fixture
DETAIL_SIMS = 1000
# how many leading sims keep their raw draw tapes (U_py, Z, U_np) in the fixture
TAPE_SIMS = 8


def get_case(name: str):
    cfg, mc, seed, n = CASES[name]()
    return copy.deepcopy(cfg), copy.deepcopy(mc), seed, n
