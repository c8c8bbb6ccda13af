"""Synthetic race workloads shared by tests, bench.py and the golden generator.

Nothing here simulates anything: it only builds the *inputs* of
``RaceSimulator.run_monte_carlo`` (reference ``src/simulation.py:59-69``) for the
BASELINE.json configs.  The reference's race-weekend cache is git-ignored and
there is no network, so every workload is a fixed, deterministic function of a
small set of constants.

The season tables below restate the *data* of the reference's
``src/config.py:7-78`` (2025 line-up, per-team DNF rates, tyre compounds,
circuits) in our own layout; ``tests/test_workloads.py`` checks them against the
reference module whenever ``/root/reference`` is present.
"""
from __future__ import annotations

import math
from typing import Any

# (driver, team) in the order of reference src/config.py:7-28 -- the order matters:
# it is the driver universe order of every synthetic grid_probs dict.
_LINEUP = (
    ("VER", "Red Bull"), ("LAW", "Red Bull"), ("NOR", "McLaren"), ("PIA", "McLaren"),
    ("LEC", "Ferrari"), ("HAM", "Ferrari"), ("RUS", "Mercedes"), ("ANT", "Mercedes"),
    ("ALO", "Aston Martin"), ("STR", "Aston Martin"), ("GAS", "Alpine"), ("DOO", "Alpine"),
    ("TSU", "Racing Bulls"), ("HAD", "Racing Bulls"), ("ALB", "Williams"), ("SAI", "Williams"),
    ("HUL", "Sauber"), ("BOR", "Sauber"), ("OCO", "Haas"), ("BEA", "Haas"),
)
DRIVER_TEAMS: dict[str, str] = dict(_LINEUP)

# per-lap team DNF rates, reference src/config.py:31-42
DEFAULT_DNF_RATES: dict[str, float] = {
    "Red Bull": 0.0015, "McLaren": 0.0012, "Ferrari": 0.0018, "Mercedes": 0.0010,
    "Aston Martin": 0.0020, "Alpine": 0.0025, "Racing Bulls": 0.0022, "Williams": 0.0025,
    "Sauber": 0.0028, "Haas": 0.0025,
}

# compound -> (pace_delta, deg_rate, optimal_laps), reference src/config.py:45-51
_COMPOUND_ROWS = (
    ("SOFT", -0.8, 0.08, 15), ("MEDIUM", 0.0, 0.05, 25), ("HARD", 0.6, 0.03, 40),
    ("INTERMEDIATE", 5.0, 0.02, 30), ("WET", 10.0, 0.01, 50),
)
TIRE_COMPOUNDS: dict[str, dict] = {
    name: {"pace_delta": pd, "deg_rate": dr, "optimal_laps": ol} for name, pd, dr, ol in _COMPOUND_ROWS
}

# circuit -> (laps, pit_loss, drs_zones, overtake_delta), reference src/config.py:54-78
_CIRCUIT_ROWS = (
    ("Bahrain", 57, 21.0, 3, 0.6), ("Saudi Arabia", 50, 20.0, 3, 0.7), ("Australia", 58, 22.0, 4, 0.5),
    ("Japan", 53, 23.0, 1, 1.0), ("China", 56, 22.0, 2, 0.6), ("Miami", 57, 21.0, 3, 0.7),
    ("Monaco", 78, 24.0, 1, 1.5), ("Canada", 70, 22.0, 2, 0.6), ("Spain", 66, 21.0, 2, 0.8),
    ("Austria", 71, 20.0, 3, 0.5), ("Great Britain", 52, 21.0, 2, 0.7), ("Hungary", 70, 22.0, 1, 1.2),
    ("Belgium", 44, 23.0, 2, 0.5), ("Netherlands", 72, 20.0, 2, 1.0), ("Italy", 53, 26.0, 2, 0.4),
    ("Azerbaijan", 51, 24.0, 2, 0.5), ("Singapore", 62, 30.0, 3, 1.1), ("United States", 56, 21.0, 2, 0.7),
    ("Mexico", 71, 22.0, 3, 0.6), ("Brazil", 71, 21.0, 2, 0.5), ("Las Vegas", 50, 21.0, 2, 0.6),
    ("Qatar", 57, 21.0, 2, 0.8), ("Abu Dhabi", 58, 22.0, 2, 0.7),
)
CIRCUITS: dict[str, dict] = {
    name: {"laps": laps, "pit_loss": pl, "drs_zones": dz, "overtake_delta": od}
    for name, laps, pl, dz, od in _CIRCUIT_ROWS
}
# fallback circuit of reference src/predictor.py:38-43
FALLBACK_CIRCUIT = {"laps": 58, "pit_loss": 22.0, "drs_zones": 2, "overtake_delta": 0.8}

# event probabilities hard-coded at the reference call site, src/predictor.py:59-61
PRODUCT_EVENT_PROBS = {"sc_probability": 0.01, "vsc_probability": 0.015, "red_flag_probability": 0.002}
# BASELINE config 3 "high safety-car rate" 
HIGH_SC_EVENT_PROBS = {"sc_probability": 0.05, "vsc_probability": 0.03, "red_flag_probability": 0.005}

# prediction-point variance multipliers, reference src/predictor.py:241-247
UNCERTAINTY_MULTIPLIER = {"fp1": 1.5, "fp2": 1.2, "fp3": 1.0, "quali": 0.9, "sprint": 0.85}


def race_config_kwargs(circuit: dict, events: dict | None = None, **overrides: Any) -> dict:
    """Keyword arguments for ``RaceConfig`` exactly as reference src/predictor.py:55-67 builds them."""
    kw = dict(
        total_laps=circuit["laps"], pit_loss=circuit["pit_loss"], overtake_delta=circuit["overtake_delta"],
        dnf_rates=dict(DEFAULT_DNF_RATES), drs_zones=circuit["drs_zones"], drs_delta=0.3,
        tire_compounds={k: dict(v) for k, v in TIRE_COMPOUNDS.items()}, driver_teams=dict(DRIVER_TEAMS),
    )
    kw.update(events or PRODUCT_EVENT_PROBS)
    kw.update(overrides)
    return kw


def gaussian_grid_probs(drivers: list[str], spread: float = 2.5) -> dict[str, list[float]]:
    """grid_probs[D_k][pos] ∝ exp(-(pos-k)²/(2·spread²)), normalised with Python floats."""
    n = len(drivers)
    gp = {}
    for k, d in enumerate(drivers):
        w = [math.exp(-((pos - k) ** 2) / (2 * spread ** 2)) for pos in range(n)]
        t = sum(w)
        gp[d] = [x / t for x in w]
    return gp


def onehot_grid_probs(drivers: list[str], order: list[int] | None = None) -> dict[str, list[float]]:
    """Actual-grid one-hot distribution as reference src/predictor.py:189-205 builds for quali/sprint points."""
    n = len(drivers)
    order = list(range(n)) if order is None else order
    return {d: [1.0 if pos == order[k] else 0.0 for pos in range(n)] for k, d in enumerate(drivers)}


def common_inputs(total_laps: int, drivers: list[str] | None = None) -> dict:
    """The per-driver inputs every BASELINE config shares (SURVEY.md §8(d) 'common synthetic inputs')."""
    D = list(DRIVER_TEAMS) if drivers is None else drivers
    return dict(
        grid_probs=gaussian_grid_probs(D),
        base_pace={d: 92.0 + 0.07 * k for k, d in enumerate(D)},
        tire_deg={d: 0.015 + 0.003 * k for k, d in enumerate(D)},
        driver_variance={d: 0.12 + 0.005 * (k % 5) for k, d in enumerate(D)},
        driver_dnf_rates={d: 0.05 / total_laps for d in D},
        track_condition="dry",
    )


def workload(name: str, **opts: Any) -> tuple[dict, dict]:
    """Return ``(race_config_kwargs, run_monte_carlo_kwargs)`` for a named BASELINE workload.

    names: ``bahrain`` (configs 1/2), ``monaco_sc`` (config 3), ``sprint19`` (config 5 sprint
    race length), ``season:<r>`` (config 4, race r of 24), ``point:<fp1|fp2|fp3|quali|sprint>``
    (config 5 prediction points on the Bahrain race).
    """
    if name == "bahrain":
        c = CIRCUITS["Bahrain"]
        return race_config_kwargs(c), common_inputs(c["laps"])
    if name == "monaco_sc":
        c = CIRCUITS["Monaco"]
        return race_config_kwargs(c, HIGH_SC_EVENT_PROBS), common_inputs(c["laps"])
    if name == "sprint19":
        c = CIRCUITS["Bahrain"]
        return race_config_kwargs(c, total_laps=19), common_inputs(19)
    if name.startswith("season:"):
        r = int(name.split(":")[1])
        circuits = [CIRCUITS[k] for k in CIRCUITS] + [FALLBACK_CIRCUIT]
        c = circuits[r]
        cfg, mc = race_config_kwargs(c), common_inputs(c["laps"])
        D = list(mc["base_pace"])
        for k, d in enumerate(D):
            mc["base_pace"][d] = mc["base_pace"][d] + 0.05 * ((7 * k + 3 * r) % 11 - 5) / 5
        return cfg, mc
    if name.startswith("point:"):
        point = name.split(":")[1]
        c = CIRCUITS["Bahrain"]
        laps = int(opts.get("total_laps", c["laps"]))
        cfg, mc = race_config_kwargs(c, total_laps=laps), common_inputs(laps)
        mult = UNCERTAINTY_MULTIPLIER[point]
        mc["driver_variance"] = {d: min(0.3, v * mult) for d, v in mc["driver_variance"].items()}
        if point in ("quali", "sprint"):
            mc["grid_probs"] = onehot_grid_probs(list(mc["grid_probs"]))
        return cfg, mc
    raise KeyError(f"unknown workload {name!r}")


N_SEASON_RACES = len(_CIRCUIT_ROWS) + 1
