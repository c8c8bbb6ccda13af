#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- records what the reference's OWN caller hands to the race simulator.

Runs the unmodified `F1Predictor.predict_weekend` of /root/reference (src/predictor.py:99-319) on synthetic session
data: `fastf1` is stubbed (5 lines), the data loader and the feature engine -- the two FastF1-bound
collaborators -- are replaced by deterministic fakes, everything else (Elo warm-up :130-157, `_predict_quali` :321-375,
`_adjust_for_penalties` :377-407, the pandas extractors :409-569, `_create_race_config` :45-67, the variance / DNF / pace
adjustments :235-281) is the reference's code.  `src.predictor.RaceSimulator` is swapped for a recording subclass of the
reference's simulator that (a) stores the `RaceConfig` instance and the keyword arguments of the one call
`run_monte_carlo(n_simulations=10000, ...)` (:283-291) and (b) seeds both global streams with 7 before running it, so
the recorded output equals `run_monte_carlo(..., seed=7)`.

Output: tests/golden/predictor_call.json -- per scenario the call (config fields, kwargs with the CPython-sum() kind of
every grid_probs item) and what predict_weekend returned (win / podium / pole probabilities, full distributions).
The reference cannot travel to the GPU box; this fixture can.   usage: python oracle/gen_predictor_call.py
"""
import json
import os
import random
import sys
import tempfile
import types

import numpy as np
import pandas as pd

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "predictor_call.json")


def import_reference_predictor():
    """src.predictor with the fastf1 stub; returns the module."""
    sys.dont_write_bytecode = True
    if "fastf1" not in sys.modules:
        ff = types.ModuleType("fastf1")
        ff.Cache = type("Cache", (), {"enable_cache": staticmethod(lambda *a, **k: None)})
        ff.get_session = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("no network"))
        ff.get_event_schedule = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("no network"))
        sys.modules["fastf1"] = ff
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import src.predictor as P
    return P


class FakeLoader:
    """Deterministic stand-in for F1DataLoader (src/data_loader.py:8-156): synthetic results and practice laps."""

    def __init__(self, drivers, rainfall=False):
        self.drivers, self.rainfall = list(drivers), rainfall

    def load_season_data(self, season):
        rng = np.random.default_rng(season)
        out = {"qualifying": [], "races": [], "sprints": [], "sprint_qualifying": []}
        for r in range(6):
            skill = {d: 90.0 + 0.06 * k + rng.normal(0, 0.15) for k, d in enumerate(self.drivers)}
            out["qualifying"].append(sorted(skill.items(), key=lambda kv: kv[1]))
            order = sorted(self.drivers, key=lambda d: skill[d] + rng.normal(0, 0.2))
            out["races"].append([(d, i + 1) for i, d in enumerate(order)])
        return out

    def load_session(self, season, race, session):
        rng = np.random.default_rng(1234 + season + sum(map(ord, session)))
        rows = []
        for k, d in enumerate(self.drivers):
            comp = ("SOFT", "MEDIUM", "HARD")[k % 3]
            for lap in range(1, 13):
                t = 93.0 + 0.07 * k + 0.02 * k % 3 + (0.03 + 0.004 * k) * lap + {"SOFT": -0.55, "MEDIUM": 0.0, "HARD": 0.45}[comp] + rng.normal(0, 0.08)
                rows.append({"Driver": d, "LapNumber": lap, "LapTime": pd.Timedelta(seconds=float(t)), "Compound": comp,
                             "PitInTime": pd.NaT, "PitOutTime": pd.NaT})
        return pd.DataFrame(rows)

    def get_weather(self, season, race, session):
        return {"air_temp": 25.0, "track_temp": 35.0, "humidity": 50.0, "rainfall": self.rainfall, "wind_speed": 2.0}


class FakeFeatures:
    """Deterministic stand-in for F1FeatureEngine (src/features.py:10-786): only the scalars that reach the simulator."""

    def __init__(self, drivers):
        self.idx = {d: k for k, d in enumerate(drivers)}

    def load_historical_data(self, seasons):
        pass

    def calculate_quali_features(self, driver, race):
        k = self.idx[driver]
        return {"teammate_delta": 0.2 * ((k % 4) - 1.5), "form_score": 0.3 * (((7 * k) % 5) - 2) / 2, "circuit_affinity": 0.1 * ((k % 3) - 1)}

    def calculate_race_features(self, driver, race, weather):
        k = self.idx[driver]
        return {"clutch_factor": 0.25 * ((k % 5) - 2), "dnf_probability": 0.03 + 0.004 * (k % 6), "team_trend": 0.1 * ((k % 7) - 3) / 3,
                "wet_performance": 0.2 * ((k % 4) - 1.5)}


def _kind(x):
    if type(x) is float:
        return 1
    if isinstance(x, (int, np.integer)) and not isinstance(x, bool) and x == 0:
        return 0
    return 2


def record_scenario(P, name, drivers, **kw):
    calls = []

    class Recorder(P.RaceSimulator):
        def run_monte_carlo(self, **kwargs):
            calls.append((self.config, kwargs))
            random.seed(7)
            np.random.seed(7)
            return super().run_monte_carlo(**kwargs)

    saved = P.RaceSimulator
    P.RaceSimulator = Recorder
    try:
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as tmp:   # F1DataLoader() does mkdir ./cache (src/data_loader.py:10-11)
            os.chdir(tmp)
            try:
                pred = P.F1Predictor()
            finally:
                os.chdir(cwd)
        pred.data_loader = FakeLoader(drivers, rainfall=kw.pop("rainfall", False))
        pred.feature_engine = FakeFeatures(drivers)
        result = pred.predict_weekend(2024, kw.pop("race"), **kw)
    finally:
        P.RaceSimulator = saved
    (config, kwargs), = calls
    cfg = {f: getattr(config, f) for f in ("total_laps", "pit_loss", "overtake_delta", "sc_probability", "vsc_probability",
                                           "red_flag_probability", "dnf_rates", "drs_zones", "drs_delta", "tire_compounds",
                                           "driver_teams", "dirty_air_threshold", "dirty_air_penalty")}
    gp = kwargs["grid_probs"]
    call = {
        "n_simulations": kwargs["n_simulations"], "track_condition": kwargs["track_condition"], "drivers": [str(d) for d in gp],
        "grid_probs": {d: [float(x) for x in row] for d, row in gp.items()},
        "grid_kinds": {d: [_kind(x) for x in row] for d, row in gp.items()},
        "base_pace": {d: float(v) for d, v in kwargs["base_pace"].items()},
        "tire_deg": {d: float(v) for d, v in kwargs["tire_deg"].items()},
        "driver_variance": {d: float(v) for d, v in kwargs["driver_variance"].items()},
        "driver_dnf_rates": {d: float(v) for d, v in kwargs["driver_dnf_rates"].items()},
    }
    out = {"win_probabilities": {str(d): float(v) for d, v in result["win_probabilities"].items()},
           "podium_probabilities": {str(d): float(v) for d, v in result["podium_probabilities"].items()},
           "pole_probabilities": {str(d): float(v) for d, v in result["pole_probabilities"].items()},
           "full_distributions": {str(d): {str(p): float(v) for p, v in cells.items()} for d, cells in result["full_distributions"].items()},
           "prediction_point": result["prediction_point"], "grid_is_actual": result["grid_is_actual"]}
    return {"name": name, "config": cfg, "call": call, "result": out, "kwargs_keys": sorted(kwargs)}


def scenarios(P):
    from src.config import DRIVER_TEAMS
    D = list(DRIVER_TEAMS)
    return [
        record_scenario(P, "bahrain_fp2_penalty", D, race="Bahrain Grand Prix", prediction_point="fp2", grid_penalties={"HAM": 5, "ALO": "engine"}),
        record_scenario(P, "monaco_quali_actual_grid", D, race="Monaco Grand Prix", prediction_point="quali",
                        actual_grid={d: ((3 * k + 1) % 20) + 1 for k, d in enumerate(D)}),
        record_scenario(P, "unknown_circuit_damp_fp1", D, race="Atlantis Grand Prix", prediction_point="fp1", rainfall=True),
    ]


def main():
    P = import_reference_predictor()
    sys.path.insert(0, ROOT)
    from oracle import ref_record
    pops = ref_record.reference_pop_choices()      # what `available.pop()` yields in THIS process
    data = {"generator": "oracle/gen_predictor_call.py", "python": sys.version.split()[0], "numpy": np.__version__,
            "pythonhashseed": os.environ.get("PYTHONHASHSEED"), "pop_choices": list(pops), "seed_both_streams": 7,
            "scenarios": scenarios(P)}
    with open(OUT, "w") as f:
        json.dump(data, f, indent=1)   # (no sort_keys: the key order of grid_probs IS the driver order, src/simulation.py:107)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", [s["name"] for s in data["scenarios"]])


if __name__ == "__main__":
    main()
