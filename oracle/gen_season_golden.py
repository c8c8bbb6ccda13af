"""TEST INFRASTRUCTURE -- fixture for BASELINE config 4 (24-race season, Brier path), from the UNMODIFIED reference.

    PYTHONHASHSEED=0 python oracle/gen_season_golden.py

Records into tests/golden/season.json:
  * winners[r]   : the winner of reference sim #0 of season race r run with seed=1000+r,
                   i.e. RaceSimulator(cfg_r).run_monte_carlo(1, ..., seed=1000+r) on the unmodified reference;
  * podiums[r]   : its top three;
  * scoring KATs : outputs of the reference's own brier_score / podium_accuracy / calibration_analysis
                   (src/validation.py:82-158, imported with a 5-line `fastf1` stub) on a fixed, synthetic
                   prediction set, so the product's scoring port can be checked without the reference tree.
"""
import importlib
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

import numpy as np  # noqa: E402

from oracle import ref_record  # noqa: E402

wl = importlib.import_module("monte-carlo-gp_b200.workloads")


def main():
    ref_sim = ref_record.import_reference()
    stub = types.ModuleType("fastf1")
    stub.Cache = types.SimpleNamespace(enable_cache=lambda *a, **k: None)
    sys.modules["fastf1"] = stub
    cwd = os.getcwd()
    os.chdir("/tmp")  # F1DataLoader() would mkdir ./cache; we never instantiate it, but stay out of the repo anyway
    import src.validation as ref_val
    os.chdir(cwd)

    winners, podiums = [], []
    for r in range(wl.N_SEASON_RACES):
        cfg, mc = wl.workload(f"season:{r}")
        sim = ref_sim.RaceSimulator(ref_sim.RaceConfig(**cfg))
        res = sim.run_monte_carlo(1, mc["grid_probs"], mc["base_pace"], mc["tire_deg"], mc["driver_variance"],
                                  mc["driver_dnf_rates"], seed=1000 + r, track_condition=mc["track_condition"])
        order = sorted(((min(v), str(d)) for d, v in res.items()))
        winners.append(order[0][1])
        podiums.append([d for _, d in order[:3]])

    # scoring KATs on a fixed synthetic prediction set
    rs = np.random.RandomState(7)
    D = list(wl.DRIVER_TEAMS)
    preds, acts = [], []
    for r in range(12):
        w = rs.dirichlet(np.ones(len(D)) * 0.3)
        p = rs.dirichlet(np.ones(len(D)) * 0.5)
        pod = np.minimum(1.0, 3 * p)
        win_probs = {d: float(x) for d, x in zip(D, w)}
        preds.append({"win_probabilities": win_probs, "podium_probabilities": {d: float(x) for d, x in zip(D, pod)}})
        top = sorted(D, key=lambda d: -win_probs[d])
        acts.append({"winner": top[r % 3], "podium": [top[(r + k) % 5] for k in range(3)]})
    preds.append({"win_probabilities": {}, "podium_probabilities": {}})        # skipped by the reference
    acts.append({"winner": None, "podium": []})
    kat = {
        "predictions": preds, "actuals": acts,
        "win_brier": float(ref_val.brier_score([p["win_probabilities"] for p in preds], [a["winner"] for a in acts])),
        "podium_accuracy": float(ref_val.podium_accuracy(preds, acts)),
        "calibration": ref_val.calibration_analysis(preds, acts),
        "empty_brier": float(ref_val.brier_score([], [])),
    }
    out = {"winners": winners, "podiums": podiums, "kat": kat, "pop": ref_record.reference_pop_choices(),
           "hashseed": os.environ.get("PYTHONHASHSEED")}
    with open(os.path.join(ROOT, "tests", "golden", "season.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("winners:", winners)
    print("win_brier KAT:", kat["win_brier"], "podium_acc:", kat["podium_accuracy"])


if __name__ == "__main__":
    main()
