#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- records golden vectors for monte-carlo-gp_b200/grid_model.py from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference): imports src/elo.py and src/predictor.py in place (fastf1,
absent here, is stubbed: the recorded functions never touch it) and writes tests/golden/grid_model.json with the
inputs and the reference's outputs as IEEE-754 hex strings.
    python oracle/gen_grid_golden.py
"""
import json, os, random, sys, types

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_functions():
    sys.dont_write_bytecode = True
    stub = types.ModuleType("fastf1")
    stub.Cache = types.SimpleNamespace(enable_cache=lambda *a, **k: None)
    sys.modules.setdefault("fastf1", stub)
    sys.path.insert(0, REF)
    try:
        from src.elo import F1EloSystem
        from src.predictor import F1Predictor
    finally:
        sys.path.remove(REF)

    def predict(drivers, ratings, features, penalties):
        elo = F1EloSystem()
        elo.ratings = {d: {"quali": r, "race": r} for d, r in ratings.items()}
        holder = types.SimpleNamespace(elo_system=elo)
        pole = elo.predict_quali_probs(drivers)
        rows = F1Predictor._predict_quali(holder, drivers, features)
        final = F1Predictor._adjust_for_penalties(holder, rows, penalties) if penalties else rows
        return pole, rows, final
    return predict


def cases():
    rnd = random.Random(20251018)
    sys.path.insert(0, ROOT)
    import mcgp_b200
    D = list(mcgp_b200.workloads.DRIVER_TEAMS)
    out = []
    out.append(dict(name="flat", drivers=D, ratings={}, features={}, penalties={}))
    out.append(dict(name="spread", drivers=D, ratings={d: 1500.0 + 40.0 * (10 - k) for k, d in enumerate(D)}, features={}, penalties={}))
    feats = {d: {"teammate_delta": rnd.uniform(-3, 3), "form_score": rnd.uniform(-1, 1), "circuit_affinity": rnd.uniform(-1, 1)}
             for d in D[::2]}
    feats[D[1]] = {"teammate_delta": 0, "form_score": 0.5}
    out.append(dict(name="features", drivers=D, ratings={d: rnd.gauss(1500, 120) for d in D[:17]}, features=feats, penalties={}))
    out.append(dict(name="penalties", drivers=D, ratings={d: rnd.gauss(1500, 80) for d in D}, features={},
                    penalties={D[0]: "engine", D[3]: 5, D[4]: "full_pu", D[7]: "gearbox", D[9]: "nonsense", D[11]: 19, D[12]: 0}))
    out.append(dict(name="small", drivers=D[:3], ratings={D[0]: 1700.0, D[1]: 1300.0}, features={D[2]: {"form_score": 1.0}},
                    penalties={D[0]: 1, D[1]: "pitlane_start"}))
    out.append(dict(name="extreme", drivers=D[:8], ratings={D[0]: 9000.0, D[1]: -4000.0, D[2]: 1500.0}, features={}, penalties={}))
    out.append(dict(name="one", drivers=D[:1], ratings={}, features={}, penalties={D[0]: 3}))
    for i in range(6):
        n = rnd.randint(2, 24)
        dr = [f"R{i}_{k}" for k in range(n)]
        out.append(dict(name=f"random{i}", drivers=dr, ratings={d: rnd.gauss(1500, 150) for d in dr if rnd.random() < 0.9},
                        features={d: {"teammate_delta": rnd.choice([0, rnd.uniform(-4, 4)]), "form_score": rnd.uniform(-1, 1),
                                      "circuit_affinity": rnd.uniform(-1, 1)} for d in dr if rnd.random() < 0.7},
                        penalties={d: rnd.choice(["engine", "gearbox", rnd.randint(0, 30)]) for d in dr if rnd.random() < 0.25}))
    return out


def hexrow(row):
    return [float(x).hex() for x in row]


if __name__ == "__main__":
    predict = reference_functions()
    golden = []
    for c in cases():
        pole, rows, final = predict(c["drivers"], c["ratings"], c["features"], c["penalties"])
        golden.append(dict(c, pole={d: float(v).hex() for d, v in pole.items()},
                           rows={d: hexrow(r) for d, r in rows.items()}, final={d: hexrow(r) for d, r in final.items()}))
    path = os.path.join(ROOT, "tests", "golden", "grid_model.json")
    with open(path, "w") as f:
        json.dump(golden, f, indent=0)
    print(path, len(golden), "cases", os.path.getsize(path), "bytes")

    # a synthetic season of qualifying / race events through the reference's F1EloSystem (src/elo.py)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, REF)
    from src.elo import F1EloSystem
    sys.path.remove(REF)
    import test_ratings
    import mcgp_b200
    D = list(mcgp_b200.workloads.DRIVER_TEAMS)
    events = test_ratings._events(random.Random(2025), D, 48)
    elo = F1EloSystem()
    ratings = test_ratings._apply(elo, events)
    pole = {d: float(v).hex() for d, v in elo.predict_quali_probs(D).items()}
    path = os.path.join(ROOT, "tests", "golden", "ratings.json")
    with open(path, "w") as f:
        json.dump(dict(drivers=D, events=events, ratings=ratings, pole=pole), f, indent=0)
    print(path, len(events), "events", os.path.getsize(path), "bytes")
