"""Grid-probability producer against the reference: bit-exact on the committed golden vectors
(recorded from the unmodified reference by oracle/gen_grid_golden.py), bit-exact against the reference itself on random
inputs when /root/reference is present, and usable as `grid_probs` of the drop-in simulator."""
import json
import os
import random

import numpy as np
import pytest

import mcgp_b200
from conftest import GOLDEN_DIR

gm = mcgp_b200.grid_model


def _hex(row):
    return [float(x).hex() for x in row]


def _golden():
    with open(os.path.join(GOLDEN_DIR, "grid_model.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("case", _golden(), ids=lambda c: c["name"])
def test_golden_vectors_bit_exact(case):
    D = case["drivers"]
    pole = gm.pole_probabilities(case["ratings"], D)
    assert {d: float(v).hex() for d, v in pole.items()} == case["pole"]
    rows = gm.quali_distributions(D, pole, case["features"])
    assert list(rows) == list(case["rows"]) and {d: _hex(r) for d, r in rows.items()} == case["rows"]
    final = gm.grid_probabilities(D, case["ratings"], case["features"], case["penalties"])
    assert {d: _hex(r) for d, r in final.items()} == case["final"]
    for d, r in final.items():
        assert len(r) == len(D) and abs(sum(r) - 1.0) < 1e-9 and min(r) >= 0.0


def test_random_inputs_equal_the_reference():
    if not os.path.isdir("/root/reference"):
        pytest.skip("reference tree not present on this box")
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_grid_golden", os.path.join(os.path.dirname(GOLDEN_DIR), "..", "oracle", "gen_grid_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    predict = gen.reference_functions()
    rnd = random.Random(7)
    for trial in range(60):
        n = rnd.randint(1, 26)
        D = [f"X{k}" for k in range(n)]
        ratings = {d: rnd.gauss(1500, 200) for d in D if rnd.random() < 0.85}
        feats = {d: {"teammate_delta": rnd.choice([0, rnd.uniform(-5, 5)]), "form_score": rnd.uniform(-1.5, 1.5),
                     "circuit_affinity": rnd.uniform(-1.5, 1.5)} for d in D if rnd.random() < 0.6}
        pens = {d: rnd.choice(["engine", "full_pu", "gearbox", "pitlane_start", "??", rnd.randint(-2, 30)]) for d in D if rnd.random() < 0.3}
        pole, rows, final = predict(D, ratings, feats, pens)
        ours = gm.grid_probabilities(D, ratings, feats, pens)
        assert {d: _hex(r) for d, r in ours.items()} == {d: _hex(r) for d, r in final.items()}, trial
        assert {d: float(v).hex() for d, v in gm.pole_probabilities(ratings, D).items()} == {d: float(v).hex() for d, v in pole.items()}


def test_nested_rating_layout_and_edge_cases():
    D = ["A", "B", "C"]
    flat = gm.pole_probabilities({"A": 1600.0, "B": 1400.0}, D)
    nested = gm.pole_probabilities({"A": {"quali": 1600.0, "race": 1.0}, "B": {"quali": 1400.0}}, D)
    assert flat == nested and abs(sum(flat.values()) - 1.0) < 1e-12 and flat["A"] > flat["C"] > flat["B"]
    assert gm.pole_probabilities({}, []) == {} and gm.quali_distributions([], {}) == {} and gm.grid_probabilities([], {}) == {}
    rows = gm.grid_probabilities(D, {}, penalties={"A": "full_pu", "B": 1})
    assert rows["A"] == [0.0, 0.0, 1.0] and rows["B"][0] == 0.0 and abs(sum(rows["B"]) - 1.0) < 1e-12
    assert all(isinstance(x, np.floating) for x in rows["C"])          # np.float64 items, as upstream (Q12)


def test_rows_feed_the_simulator_parameter_block():
    """The producer's rows are valid `grid_probs`: np.float64 items are marshalled as the plain-summation kind."""
    cfg, mc = mcgp_b200.workloads.workload("bahrain")
    D = list(mc["grid_probs"])
    gp = gm.grid_probabilities(D, {d: 1500.0 + 30.0 * (10 - k) for k, d in enumerate(D)}, penalties={D[0]: "gearbox"})
    sim = mcgp_b200.simulation.RaceSimulator(mcgp_b200.simulation.RaceConfig(**cfg), pop_no_medium="SOFT", pop_no_soft="MEDIUM")
    p = sim._params(gp, mc["base_pace"], mc["tire_deg"], mc["driver_variance"], mc["driver_dnf_rates"], "dry")
    assert p.n_drivers == 20
    assert p.grid_probs[0][0] == 0.0 and abs(sum(p.grid_probs[0][i] for i in range(20)) - 1.0) < 1e-12
    assert {p.grid_kind[1][i] for i in range(20)} == {mcgp_b200.simulation._item_kind(np.float64(0.5))}
