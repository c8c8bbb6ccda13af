"""The drop-in behind the reference's OWN call site (src/predictor.py:45-67, :255-291, :302-314).

`tests/golden/predictor_call.json` (generator: oracle/gen_predictor_call.py) holds what the unmodified
`F1Predictor.predict_weekend` hands to `RaceSimulator(race_config).run_monte_carlo(...)` in three scenarios (predicted
grid with penalties, actual grid at Monaco, unknown circuit on a damp track) and what it made of the result.

CPU (here, where /root/reference exists): the reference caller is run again with a recording simulator and must
produce exactly the committed call; the drop-in's marshaller turns a *reference* `RaceConfig` instance + the recorded
keyword arguments into the same dense block as the oracle's marshaller, field by field; the oracle reproduces the
recorded output bit for bit.  GPU: the drop-in `RaceSimulator` takes the recorded call and the three consumers of the
result (win, podium, full distributions) agree with the reference within 3 sigma.
"""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from stats_util import assert_agree_two_stage

HAVE_REF = os.path.isdir("/root/reference/src")
FIX = json.load(open(os.path.join(GOLDEN_DIR, "predictor_call.json")))
SCENARIOS = {s["name"]: s for s in FIX["scenarios"]}
CFG_FIELDS = ("total_laps", "pit_loss", "overtake_delta", "sc_probability", "vsc_probability", "red_flag_probability",
              "dnf_rates", "drs_zones", "drs_delta", "tire_compounds", "driver_teams", "dirty_air_threshold", "dirty_air_penalty")


def _typed_grid(call):
    """grid_probs with the item types the reference caller produced (float vs np.float64: SURVEY Q12)."""
    return {d: [float(v) if k == 1 else (0 if k == 0 else np.float64(v)) for v, k in zip(row, call["grid_kinds"][d])]
            for d, row in call["grid_probs"].items()}


def _kwargs(call):
    return dict(grid_probs=_typed_grid(call), base_pace=call["base_pace"], tire_deg=call["tire_deg"],
                driver_variance=call["driver_variance"], driver_dnf_rates=call["driver_dnf_rates"],
                track_condition=call["track_condition"])


def _struct_fields(p, names):
    return {n: np.ctypeslib.as_array(getattr(p, n)).copy() if hasattr(getattr(p, n), "_length_") else getattr(p, n) for n in names}


def _assert_same_block(ours, orc):
    """mcgp_race_params (product marshaller) vs orc_params (oracle marshaller), field by field."""
    pairs = [("n_drivers", "n_drivers"), ("total_laps", "total_laps"), ("track_condition", "track_condition"),
             ("pop_no_medium", "pop_no_medium"), ("pop_no_soft", "pop_no_soft"), ("pit_loss", "pit_loss"),
             ("overtake_delta", "overtake_delta"), ("sc_probability", "sc_p"), ("vsc_probability", "vsc_p"),
             ("red_flag_probability", "red_p"), ("drs_delta", "drs_delta"), ("dirty_air_threshold", "dirty_thr"),
             ("dirty_air_penalty", "dirty_pen"), ("compound_pace_delta", "compound_pace_delta"),
             ("compound_deg_rate", "compound_deg_rate"), ("compound_optimal_laps", "compound_optimal"),
             ("base_pace", "base_pace"), ("tire_deg", "tire_deg"), ("tire_deg_pit", "tire_deg_pit"),
             ("driver_variance", "variance"), ("dnf_rate", "dnf_rate"), ("team_dnf_rate", "team_rate"),
             ("grid_probs", "grid_probs"), ("grid_kind", "grid_kind")]
    for a, b in pairs:
        va, vb = getattr(ours, a), getattr(orc, b)
        if hasattr(va, "_length_"):
            va, vb = np.ctypeslib.as_array(va), np.ctypeslib.as_array(vb)
            assert va.tobytes() == vb.tobytes(), f"field {a} differs"
        else:
            assert va == vb, f"field {a}: {va} != {vb}"


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_recorded_call_marshals_like_the_oracle(oracle, name):
    import mcgp_b200
    S = mcgp_b200.simulation
    sc = SCENARIOS[name]
    cfg, call = sc["config"], sc["call"]
    assert sc["kwargs_keys"] == sorted(["n_simulations", "grid_probs", "base_pace", "tire_deg", "driver_variance",
                                        "driver_dnf_rates", "track_condition"])  # src/predictor.py:283-291
    kw = _kwargs(call)
    ours = S.build_race_params(S.RaceConfig(**cfg), **kw, pop_no_medium=FIX["pop_choices"][0], pop_no_soft=FIX["pop_choices"][1])
    mc = dict(kw)
    orc = oracle.make_params(cfg, mc, *FIX["pop_choices"])
    _assert_same_block(ours, orc)


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_oracle_reproduces_the_recorded_prediction(oracle, name):
    """The recorded output is `run_monte_carlo(..., seed=7)` of the reference: the oracle must give the same table,
    and the consumers of src/predictor.py:307-314 the same win / podium numbers."""
    sc = SCENARIOS[name]
    cfg, call, res = sc["config"], sc["call"], sc["result"]
    n = call["n_simulations"]
    hist = oracle.run_monte_carlo(cfg, _kwargs(call), n, FIX["seed_both_streams"], *FIX["pop_choices"])
    D = list(call["grid_probs"])
    for i, d in enumerate(D):
        cells = {str(p + 1): int(c) / n for p, c in enumerate(hist[i]) if c}
        assert cells == res["full_distributions"].get(d, {}), d
        assert res["win_probabilities"][d] == cells.get("1", 0)
        assert res["podium_probabilities"][d] == sum(cells.get(str(p), 0) for p in (1, 2, 3))


@pytest.mark.skipif(not HAVE_REF, reason="needs /root/reference (build box only)")
def test_reference_caller_still_produces_the_recorded_call(oracle):
    """Runs the unmodified predict_weekend with a recording simulator that does NOT simulate: config + kwargs must be the
    committed fixture (so the fixture cannot rot), and a reference RaceConfig INSTANCE goes through the drop-in's
    marshaller unchanged (the dataclass is duck-typed: src/predictor.py:55-67 builds it by keyword)."""
    import importlib.util
    import mcgp_b200
    spec = importlib.util.spec_from_file_location("gen_predictor_call", os.path.join(os.path.dirname(GOLDEN_DIR), "..", "oracle", "gen_predictor_call.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    P = gen.import_reference_predictor()
    from src.config import DRIVER_TEAMS
    D = list(DRIVER_TEAMS)
    calls = []

    class Recorder(P.RaceSimulator):
        def run_monte_carlo(self, **kwargs):
            calls.append((self.config, kwargs))
            return {}

    saved = P.RaceSimulator
    P.RaceSimulator = Recorder
    cwd = os.getcwd()
    try:
        import tempfile
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)
            pred = P.F1Predictor()
        os.chdir(cwd)
        pred.data_loader, pred.feature_engine = gen.FakeLoader(D), gen.FakeFeatures(D)
        out = pred.predict_weekend(2024, "Bahrain Grand Prix", prediction_point="fp2", grid_penalties={"HAM": 5, "ALO": "engine"})
    finally:
        os.chdir(cwd)
        P.RaceSimulator = saved
    (ref_config, kwargs), = calls
    sc = SCENARIOS["bahrain_fp2_penalty"]
    for f in CFG_FIELDS:
        assert getattr(ref_config, f) == sc["config"][f], f
    assert kwargs["n_simulations"] == 10000 and kwargs["track_condition"] == sc["call"]["track_condition"]
    for key in ("base_pace", "tire_deg", "driver_variance", "driver_dnf_rates"):
        assert {d: float(v) for d, v in kwargs[key].items()} == sc["call"][key], key
    assert {d: [float(x) for x in r] for d, r in kwargs["grid_probs"].items()} == sc["call"]["grid_probs"]
    # an empty result is what the consumers must survive too (:307-314 use .get(d, {}).get(1, 0))
    assert set(out["win_probabilities"].values()) == {0}
    # the reference's own RaceConfig instance + its own kwargs through the drop-in marshaller == the oracle's block
    S = mcgp_b200.simulation
    kw = {k: v for k, v in kwargs.items() if k != "n_simulations"}
    ours = S.build_race_params(ref_config, **kw, pop_no_medium=FIX["pop_choices"][0], pop_no_soft=FIX["pop_choices"][1])
    _assert_same_block(ours, oracle.make_params(sc["config"], kw, *FIX["pop_choices"]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_dropin_serves_the_reference_call_site(oracle, name):
    """GPU twin: the recorded call through the drop-in RaceSimulator; win / podium / every cell within 3 sigma of the
    reference (two-stage), the returned dict readable by the three consumers of src/predictor.py:307-314."""
    import mcgp_b200
    S = mcgp_b200.simulation
    sc = SCENARIOS[name]
    cfg, call = sc["config"], sc["call"]
    kw = _kwargs(call)
    sim = S.RaceSimulator(S.RaceConfig(**cfg), pop_no_medium=FIX["pop_choices"][0], pop_no_soft=FIX["pop_choices"][1])
    D = list(call["grid_probs"])
    probs = sim.run_monte_carlo(n_simulations=call["n_simulations"], **kw)          # the call of :283-291, verbatim
    win = {d: probs.get(d, {}).get(1, 0) for d in D}                                  # :307-309
    podium = {d: sum(probs.get(d, {}).get(p, 0) for p in [1, 2, 3]) for d in D}      # :310-313
    assert abs(sum(win.values()) - 1.0) < 1e-9 and abs(sum(podium.values()) - 3.0) < 1e-9
    assert all(isinstance(p, int) and 1 <= p <= len(D) for cells in probs.values() for p in cells)

    def gpu(n, stage):
        return sim.run_monte_carlo_counts(n, **kw, seed=500 + stage)

    def ref(n, stage):
        return oracle.run_monte_carlo(cfg, kw, n, 900 + stage, *FIX["pop_choices"], threads=8)

    print(name, assert_agree_two_stage(gpu, ref, 2_000_000, 200_000, name))
