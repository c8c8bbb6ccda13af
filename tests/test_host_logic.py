"""Host-side mirror of the reference interface: marshalling (.get defaults), output dict shape (Q9),
workload tables vs the reference's config module (when /root/reference is present)."""
import ctypes as C
import dataclasses
import os
import sys

import numpy as np
import pytest

import golden_cases as gc

REF = "/root/reference"


@pytest.fixture(scope="module")
def mcgp():
    import mcgp_b200
    return mcgp_b200


@pytest.mark.parametrize("case", sorted(gc.CASES))
def test_product_and_oracle_marshallers_agree(mcgp, oracle, case):
    """Two independently written marshallers (product: simulation.build_race_params, checker:
    oracle.pyoracle.make_params) must produce the same dense block from the reference-shaped arguments."""
    cfg, mc, seed, _ = gc.get_case(case)
    sim = mcgp.simulation.RaceSimulator(mcgp.simulation.RaceConfig(**cfg), pop_no_medium="HARD", pop_no_soft="HARD")
    p = sim._params(mc["grid_probs"], mc["base_pace"], mc["tire_deg"], mc["driver_variance"],
                    mc.get("driver_dnf_rates"), mc.get("track_condition", "dry"))
    o = oracle.make_params(cfg, mc, "HARD", "HARD")
    n = p.n_drivers
    assert (p.n_drivers, p.total_laps, p.track_condition, p.pop_no_medium, p.pop_no_soft) == \
           (o.n_drivers, o.total_laps, o.track_condition, o.pop_no_medium, o.pop_no_soft)
    pairs = [("pit_loss", "pit_loss"), ("overtake_delta", "overtake_delta"), ("sc_probability", "sc_p"),
             ("vsc_probability", "vsc_p"), ("red_flag_probability", "red_p"), ("drs_delta", "drs_delta"),
             ("dirty_air_threshold", "dirty_thr"), ("dirty_air_penalty", "dirty_pen")]
    for a, b in pairs:
        assert getattr(p, a) == getattr(o, b), a
    for a, b in [("compound_pace_delta", "compound_pace_delta"), ("compound_deg_rate", "compound_deg_rate"),
                 ("compound_optimal_laps", "compound_optimal")]:
        assert list(getattr(p, a)) == list(getattr(o, b)), a
    for a, b in [("base_pace", "base_pace"), ("tire_deg", "tire_deg"), ("tire_deg_pit", "tire_deg_pit"),
                 ("driver_variance", "variance"), ("dnf_rate", "dnf_rate"), ("team_dnf_rate", "team_rate")]:
        assert list(getattr(p, a))[:n] == list(getattr(o, b))[:n], a
    for d in range(n):
        assert list(p.grid_probs[d])[:n] == list(o.grid_probs[d])[:n]
        assert list(p.grid_kind[d])[:n] == list(o.grid_kind[d])[:n]


def test_get_defaults(mcgp):
    """pace 90.0, deg 0.05 (lap time / overtakes) but 0.0 (pit window), variance 0.15, team 'Unknown'
    -> 0.002, missing compound -> deg 0.05 / delta 0 / optimal 30."""
    S = mcgp.simulation
    cfg = S.RaceConfig(total_laps=10, pit_loss=20.0, overtake_delta=0.5, sc_probability=0.0, vsc_probability=0.0,
                       red_flag_probability=0.0, dnf_rates={"T": 0.01}, drs_zones=1, drs_delta=0.3,
                       tire_compounds={"SOFT": {"deg_rate": 0.1}}, driver_teams={"A": "T"})
    p = S.build_race_params(cfg, {"A": [0.5, 0.5], "B": [1]}, {}, {}, {}, None, "monsoon", "SOFT", "HARD")
    assert (p.base_pace[0], p.tire_deg[0], p.tire_deg_pit[0], p.driver_variance[0]) == (90.0, 0.05, 0.0, 0.15)
    assert (p.dnf_rate[0], p.team_dnf_rate[0], p.dnf_rate[1], p.team_dnf_rate[1]) == (0.01, 0.01, 0.002, 0.002)
    assert list(p.compound_deg_rate) == [0.1, 0.05, 0.05, 0.05, 0.05]
    assert list(p.compound_pace_delta) == [0.0] * 5 and list(p.compound_optimal_laps) == [30.0] * 5
    assert p.track_condition == 0                                  # unknown condition behaves as dry
    assert (p.dirty_air_threshold, p.dirty_air_penalty) == (2.0, 0.5)
    assert list(p.grid_kind[1])[:2] == [mcgp.capi.ITEM_FLOAT, mcgp.capi.ITEM_INT0]      # int 1, then pos >= len(row)
    assert (p.pop_no_medium, p.pop_no_soft) == (0, 2)
    q = S.build_race_params(cfg, {"A": [np.float64(0.5), 0.5]}, {}, {}, {}, {"A": 0.5})
    assert list(q.grid_kind[0])[:1] == [mcgp.capi.ITEM_NPFLOAT] and q.dnf_rate[0] == 0.5


def test_counts_to_probabilities(mcgp):
    hist = np.array([[3, 0, 1], [1, 3, 0], [0, 1, 3]], np.uint64)
    out = mcgp.simulation.counts_to_probabilities(hist, ["A", "B", "C"], 4)
    assert out == {"A": {1: 0.75, 3: 0.25}, "B": {1: 0.25, 2: 0.75}, "C": {2: 0.25, 3: 0.75}}
    assert all(isinstance(k, np.str_) for k in out)                 # Q9: the reference's keys come from np.random.choice
    assert mcgp.simulation.counts_to_probabilities(np.zeros((2, 2), np.uint64), ["A", "B"], 1) == {}


def test_pop_choice_knobs(mcgp, monkeypatch):
    a, b = mcgp.simulation.default_pop_choices()
    assert a in ("SOFT", "HARD") and b in ("MEDIUM", "HARD")
    monkeypatch.setenv("MCGP_POP_NO_MEDIUM", "HARD")
    monkeypatch.setenv("MCGP_POP_NO_SOFT", "HARD")
    assert mcgp.simulation.default_pop_choices() == ("HARD", "HARD")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")
def test_interface_and_tables_match_reference(mcgp):
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    try:
        import src.config as rc
        import src.simulation as rs
    finally:
        sys.path.remove(REF)
    wl = mcgp.workloads
    assert wl.DRIVER_TEAMS == rc.DRIVER_TEAMS and list(wl.DRIVER_TEAMS) == list(rc.DRIVER_TEAMS)
    assert wl.DEFAULT_DNF_RATES == rc.DEFAULT_DNF_RATES and wl.TIRE_COMPOUNDS == rc.TIRE_COMPOUNDS
    assert wl.CIRCUITS == rc.CIRCUITS and list(wl.CIRCUITS) == list(rc.CIRCUITS)
    ours = [(f.name, f.default) for f in dataclasses.fields(mcgp.simulation.RaceConfig)]
    theirs = [(f.name, f.default) for f in dataclasses.fields(rs.RaceConfig)]
    assert ours == theirs
    assert [f.name for f in dataclasses.fields(mcgp.simulation.CarState)] == [f.name for f in dataclasses.fields(rs.CarState)]
    import inspect
    for meth in ("run_monte_carlo", "simulate_race"):
        a = inspect.signature(getattr(mcgp.simulation.RaceSimulator, meth))
        b = inspect.signature(getattr(rs.RaceSimulator, meth))
        assert list(a.parameters) == list(b.parameters), meth
        assert [p.default for p in a.parameters.values()] == [p.default for p in b.parameters.values()], meth


def test_workloads_cover_baseline_configs(mcgp):
    wl = mcgp.workloads
    cfg, mc = wl.workload("bahrain")
    assert cfg["total_laps"] == 57 and len(mc["grid_probs"]) == 20
    assert abs(sum(mc["grid_probs"]["VER"]) - 1.0) < 1e-12
    cfg, mc = wl.workload("monaco_sc")
    assert (cfg["total_laps"], cfg["sc_probability"], cfg["red_flag_probability"]) == (78, 0.05, 0.005)
    assert wl.N_SEASON_RACES == 24 and wl.workload("season:23")[0]["total_laps"] == 58
    cfg, mc = wl.workload("point:quali")
    assert mc["grid_probs"]["LAW"][1] == 1.0 and abs(mc["driver_variance"]["VER"] - 0.12 * 0.9) < 1e-15
    assert max(wl.workload("point:fp1")[1]["driver_variance"].values()) <= 0.3
