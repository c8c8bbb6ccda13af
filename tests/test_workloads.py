"""The synthetic workloads restate DATA of the reference (src/config.py:7-78, src/predictor.py:38-67, :241-252).

When /root/reference is present (this container, not the GPU box) the tables are compared with the reference
modules themselves; everywhere else the structural checks still run."""
import math
import os
import sys

import pytest

import mcgp_b200

wl = mcgp_b200.workloads
REF = "/root/reference"


def _ref_config():
    if not os.path.isdir(REF):
        pytest.skip("reference tree not present on this box")
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    try:
        from src import config as ref_config
    finally:
        sys.path.remove(REF)
    return ref_config


def test_tables_equal_the_reference_config():
    rc = _ref_config()
    assert list(wl.DRIVER_TEAMS.items()) == list(rc.DRIVER_TEAMS.items())          # order matters (driver universe)
    assert wl.DEFAULT_DNF_RATES == rc.DEFAULT_DNF_RATES
    assert wl.TIRE_COMPOUNDS == rc.TIRE_COMPOUNDS
    assert list(wl.CIRCUITS) == list(rc.CIRCUITS)
    for name, c in rc.CIRCUITS.items():
        assert wl.CIRCUITS[name] == {k: c[k] for k in ("laps", "pit_loss", "drs_zones", "overtake_delta")}, name


def test_race_config_kwargs_build_the_reference_dataclass():
    _ref_config()
    sys.path.insert(0, REF)
    try:
        from src.simulation import RaceConfig as RefRaceConfig
    finally:
        sys.path.remove(REF)
    for name in ("bahrain", "monaco_sc", "sprint19", "season:23", "point:quali"):
        cfg, mc = wl.workload(name)
        ref = RefRaceConfig(**cfg)                      # same field names as src/simulation.py:39-52
        ours = mcgp_b200.simulation.RaceConfig(**cfg)
        for f in ref.__dataclass_fields__:
            assert getattr(ref, f) == getattr(ours, f), (name, f)


def test_common_inputs_follow_survey_8d():
    cfg, mc = wl.workload("bahrain")
    D = list(wl.DRIVER_TEAMS)
    assert cfg["total_laps"] == 57 and cfg["pit_loss"] == 21.0 and cfg["overtake_delta"] == 0.6
    assert (cfg["sc_probability"], cfg["vsc_probability"], cfg["red_flag_probability"]) == (0.01, 0.015, 0.002)
    assert list(mc["grid_probs"]) == D and all(len(v) == 20 for v in mc["grid_probs"].values())
    for k, d in enumerate(D):
        row = mc["grid_probs"][d]
        assert abs(sum(row) - 1.0) < 1e-12 and max(range(20), key=row.__getitem__) == k
        assert all(type(x) is float for x in row)       # exact Python floats: the Neumaier sum() path (Q12)
        assert mc["base_pace"][d] == 92.0 + 0.07 * k and mc["tire_deg"][d] == 0.015 + 0.003 * k
        assert mc["driver_variance"][d] == 0.12 + 0.005 * (k % 5) and mc["driver_dnf_rates"][d] == 0.05 / 57
    w = [math.exp(-((p - 3) ** 2) / 12.5) for p in range(20)]
    assert mc["grid_probs"][D[3]] == [x / sum(w) for x in w]


def test_named_workloads():
    cfg, mc = wl.workload("monaco_sc")
    assert cfg["total_laps"] == 78 and cfg["sc_probability"] == 0.05 and cfg["overtake_delta"] == 1.5
    assert next(iter(mc["driver_dnf_rates"].values())) == 0.05 / 78
    assert wl.workload("sprint19")[0]["total_laps"] == 19
    assert wl.N_SEASON_RACES == 24
    laps = [wl.workload(f"season:{r}")[0]["total_laps"] for r in range(24)]
    assert laps[0] == 57 and laps[6] == 78 and laps[23] == 58        # ... + the fallback circuit of src/predictor.py:38-43
    p0, p5 = wl.workload("season:0")[1]["base_pace"], wl.workload("season:5")[1]["base_pace"]
    assert p0 != p5 and abs(p0["VER"] - (92.0 + 0.05 * (0 % 11 - 5) / 5)) < 1e-12
    base = wl.workload("bahrain")[1]["driver_variance"]
    for point, mult in wl.UNCERTAINTY_MULTIPLIER.items():
        cfg, mc = wl.workload(f"point:{point}")
        assert mc["driver_variance"] == {d: min(0.3, v * mult) for d, v in base.items()}
        onehot = all(sorted(row) == [0.0] * 19 + [1.0] for row in mc["grid_probs"].values())
        assert onehot == (point in ("quali", "sprint"))
    with pytest.raises(KeyError):
        wl.workload("nope")
