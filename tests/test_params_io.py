"""Race-parameter files: byte-exact round trip of the C parameter blocks."""
import ctypes as C

import numpy as np
import pytest

import mcgp_b200


def _blocks():
    wl, out, names = mcgp_b200.workloads, [], []
    for r, name in enumerate(("bahrain", "monaco_sc", "sprint19", "point:quali", "season:23")):
        cfg, mc = wl.workload(name)
        sim = mcgp_b200.simulation.RaceSimulator(mcgp_b200.simulation.RaceConfig(**cfg), pop_no_medium="SOFT", pop_no_soft="HARD")
        out.append(sim._params(mc["grid_probs"], mc["base_pace"], mc["tire_deg"], mc["driver_variance"], mc["driver_dnf_rates"],
                               mc["track_condition"], stream=r))
        names.append(list(mc["grid_probs"]))
    return out, names


def test_round_trip_is_byte_exact(tmp_path):
    blocks, names = _blocks()
    path = str(tmp_path / "weekend.npz")
    mcgp_b200.params_io.save_race_params(path, blocks, drivers=names, labels=["a", "b", "c", "d", "e"])
    loaded, meta = mcgp_b200.params_io.load_race_params(path)
    assert meta["n_races"] == 5 and meta["drivers"] == names and meta["labels"][2] == "c"
    assert len(loaded) == 5
    for a, b in zip(blocks, loaded):
        assert bytes(a) == bytes(b) and C.sizeof(a) == C.sizeof(mcgp_b200.capi.McgpRaceParams)


def test_foreign_files_are_rejected(tmp_path):
    path = str(tmp_path / "x.npz")
    np.savez(path, a=np.zeros(3))
    with pytest.raises(ValueError, match="meta"):
        mcgp_b200.params_io.load_race_params(path)
    np.savez(path, meta='{"format": "other/9"}')
    with pytest.raises(ValueError, match="unsupported format"):
        mcgp_b200.params_io.load_race_params(path)
    blocks, _ = _blocks()
    arrays = mcgp_b200.params_io.params_to_arrays(blocks[:1])
    del arrays["grid_kind"]
    np.savez(path, meta='{"format": "mcgp-race-params/1"}', **arrays)
    with pytest.raises(ValueError, match="grid_kind"):
        mcgp_b200.params_io.load_race_params(path)


@pytest.mark.gpu
def test_loaded_blocks_simulate_identically(tmp_path):
    blocks, names = _blocks()
    blocks = [blocks[0], blocks[4]]                       # same field size: one batch
    path = str(tmp_path / "two.npz")
    mcgp_b200.params_io.save_race_params(path, blocks)
    loaded, _ = mcgp_b200.params_io.load_race_params(path)
    eng = mcgp_b200.capi.get_engine(0)
    assert np.array_equal(eng.run_native(blocks, 50000, 0, 9), eng.run_native(loaded, 50000, 0, 9))
