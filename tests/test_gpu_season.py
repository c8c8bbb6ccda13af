"""The device-resident season loop (include/mcgp.h: mcgp_run_season; csrc/season_kernels.cu): race r's count table ->
its "actual" result -> pairwise Elo update -> grid rows of race r + 1 -> next launch, with no host round trip.

Held against the host ports that are bit-exact to the reference (ratings.PairwiseElo == src/elo.py, grid_model ==
src/predictor.py:321-407): fed with the device's actual results they must reproduce the device's rating history to
1e-9 rating points and its grid rows to 1e-12 (the only difference is CUDA's exp / pow vs numpy's / libm's, a few ulp);
every count table must be IDENTICAL to a standalone launch of that race from the derived rows; the scores must equal
the checker's on the same counts."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

POP = ("SOFT", "MEDIUM")


@pytest.fixture(scope="module")
def mcgp():
    import mcgp_b200
    return mcgp_b200


@pytest.fixture(scope="module")
def season(mcgp):
    n_sims, seed = 200_000, 77
    pen = np.zeros((24, 20), np.int32)
    pen[3, 5], pen[7, 0], pen[7, 11], pen[20, 19] = 5, 25, 10, 3      # gearbox-like, back of the grid, engine, small
    rng = np.random.default_rng(3)
    q0 = 1500.0 + rng.normal(0, 60, 20)
    r0 = 1500.0 + rng.normal(0, 60, 20)
    out = mcgp.season.run_device_season(n_sims, seed, quali0=q0, race0=r0, k_factor=32.0, penalties=pen,
                                        pop_no_medium=POP[0], pop_no_soft=POP[1])
    return dict(out=out, n_sims=n_sims, seed=seed, pen=pen, q0=q0, r0=r0)


def test_season_outputs_are_consistent(mcgp, season):
    out, n = season["out"], season["n_sims"]
    assert out["hist"].shape == (24, 20, 20)
    assert (out["hist"].sum(1) == n).all() and (out["hist"].sum(2) == n).all()        # every table is doubly stochastic x n
    for r in range(24):                                                                 # actual results are permutations
        assert sorted(out["actual_grid"][r].tolist()) == list(range(20)) == sorted(out["actual_finish"][r].tolist())
    assert np.array_equal(out["quali"][0], season["q0"]) and np.array_equal(out["race"][0], season["r0"])
    assert np.abs(out["grid_rows"].sum(2) - 1.0).max() < 1e-12
    # pairwise Elo is zero-sum up to rounding
    assert np.abs(out["quali"].sum(1) - season["q0"].sum()).max() < 1e-8
    assert np.abs(np.diff(out["quali"], axis=0)).max() > 1.0                              # and the ratings do move


def test_season_matches_the_host_ports(mcgp, season):
    """(f)2 + (f)4: rating history and grid rows vs ratings.PairwiseElo / grid_model driven by the same actual results."""
    out = season["out"]
    host = mcgp.season.replay_season_on_host(out["drivers"], out["actual_grid"], out["actual_finish"], season["q0"],
                                             season["r0"], 32.0, season["pen"])
    assert np.abs(host["quali"] - out["quali"]).max() < 1e-9, np.abs(host["quali"] - out["quali"]).max()
    assert np.abs(host["race"] - out["race"]).max() < 1e-9
    assert np.abs(host["grid_rows"] - out["grid_rows"]).max() < 1e-12, np.abs(host["grid_rows"] - out["grid_rows"]).max()
    # penalties: a driver sent to the back starts last with certainty, a 5-place penalty empties the first five cells
    assert out["grid_rows"][7, 0, 19] == 1.0 and out["grid_rows"][7, 0, :19].sum() == 0.0
    assert out["grid_rows"][3, 5, :5].sum() == 0.0


def test_season_tables_equal_standalone_launches(mcgp, season):
    """Race r of the loop == one ordinary launch of race r whose grid_probs are the rows the device derived (downloaded
    as FP64): identical count tables, identical actual sim.  Nothing of the loop leaks into the simulation."""
    out, n_sims, seed = season["out"], season["n_sims"], season["seed"]
    params, drivers, dev = mcgp.scoring.season_params(pop_no_medium=POP[0], pop_no_soft=POP[1])
    eng = mcgp.capi.Engine(dev)
    for r in (0, 1, 7, 23):
        p = params[r]
        for d in range(20):
            for pos in range(20):
                p.grid_probs[d][pos] = out["grid_rows"][r, d, pos]
                p.grid_kind[d][pos] = mcgp.capi.ITEM_NPFLOAT
        hist = eng.run_native([p], n_sims, 0, seed)
        assert np.array_equal(hist[0], out["hist"][r]), r
        _, finish = eng.run_native([p], 1, n_sims, seed, want_finish=True)
        assert np.array_equal(finish[0, 0], out["actual_finish"][r]), r


def test_season_scores_equal_checker(mcgp, season):
    """(f)1 inside the loop: Brier terms, podium hits, tallies, calibration of the device == checker on the same counts."""
    from oracle import scoring_oracle as chk
    out, n_sims = season["out"], season["n_sims"]
    D = out["drivers"]
    preds = [mcgp.scoring.predictions_from_counts(out["hist"][r], D, n_sims) for r in range(24)]
    acts = [{"winner": D[int(out["actual_finish"][r, 0])], "podium": [D[int(i)] for i in out["actual_finish"][r, :3]]} for r in range(24)]
    b = chk.brier_score([p["win_probabilities"] for p in preds], [a["winner"] for a in acts])
    assert abs(out["win_brier"] - b) <= 1e-12
    assert out["podium_accuracy"] == chk.podium_accuracy(preds, acts)
    assert np.array_equal(out["tallies"][:, 0].astype(np.int64), out["hist"][:, :, 0].astype(np.int64))
    assert np.array_equal(out["tallies"][:, 1].astype(np.int64), out["hist"][:, :, :3].sum(2).astype(np.int64))
    assert np.array_equal(out["tallies"][:, 2].astype(np.int64), out["hist"][:, :, :10].sum(2).astype(np.int64))
    host = mcgp.scoring.score_counts(out["hist"], n_sims, out["actual_finish"][:, 0].astype(int), out["actual_finish"][:, :3].astype(int))
    assert np.allclose(out["brier"], host["brier_terms"], rtol=0, atol=1e-12)
    assert np.array_equal(out["podium_hits"], host["podium_hits"])


def test_season_is_reproducible_and_seed_dependent(mcgp, season):
    again = mcgp.season.run_device_season(season["n_sims"], season["seed"], quali0=season["q0"], race0=season["r0"], k_factor=32.0,
                                          penalties=season["pen"], pop_no_medium=POP[0], pop_no_soft=POP[1])
    assert np.array_equal(again["hist"], season["out"]["hist"]) and np.array_equal(again["quali"], season["out"]["quali"])
    other = mcgp.season.run_device_season(20_000, season["seed"] + 1, quali0=season["q0"], race0=season["r0"], races=[0, 1, 2],
                                          pop_no_medium=POP[0], pop_no_soft=POP[1])
    assert other["hist"].shape == (3, 20, 20) and not np.array_equal(other["actual_finish"][0], season["out"]["actual_finish"][0])
