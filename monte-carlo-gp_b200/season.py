"""A whole season resident on the GPU: simulate race r -> Elo update from its result -> race r+1's grid probabilities
-> next launch, one stream, no host round trip (SURVEY.md §8(f) rows 2 + 4; include/mcgp.h: mcgp_run_season).

Upstream this is the loop of ``backtest_model`` (src/validation.py:176-198) around ``F1EloSystem`` (src/elo.py:45-141)
and ``_predict_quali`` / ``_adjust_for_penalties`` (src/predictor.py:321-407), with FastF1 results as the "actual"
outcomes; here the actual outcome of race r is one extra simulated race (global sim index n_sims), so the loop closes
on the device.  ``replay_season_on_host`` runs the same loop with the host ports (``ratings.PairwiseElo``,
``grid_model``) fed with the device's actual results: what the GPU tests hold the device loop against.
"""
from __future__ import annotations

import numpy as np

from . import capi, grid_model, ratings, scoring


def run_device_season(n_simulations: int, seed: int, quali0=None, race0=None, k_factor: float = 32.0, penalties=None,
                      races: list[int] | None = None, device: int | None = None, pop_no_medium: str | None = None,
                      pop_no_soft: str | None = None, flags: int = 0) -> dict:
    """The synthetic season of BASELINE config 4 with grids driven by Elo ratings that evolve on the device.
    Returns count tables, the rating history, the derived grid rows, the actual results and the scores."""
    params, drivers, dev = scoring.season_params(races, device, pop_no_medium, pop_no_soft)
    n = len(drivers)
    q0 = np.full(n, grid_model.INITIAL_RATING) if quali0 is None else np.asarray(quali0, np.float64)
    r0 = np.full(n, grid_model.INITIAL_RATING) if race0 is None else np.asarray(race0, np.float64)
    out = capi.Engine(dev).run_season(params, int(n_simulations), int(seed), q0, r0, k_factor, penalties, flags)
    terms = out["brier"]
    out["drivers"] = drivers
    out["win_brier"] = float(np.mean(terms[~np.isnan(terms)])) if (~np.isnan(terms)).any() else 1.0
    done = out["podium_hits"] >= 0
    out["podium_accuracy"] = float(out["podium_hits"][done].sum() / (3 * done.sum())) if done.any() else 0.0
    return out


def replay_season_on_host(drivers: list[str], actual_grid: np.ndarray, actual_finish: np.ndarray, quali0, race0,
                          k_factor: float = 32.0, penalties=None) -> dict:
    """The same season loop on the host ports, driven by given actual results (driver INDEX per grid slot / finishing
    position): rating history [R + 1, n] (quali, race) and the grid rows [R, n, n] each race would start from."""
    R, n = np.asarray(actual_grid).shape
    elo = ratings.PairwiseElo(k_factor=k_factor)
    elo.ratings = {d: {"quali": float(q), "race": float(r)} for d, q, r in zip(drivers, quali0, race0)}
    quali, race, rows = [], [], []
    for r in range(R):
        quali.append([elo.ratings[d]["quali"] for d in drivers])
        race.append([elo.ratings[d]["race"] for d in drivers])
        pen = None if penalties is None else {d: int(p) for d, p in zip(drivers, np.asarray(penalties).reshape(R, n)[r]) if p}
        g = grid_model.grid_probabilities(drivers, elo.ratings, None, pen)
        rows.append([[float(x) for x in g[d]] for d in drivers])
        elo.update_quali_ratings([(drivers[int(i)], float(slot)) for slot, i in enumerate(actual_grid[r])])
        elo.update_race_ratings([(drivers[int(i)], pos + 1) for pos, i in enumerate(actual_finish[r])])
    quali.append([elo.ratings[d]["quali"] for d in drivers])
    race.append([elo.ratings[d]["race"] for d in drivers])
    return {"quali": np.array(quali), "race": np.array(race), "grid_rows": np.array(rows)}
