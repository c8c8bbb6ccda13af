"""Backtest scoring of simulated seasons: the consumer side of the count tables.

What the reference computes with loops over prediction dicts (``src/validation.py``: ``brier_score`` :82-106,
``podium_accuracy`` :109-130, ``calibration_analysis`` :133-158) is computed here on the ``[R, n, n]`` count tensor a
batched launch produces -- on the GPU, without the tables leaving it (``score_counts_device`` ->
``mcgp_score_counts``, csrc/season_kernels.cu), or vectorised over the tensor on the host (``score_counts``).  The
dict-shaped entry points of the reference (same names, arguments, skipping rules and return values) are thin
adapters over the same array code, so a caller of ``src/validation.py`` can switch without changes.
Every floating-point sum runs in the reference's order (``np.add.accumulate`` is sequential), so the results are the
reference's bit for bit (tests/test_scoring.py: KATs produced by the reference, and the reference's own functions
where it is importable).  The FastF1 fetchers (:8-79) are out of scope (network data); the backtest loop itself is
``season.run_device_season``.
"""
from __future__ import annotations

import numpy as np

from . import simulation, workloads


# ---- array core -------------------------------------------------------------------------------------
def _seq_sum(a: np.ndarray, axis: int = -1) -> np.ndarray:
    """Left-to-right sum along `axis` (np.sum is pairwise: different rounding than the reference's += loops)."""
    a = np.asarray(a, np.float64)
    if a.shape[axis] == 0:
        return np.zeros(np.delete(a.shape, axis))
    return np.take(np.add.accumulate(a, axis=axis), -1, axis=axis)


def brier_terms(win_probs: np.ndarray, winners: np.ndarray) -> np.ndarray:
    """Per-race Brier term mean_d (p_d - [d == winner])^2 of win probabilities [R, n]; NaN where winners[r] < 0."""
    p = np.asarray(win_probs, np.float64)
    w = np.asarray(winners, np.int64)
    R, n = p.shape
    outcome = np.zeros_like(p)
    ok = w >= 0
    outcome[np.nonzero(ok)[0], w[ok]] = 1.0
    terms = _seq_sum((p - outcome) ** 2, 1) / n
    terms[~ok] = np.nan
    return terms


def top3_hits(podium_probs: np.ndarray, podiums: np.ndarray) -> np.ndarray:
    """|three highest podium probabilities (ties: lower driver index first) ∩ actual podium| per race; -1 = skipped."""
    p = np.asarray(podium_probs, np.float64)
    a = np.asarray(podiums, np.int64).reshape(len(p), 3)
    top = np.argsort(-p, axis=1, kind="stable")[:, :3]
    hits = (top[:, :, None] == a[:, None, :]).any(2).sum(1).astype(np.int32)
    hits[a[:, 0] < 0] = -1
    return hits


def calibration_bins(win_probs: np.ndarray, winners: np.ndarray) -> dict:
    """Uniform-bin reliability curve of the win probabilities over the races with a winner (:133-158)."""
    p = np.asarray(win_probs, np.float64)
    w = np.asarray(winners, np.int64)
    ok = w >= 0
    y_prob = p[ok].ravel()
    if y_prob.size == 0:
        return {"prob_true": [], "prob_pred": []}
    y_true = np.zeros_like(p[ok])
    y_true[np.arange(ok.sum()), w[ok]] = 1.0
    y_true = y_true.ravel()
    if y_prob.min() < 0 or y_prob.max() > 1:
        return {"prob_true": [], "prob_pred": []}          # sklearn raises ValueError, the reference returns empties
    n_bins = min(10, max(2, y_prob.size // 10))
    edges = np.linspace(0.0, 1.0, n_bins + 1)
    ids = np.searchsorted(edges[1:-1], y_prob)
    total = np.bincount(ids, minlength=len(edges))
    true = np.bincount(ids, weights=y_true, minlength=len(edges))
    pred = np.bincount(ids, weights=y_prob, minlength=len(edges))
    nz = total != 0
    return {"prob_true": (true[nz] / total[nz]).tolist(), "prob_pred": (pred[nz] / total[nz]).tolist()}


def score_counts(hist: np.ndarray, n_simulations: int, winners, podiums=None) -> dict:
    """Host-side scoring of count tables hist[R, n, n] (driver, position) against actual winners / podiums given as
    driver INDICES (-1 = unknown): tallies, Brier terms + season score, podium hits + accuracy, calibration curve."""
    h = np.asarray(hist)
    R, n, _ = h.shape
    win = h[:, :, 0] / n_simulations
    # sum(race_probs[d].get(p, 0) for p in [1, 2, 3]) (src/predictor.py:310-313): three rounded quotients, added in order
    pod = _seq_sum(h[:, :, : min(3, n)] / n_simulations, 2)
    terms = brier_terms(win, winners)
    scored = terms[~np.isnan(terms)]
    out = {"tallies": np.stack([h[:, :, 0], h[:, :, : min(3, n)].sum(2), h[:, :, : min(10, n)].sum(2)], 1),
           "brier_terms": terms, "win_brier": float(np.mean(scored)) if scored.size else 1.0,
           "calibration_curve": calibration_bins(win, winners)}
    if podiums is not None:
        hits = top3_hits(pod, podiums)
        done = hits >= 0
        out["podium_hits"] = hits
        out["podium_accuracy"] = float(hits[done].sum() / (3 * done.sum())) if done.any() else 0.0
    return out


def score_counts_device(hist_dev, n_races: int, n: int, n_simulations: int, winners, podiums=None, device: int = 0,
                        stream=None, engine=None) -> dict:
    """The same scores from count tables that live on the GPU: `hist_dev` is a device pointer (e.g. a torch tensor's
    data_ptr()) to uint64 / int64 [R, n, n]; one small kernel, only the results are copied back."""
    from . import capi
    eng = engine or capi.get_engine(device)
    raw = eng.score_counts(hist_dev, n_races, n, n_simulations, winners, podiums, stream)
    terms = raw["brier"]
    scored = terms[~np.isnan(terms)]
    bins = int(raw["calib_bins"][0])
    total, true, pred = raw["calib"][0, : bins], raw["calib"][1, : bins], raw["calib"][2, : bins]
    nz = total != 0
    out = {"tallies": raw["tallies"], "brier_terms": terms, "win_brier": float(np.mean(scored)) if scored.size else 1.0,
           "calibration_curve": ({"prob_true": (true[nz] / total[nz]).tolist(), "prob_pred": (pred[nz] / total[nz]).tolist()}
                                 if total.sum() else {"prob_true": [], "prob_pred": []})}
    if podiums is not None:
        hits = raw["podium_hits"]
        done = hits >= 0
        out["podium_hits"] = hits
        out["podium_accuracy"] = float(hits[done].sum() / (3 * done.sum())) if done.any() else 0.0
    return out


# ---- the reference's dict-shaped entry points (src/validation.py:82-158) as adapters --------------
def _rows(dicts: list[dict]):
    """[{driver: p}] with one common key order -> array [R, n]; None if the races do not share one driver list."""
    keys = list(dicts[0])
    if any(list(d) != keys for d in dicts):
        return None, None
    return keys, np.array([[d[k] for k in keys] for d in dicts], np.float64)


def brier_score(predictions: list[dict], actuals: list) -> float:
    """src/validation.py:82-106: mean over the scored races of mean_d (p_d - [d == actual])^2; 1.0 if none is scored."""
    terms = []
    for pred, actual in zip(predictions, actuals):
        if actual is None or not pred:
            continue
        keys, p = list(pred), np.array(list(pred.values()), np.float64)
        if not ((p >= 0) & (p <= 1)).all():
            continue                                         # invalid probabilities: the race is skipped (:95-97)
        w = keys.index(actual) if actual in pred else -1
        outcome = np.zeros_like(p)
        if w >= 0:
            outcome[w] = 1.0
        terms.append(_seq_sum((p - outcome) ** 2) / len(p))
    return float(np.mean(terms)) if terms else 1.0


def podium_accuracy(predictions: list[dict], actuals: list[dict]) -> float:
    """src/validation.py:109-130: share of the actual podium found among the three highest podium probabilities."""
    hits = races = 0
    for pred, act in zip(predictions, actuals):
        probs = pred.get("podium_probabilities", {})
        if not act.get("podium") or not probs:
            continue
        keys = list(probs)
        order = np.argsort(-np.array(list(probs.values()), np.float64), kind="stable")[:3]
        hits += len({keys[i] for i in order} & set(act["podium"]))
        races += 1
    return hits / (3 * races) if races else 0.0


def calibration_analysis(predictions: list[dict], actuals: list[dict]) -> dict:
    """src/validation.py:133-158 (sklearn's uniform-bin calibration_curve written out: no sklearn in the product)."""
    probs, outcomes = [], []
    for pred, act in zip(predictions, actuals):
        win = pred.get("win_probabilities", {})
        if not act.get("winner") or not win:
            continue
        probs.extend(win.values())
        outcomes.extend(1.0 if d == act["winner"] else 0.0 for d in win)
    if not probs:
        return {"prob_true": [], "prob_pred": []}
    y_prob, y_true = np.array(probs, np.float64), np.array(outcomes, np.float64)
    if y_prob.min() < 0 or y_prob.max() > 1:
        return {"prob_true": [], "prob_pred": []}
    n_bins = min(10, max(2, len(probs) // 10))
    edges = np.linspace(0.0, 1.0, n_bins + 1)
    ids = np.searchsorted(edges[1:-1], y_prob)
    total = np.bincount(ids, minlength=len(edges))
    nz = total != 0
    return {"prob_true": (np.bincount(ids, weights=y_true, minlength=len(edges))[nz] / total[nz]).tolist(),
            "prob_pred": (np.bincount(ids, weights=y_prob, minlength=len(edges))[nz] / total[nz]).tolist()}


# ---- glue: count tables -> the prediction dicts of predict_weekend ------------------------------------
def predictions_from_counts(hist: np.ndarray, drivers: list[str], n_simulations: int) -> dict:
    """The prediction dict of predict_weekend (src/predictor.py:302-314) from one race's count table."""
    race_probs = simulation.counts_to_probabilities(hist, drivers, n_simulations)
    return {
        'win_probabilities': {d: race_probs.get(d, {}).get(1, 0) for d in drivers},
        'podium_probabilities': {d: sum(race_probs.get(d, {}).get(p, 0) for p in [1, 2, 3]) for d in drivers},
        'full_distributions': race_probs,
    }


def season_params(races: list[int] | None = None, device: int | None = None, pop_no_medium: str | None = None,
                  pop_no_soft: str | None = None):
    """The mcgp_race_params blocks of the synthetic season (BASELINE config 4), race r on draw stream r."""
    races = list(range(workloads.N_SEASON_RACES)) if races is None else list(races)
    params, drivers, sim = [], None, None
    for r in races:
        cfg, mc = workloads.workload(f"season:{r}")
        sim = simulation.RaceSimulator(simulation.RaceConfig(**cfg), device=device, pop_no_medium=pop_no_medium,
                                       pop_no_soft=pop_no_soft)
        params.append(sim._params(mc['grid_probs'], mc['base_pace'], mc['tire_deg'], mc['driver_variance'],
                                  mc['driver_dnf_rates'], mc['track_condition'], stream=r))
        drivers = list(mc['grid_probs'])
    return params, drivers, sim.device


def simulate_season(n_simulations: int, seed: int, races: list[int] | None = None, device: int | None = None,
                    pop_no_medium: str | None = None, pop_no_soft: str | None = None):
    """BASELINE config 4: every race of the synthetic 24-race season in ONE kernel launch (race r -> stream r).
    Returns (count tables [R, n, n] uint64, list of prediction dicts)."""
    from . import capi
    params, drivers, dev = season_params(races, device, pop_no_medium, pop_no_soft)
    hist = capi.get_engine(dev).run_native(params, int(n_simulations), 0, int(seed) & (2 ** 64 - 1))
    return hist, [predictions_from_counts(hist[i], drivers, n_simulations) for i in range(len(params))]


def brier_mc_sigma(win_probs: np.ndarray, winners_idx: np.ndarray, n_simulations: int) -> float:
    """1-sigma Monte Carlo error of the season Brier score: delta method over the multinomial win counts.
    win_probs [R, n]; B = mean_r mean_d (p_rd - o_rd)^2  =>  dB/dp_rd = 2 (p_rd - o_rd) / (n R)."""
    p = np.asarray(win_probs, np.float64)
    R, n = p.shape
    o = np.zeros_like(p)
    o[np.arange(R), winners_idx] = 1.0
    a = 2.0 * (p - o) / (n * R)
    var = ((a * a * p).sum(1) - ((a * p).sum(1)) ** 2) / n_simulations
    return float(np.sqrt(max(var.sum(), 0.0)))
