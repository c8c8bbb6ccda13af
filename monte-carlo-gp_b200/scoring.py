"""Backtest scoring of simulated seasons: the consumer side of the count tables.

Mirrors the pure functions of the reference's ``src/validation.py`` -- ``brier_score`` (:82-106),
``podium_accuracy`` (:109-130), ``calibration_analysis`` (:133-158) -- with the same arguments, skipping rules and
return values, and adds the glue that turns one batched GPU launch (BASELINE config 4: a 24-race season,
``simulation.run_batch``) into the prediction dicts those functions consume (``src/predictor.py:302-314``).
The reference's FastF1 fetchers (:8-79) and the `backtest_model` loop around them are out of scope (network data).
"""
from __future__ import annotations

import numpy as np

from . import simulation, workloads


def brier_score(predictions: list[dict], actuals: list) -> float:
    """Mean over races of mean_d (p_d - [d == actual])^2; races with no actual / empty or invalid predictions are
    skipped; 1.0 when nothing is scored (src/validation.py:82-106)."""
    race_scores = []
    for pred, actual in zip(predictions, actuals):
        if actual is None or not pred:
            continue
        probs = list(pred.values())
        if not all(0 <= p <= 1 for p in probs):
            continue
        race_score = 0.0
        for driver, prob in pred.items():
            outcome = 1.0 if driver == actual else 0.0
            race_score += (prob - outcome) ** 2
        race_scores.append(race_score / len(pred))
    return float(np.mean(race_scores)) if race_scores else 1.0


def podium_accuracy(predictions: list[dict], actuals: list[dict]) -> float:
    """Share of the actual podium found among the three highest podium probabilities (src/validation.py:109-130)."""
    correct = total = 0
    for pred, act in zip(predictions, actuals):
        if not act.get('podium'):
            continue
        podium_probs = pred.get('podium_probabilities', {})
        if not podium_probs:
            continue
        predicted = sorted(podium_probs.items(), key=lambda x: x[1], reverse=True)[:3]  # stable, like the reference
        correct += len({d for d, _ in predicted} & set(act['podium']))
        total += 3
    return correct / total if total > 0 else 0.0


def calibration_analysis(predictions: list[dict], actuals: list[dict]) -> dict:
    """Reliability curve of the win probabilities (src/validation.py:133-158).  The reference calls sklearn's
    ``calibration_curve(..., n_bins)`` (uniform bins); restated here so the product path has no sklearn dependency."""
    all_probs, all_outcomes = [], []
    for pred, act in zip(predictions, actuals):
        if not act.get('winner'):
            continue
        win_probs = pred.get('win_probabilities', {})
        if not win_probs:
            continue
        for driver, prob in win_probs.items():
            all_probs.append(prob)
            all_outcomes.append(1 if driver == act['winner'] else 0)
    if not all_probs:
        return {'prob_true': [], 'prob_pred': []}
    n_bins = min(10, max(2, len(all_probs) // 10))
    y_true, y_prob = np.asarray(all_outcomes, np.float64), np.asarray(all_probs, np.float64)
    if y_prob.min() < 0 or y_prob.max() > 1 or len(np.unique(y_true)) > 2:
        return {'prob_true': [], 'prob_pred': []}  # sklearn raises ValueError, the reference returns empties
    bins = np.linspace(0.0, 1.0, n_bins + 1)
    binids = np.searchsorted(bins[1:-1], y_prob)
    bin_sums = np.bincount(binids, weights=y_prob, minlength=len(bins))
    bin_true = np.bincount(binids, weights=y_true, minlength=len(bins))
    bin_total = np.bincount(binids, minlength=len(bins))
    nonzero = bin_total != 0
    return {'prob_true': (bin_true[nonzero] / bin_total[nonzero]).tolist(),
            'prob_pred': (bin_sums[nonzero] / bin_total[nonzero]).tolist()}


def predictions_from_counts(hist: np.ndarray, drivers: list[str], n_simulations: int) -> dict:
    """The prediction dict of predict_weekend (src/predictor.py:302-314) from one race's count table."""
    race_probs = simulation.counts_to_probabilities(hist, drivers, n_simulations)
    return {
        'win_probabilities': {d: race_probs.get(d, {}).get(1, 0) for d in drivers},
        'podium_probabilities': {d: sum(race_probs.get(d, {}).get(p, 0) for p in [1, 2, 3]) for d in drivers},
        'full_distributions': race_probs,
    }


def simulate_season(n_simulations: int, seed: int, races: list[int] | None = None, device: int | None = None,
                    pop_no_medium: str | None = None, pop_no_soft: str | None = None):
    """BASELINE config 4: every race of the synthetic 24-race season in ONE kernel launch (race r -> stream r).
    Returns (count tables [R, n, n] uint64, list of prediction dicts)."""
    from . import capi
    races = list(range(workloads.N_SEASON_RACES)) if races is None else list(races)
    params, drivers = [], None
    for r in races:
        cfg, mc = workloads.workload(f"season:{r}")
        sim = simulation.RaceSimulator(simulation.RaceConfig(**cfg), device=device, pop_no_medium=pop_no_medium,
                                       pop_no_soft=pop_no_soft)
        params.append(sim._params(mc['grid_probs'], mc['base_pace'], mc['tire_deg'], mc['driver_variance'],
                                  mc['driver_dnf_rates'], mc['track_condition'], stream=r))
        drivers = list(mc['grid_probs'])
    hist = capi.get_engine(sim.device).run_native(params, int(n_simulations), 0, int(seed) & (2 ** 64 - 1))
    return hist, [predictions_from_counts(hist[i], drivers, n_simulations) for i in range(len(races))]


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
