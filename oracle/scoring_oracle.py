"""TEST INFRASTRUCTURE -- the scoring CHECKER: a plain restatement of the reference's scoring functions
``src/validation.py`` -- ``brier_score`` (:82-106), ``podium_accuracy`` (:109-130), ``calibration_analysis``
(:133-158, with sklearn's uniform-bin ``calibration_curve`` written out) -- loop for loop as upstream, so that the
product's vectorised / on-device scoring (monte-carlo-gp_b200/scoring.py, csrc/season_kernels.cu) has something to be
held against on the GPU box, where /root/reference does not exist.  Pinned to the reference itself by the KATs of
tests/golden/season.json (generator: oracle/gen_season_golden.py) and, where the reference is importable, against
its own functions (tests/test_scoring.py).  Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np


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
