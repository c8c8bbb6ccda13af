"""Grid-probability producer: the step just before the race simulator.

Turns qualifying Elo ratings (+ optional per-driver features and grid penalties) into the ``grid_probs`` argument of
``RaceSimulator.run_monte_carlo``.  Behavioural source, reproduced bit for bit (same IEEE operations in the same
order, checked against the unmodified reference in ``tests/test_grid_model.py``):

  * ``pole_probabilities``   reference ``src/elo.py:124-141``  (softmax of rating / 100, max-subtracted)
  * ``quali_distributions``  reference ``src/predictor.py:321-375`` (teammate boost, form / circuit adjustment,
                             Gaussian position spread around ``(1 - p) * n`` -- with sum(p) = 1 that centres every
                             driver near the back of the grid, a quirk the reference has and this keeps)
  * ``apply_grid_penalties`` reference ``src/predictor.py:377-407`` (+ ``PENALTY_TYPES``, ``src/config.py:81-86``)
  * ``grid_probabilities``   the composition used at the fp1/fp2/fp3 prediction points (``src/predictor.py:214-220``)

This is host-side glue of O(n^2) work per race; it exists so that a whole season of races can be prepared and
handed to ONE batched launch (``simulation.run_batch``).  The rows it returns hold ``np.float64`` items, exactly like
the reference's, so the simulator's grid sampling takes the plain-summation path of CPython's ``sum()``.
"""
from __future__ import annotations

from typing import Mapping

import numpy as np

# grid places lost per penalty type, reference src/config.py:81-86
PENALTY_TYPES: dict[str, int] = {"engine": 10, "full_pu": 20, "gearbox": 5, "pitlane_start": 20}

ELO_SCALE = 100          # rating points per e-fold of pole probability (src/elo.py:134)
INITIAL_RATING = 1500.0  # F1EloSystem default (src/elo.py:7)


def _quali_rating(ratings: Mapping, driver: str, initial: float) -> float:
    """Accepts the reference's nested layout {driver: {'quali': r, 'race': r}} or a flat {driver: r}."""
    entry = ratings.get(driver)
    if entry is None:
        return initial
    if isinstance(entry, Mapping):
        return entry.get("quali", initial)
    return entry


def pole_probabilities(ratings: Mapping, drivers: list[str], initial: float = INITIAL_RATING) -> dict:
    """P(pole) per driver from the qualifying ratings (src/elo.py:124-141)."""
    if not drivers:
        return {}
    scaled = {d: _quali_rating(ratings, d, initial) / ELO_SCALE for d in drivers}
    top = max(scaled.values())
    weight = {d: np.exp(v - top) for d, v in scaled.items()}
    norm = sum(weight.values())
    if norm > 0:
        return {d: w / norm for d, w in weight.items()}
    return {d: 1.0 / len(drivers) for d in drivers}


def quali_distributions(drivers: list[str], pole_probs: Mapping, features: Mapping | None = None) -> dict:
    """Per-driver distribution over grid positions (src/predictor.py:332-375), given the pole probabilities."""
    if not drivers:
        return {}
    features = features or {}
    probs = dict(pole_probs)
    # teammate comparison: +-25 % per unit of delta, clamped to [0.5, 1.5]  (:333-339)
    for d in drivers:
        delta = features.get(d, {}).get("teammate_delta", 0)
        if delta != 0 and d in probs:
            probs[d] = probs[d] * max(0.5, min(1.5, 1 + (delta * 0.25)))
    norm = sum(probs.values())
    if norm > 0:
        probs = {d: p / norm for d, p in probs.items()}

    n = len(drivers)
    spread = max(1.0, n / 4)                                   # :362
    out = {}
    for d in drivers:
        f = features.get(d, {})
        form = f.get("form_score", 0) * 0.15                   # :353
        circuit = f.get("circuit_affinity", 0) * 0.10          # :354
        p = probs.get(d, 1 / n) * (1 + form + circuit)
        p = max(0.001, min(0.999, p))                          # :357
        centre = (1 - p) * n                                   # :364
        bell = [np.exp(-((pos - centre) ** 2) / (2 * spread ** 2)) for pos in range(n)]
        norm = sum(bell)
        out[d] = [b / norm for b in bell] if norm > 0 else [1.0 / n] * n
    return out


def apply_grid_penalties(quali_probs: Mapping, penalties: Mapping | None) -> dict:
    """Shift penalised drivers' distributions towards the back (src/predictor.py:377-407)."""
    penalties = penalties or {}
    out = {}
    for d, row in quali_probs.items():
        places = penalties.get(d, 0)
        if isinstance(places, str):
            places = PENALTY_TYPES.get(places, 0)
        n = len(row)
        if not (places > 0 and n > 0):
            out[d] = row
        elif places >= n:
            out[d] = [0.0] * (n - 1) + [1.0]
        else:
            moved = [0.0] * n
            for i, p in enumerate(row):
                moved[min(i + places, n - 1)] += p
            out[d] = moved
    return out


def grid_probabilities(drivers: list[str], ratings: Mapping, features: Mapping | None = None,
                       penalties: Mapping | None = None, initial: float = INITIAL_RATING) -> dict:
    """ratings (+ features, penalties) -> ``grid_probs`` for ``RaceSimulator.run_monte_carlo`` (:214-220)."""
    rows = quali_distributions(drivers, pole_probabilities(ratings, drivers, initial), features)
    return apply_grid_penalties(rows, penalties) if penalties else rows
