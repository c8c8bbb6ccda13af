"""Pairwise Elo ratings for qualifying and race results.

Behavioural source: reference ``src/elo.py:6-145`` (``F1EloSystem``), reproduced bit for bit -- every update adds
``k * (actual - expected) / (n - 1)`` per opponent in result-list order, against the ratings as they were BEFORE the
event (checked against the unmodified reference in ``tests/test_ratings.py``).  Together with ``grid_model`` and the
batched simulator this closes the loop "simulate a race -> update ratings -> next race" without the reference's
pandas / FastF1 layers.  O(n^2) host work per event.
"""
from __future__ import annotations

from . import grid_model


class PairwiseElo:
    """Drop-in for the reference's ``F1EloSystem`` (same attribute and method names)."""

    def __init__(self, k_factor: float = 32, initial_rating: float = 1500):
        self.base_k = k_factor
        self.k = k_factor
        self.initial = initial_rating
        self.ratings: dict = {}          # driver -> {'quali': r, 'race': r}

    # src/elo.py:13-37
    def set_recency_weight(self, years_ago: float, race_index: int = 0, total_races: int = 24):
        if years_ago <= 0:               # current season: 0.75x for the first race ... 1.5x for the last
            self.k = self.base_k * (0.75 + (0.75 * race_index / max(1, total_races - 1)))
        elif years_ago <= 1:
            self.k = self.base_k * 1.0
        elif years_ago <= 2:
            self.k = self.base_k * 0.7
        else:
            self.k = self.base_k * 0.5

    # src/elo.py:39-42
    def expected_score(self, rating_a: float, rating_b: float) -> float:
        return 1 / (1 + 10 ** max(-10, min(10, (rating_b - rating_a) / 400)))

    def _update(self, kind: str, results: list[tuple]) -> None:
        """One event: `results` is [(driver, value)], lower value = better (lap time or finishing position)."""
        n = len(results)
        if n < 2:
            return
        for driver, _ in results:
            if driver not in self.ratings:
                self.ratings[driver] = {"quali": self.initial, "race": self.initial}
        before = [self.ratings[d][kind] for d, _ in results]
        change = {}
        for i, (driver, mine) in enumerate(results):
            acc = 0
            for j, (_, theirs) in enumerate(results):
                if j == i:
                    continue
                outcome = 1.0 if mine < theirs else 0.0 if mine > theirs else 0.5
                acc += self.k * (outcome - self.expected_score(before[i], before[j])) / (n - 1)
            change[driver] = acc         # (a driver listed twice keeps the LAST delta, as upstream's dict does)
        for driver, acc in change.items():
            self.ratings[driver][kind] += acc

    def update_quali_ratings(self, quali_results: list[tuple[str, float]]):   # src/elo.py:45-82
        self._update("quali", quali_results)

    def update_race_ratings(self, race_results: list[tuple[str, int]]):       # src/elo.py:84-122
        self._update("race", race_results)

    def predict_quali_probs(self, drivers: list[str]) -> dict:                # src/elo.py:124-141
        return grid_model.pole_probabilities(self.ratings, drivers, self.initial)

    def get_rating(self, driver: str, rating_type: str = "quali") -> float:   # src/elo.py:143-145
        return self.ratings.get(driver, {}).get(rating_type, self.initial)
