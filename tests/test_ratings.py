"""Pairwise Elo: bit-exact to the reference's F1EloSystem on a committed golden season and,
when /root/reference is present, on random event streams."""
import json
import os
import random
import sys

import pytest

import mcgp_b200
from conftest import GOLDEN_DIR

PairwiseElo = mcgp_b200.ratings.PairwiseElo


def _events(rnd, drivers, n_events):
    ev = []
    for e in range(n_events):
        field = [d for d in drivers if rnd.random() < 0.9] or drivers[:2]
        rnd.shuffle(field)
        if rnd.random() < 0.5:
            times = [round(rnd.uniform(78.0, 82.0), 3 if rnd.random() < 0.8 else 1) for _ in field]   # coarse times -> ties
            ev.append(("quali", [[d, t] for d, t in zip(field, times)], rnd.choice([0, 0, 1, 2, 3]), e % 24))
        else:
            pos = list(range(1, len(field) + 1))
            if rnd.random() < 0.2 and len(pos) > 2:
                pos[1] = pos[0]
            ev.append(("race", [[d, p] for d, p in zip(field, pos)], rnd.choice([0, 0, 1, 2, 3]), e % 24))
    return ev


def _apply(elo, events):
    for kind, results, years_ago, race_index in events:
        elo.set_recency_weight(years_ago, race_index)
        res = [tuple(r) for r in results]
        (elo.update_quali_ratings if kind == "quali" else elo.update_race_ratings)(res)
    return {d: {k: float(v).hex() for k, v in r.items()} for d, r in elo.ratings.items()}


def test_golden_season_bit_exact():
    with open(os.path.join(GOLDEN_DIR, "ratings.json")) as f:
        g = json.load(f)
    elo = PairwiseElo()
    assert _apply(elo, g["events"]) == g["ratings"]
    assert {d: float(v).hex() for d, v in elo.predict_quali_probs(g["drivers"]).items()} == g["pole"]


def test_random_event_streams_equal_the_reference():
    if not os.path.isdir("/root/reference"):
        pytest.skip("reference tree not present on this box")
    sys.dont_write_bytecode = True
    sys.path.insert(0, "/root/reference")
    try:
        from src.elo import F1EloSystem
    finally:
        sys.path.remove("/root/reference")
    rnd = random.Random(99)
    for trial in range(20):
        D = [f"E{k}" for k in range(rnd.randint(2, 22))]
        ev = _events(rnd, D, rnd.randint(1, 30))
        k, init = rnd.choice([(32, 1500), (24.0, 1400.0)])
        assert _apply(PairwiseElo(k, init), ev) == _apply(F1EloSystem(k, init), ev), trial


def test_small_fields_and_accessors():
    elo = PairwiseElo()
    elo.update_quali_ratings([("A", 80.0)])                  # fewer than two drivers: ignored
    assert elo.ratings == {} and elo.get_rating("A") == 1500 and elo.get_rating("A", "race") == 1500
    elo.update_race_ratings([("A", 1), ("B", 2)])
    assert elo.get_rating("A", "race") == 1516.0 and elo.get_rating("B", "race") == 1484.0 and elo.get_rating("A") == 1500
    assert abs(elo.expected_score(1500, 1900) - 1 / 11) < 1e-15 and elo.expected_score(0, 1e9) == 1 / (1 + 10 ** 10)
    elo.set_recency_weight(0, 23)
    assert elo.k == 48.0
    elo.set_recency_weight(5)
    assert elo.k == 16.0
