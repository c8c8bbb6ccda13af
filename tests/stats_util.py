"""Monte Carlo agreement checks between two finish-position count tables."""
import numpy as np


def z_table(a, na: int, b, nb: int) -> np.ndarray:
    """Two-sample z score per cell with the pooled-proportion variance p(1-p)(1/na + 1/nb)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    pool = (a + b) / (na + nb)
    var = pool * (1 - pool) * (1.0 / na + 1.0 / nb)
    z = np.zeros_like(pool)
    nz = var > 0
    z[nz] = (a[nz] / na - b[nz] / nb) / np.sqrt(var[nz])
    return z


def compare_tables(a, na, b, nb) -> dict:
    """z scores of the statistics north_star names: per-driver win, podium, and every position cell."""
    a = np.asarray(a, np.int64)
    b = np.asarray(b, np.int64)
    k = min(3, a.shape[0])
    return dict(win=z_table(a[:, 0], na, b[:, 0], nb), podium=z_table(a[:, :k].sum(1), na, b[:, :k].sum(1), nb),
                cells=z_table(a, na, b, nb))


def violations(z: dict) -> list:
    """Statistics outside the gate.  win / podium: every driver within 3 sigma.  Position table: with n*n cells
    ~0.27 % land beyond 3 sigma by chance alone, so the table is held to <= 1.5 % of cells beyond 3 sigma and none
    beyond 4.5 sigma (P(any of 400 beyond 4.5) = 0.3 %)."""
    out = [("win", int(i)) for i in np.nonzero(np.abs(z["win"]) > 3.0)[0]]
    out += [("podium", int(i)) for i in np.nonzero(np.abs(z["podium"]) > 3.0)[0]]
    zc = np.abs(z["cells"])
    out += [("cell", int(i), int(j)) for i, j in np.argwhere(zc > 4.5)]
    if (zc > 3.0).sum() > max(1, int(0.015 * zc.size)):
        out.append(("cells>3", int((zc > 3.0).sum())))
    return out


def summary(z: dict) -> dict:
    return dict(max_win=round(float(np.abs(z["win"]).max()), 2), max_podium=round(float(np.abs(z["podium"]).max()), 2),
                max_cell=round(float(np.abs(z["cells"]).max()), 2), cells_over_3=int((np.abs(z["cells"]) > 3).sum()))


def assert_agree_two_stage(sample_a, sample_b, na, nb, what="", factor=4):
    """'Agree within 3 sigma Monte Carlo error' as a two-stage test.

    sample_a(n, stage) / sample_b(n, stage) return count tables from independent streams per stage.  Stage 1 checks
    every statistic.  Dozens of statistics are examined per case, so a >3 sigma excursion somewhere is not rare under
    perfect agreement; anything flagged is therefore re-measured in stage 2 with `factor` x the sims and fresh
    streams, where a chance excursion vanishes and a real bias grows by sqrt(factor).  Stage 2 must be clean on the
    flagged statistics."""
    z1 = compare_tables(sample_a(na, 1), na, sample_b(nb, 1), nb)
    v1 = violations(z1)
    if not v1:
        return dict(stage=1, **summary(z1))
    na2, nb2 = na * factor, nb * factor
    z2 = compare_tables(sample_a(na2, 2), na2, sample_b(nb2, 2), nb2)
    v2 = [v for v in violations(z2) if v in v1 or v[0] == "cells>3"]
    assert not v2, f"{what}: disagreement confirmed at {factor}x the sims: {v2} (stage 1 flagged {v1}); {summary(z2)}"
    return dict(stage=2, flagged=v1, **summary(z2))
