"""Monte Carlo agreement checks between two finish-position count tables."""
import numpy as np


def z_table(a: np.ndarray, na: int, b: np.ndarray, nb: int) -> np.ndarray:
    """Two-sample z score per cell with the pooled-proportion variance p(1-p)(1/na + 1/nb)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    pool = (a + b) / (na + nb)
    var = pool * (1 - pool) * (1.0 / na + 1.0 / nb)
    z = np.zeros_like(pool)
    nz = var > 0
    z[nz] = (a[nz] / na - b[nz] / nb) / np.sqrt(var[nz])
    return z


def assert_tables_agree(a, na, b, nb, what=""):
    """north_star: 'per-driver win/podium/position probabilities agree within 3 sigma Monte Carlo error'.

    win and podium (2n statistics): every one within 3 sigma.  Position table (n*n cells): with 400 cells
    ~1 cell is expected beyond 3 sigma by chance alone (P(|z|>3) = 0.27 %), so the table is held to: at most
    1.5 % of the cells beyond 3 sigma and none beyond 4.5 sigma (P(any of 400 > 4.5) = 0.3 %)."""
    a = np.asarray(a, np.int64)
    b = np.asarray(b, np.int64)
    n = a.shape[0]
    zc = z_table(a, na, b, nb)
    zw = z_table(a[:, 0], na, b[:, 0], nb)
    k = min(3, n)
    zp = z_table(a[:, :k].sum(1), na, b[:, :k].sum(1), nb)
    assert np.abs(zw).max() <= 3.0, f"{what}: win probability off by {np.abs(zw).max():.2f} sigma (driver {np.abs(zw).argmax()})"
    assert np.abs(zp).max() <= 3.0, f"{what}: podium probability off by {np.abs(zp).max():.2f} sigma (driver {np.abs(zp).argmax()})"
    over3 = int((np.abs(zc) > 3.0).sum())
    assert np.abs(zc).max() <= 4.5, f"{what}: cell {np.unravel_index(np.abs(zc).argmax(), zc.shape)} off by {np.abs(zc).max():.2f} sigma"
    assert over3 <= max(1, int(0.015 * zc.size)), f"{what}: {over3} of {zc.size} cells beyond 3 sigma"
    return dict(max_cell=float(np.abs(zc).max()), over3=over3, max_win=float(np.abs(zw).max()), max_podium=float(np.abs(zp).max()))
