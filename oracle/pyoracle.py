"""TEST INFRASTRUCTURE -- ctypes front-end of the CPU oracle (oracle/race_oracle.c).

Marshals the reference's call signature (RaceConfig kwargs + run_monte_carlo kwargs) into the
oracle's dense parameter block, applying the reference's own `.get` defaults.  This
marshaller is deliberately independent of the product's (monte-carlo-gp_b200/simulation.py) so
that the parity tests also cross-check the host-side marshalling.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
MAXD, NC = 32, 5
COMPOUNDS = ("SOFT", "MEDIUM", "HARD", "INTERMEDIATE", "WET")
TRACK = {"dry": 0, "damp": 1, "wet": 2}


class OrcParams(C.Structure):
    _fields_ = [
        ("n_drivers", C.c_int32), ("total_laps", C.c_int32), ("track_condition", C.c_int32),
        ("pop_no_medium", C.c_int32), ("pop_no_soft", C.c_int32), ("_pad", C.c_int32),
        ("pit_loss", C.c_double), ("overtake_delta", C.c_double), ("sc_p", C.c_double),
        ("vsc_p", C.c_double), ("red_p", C.c_double), ("drs_delta", C.c_double),
        ("dirty_thr", C.c_double), ("dirty_pen", C.c_double),
        ("compound_pace_delta", C.c_double * NC), ("compound_deg_rate", C.c_double * NC),
        ("compound_optimal", C.c_double * NC),
        ("base_pace", C.c_double * MAXD), ("tire_deg", C.c_double * MAXD),
        ("tire_deg_pit", C.c_double * MAXD), ("variance", C.c_double * MAXD),
        ("dnf_rate", C.c_double * MAXD), ("team_rate", C.c_double * MAXD),
        ("grid_probs", (C.c_double * MAXD) * MAXD), ("grid_kind", (C.c_uint8 * MAXD) * MAXD),
    ]


class OrcMT(C.Structure):
    _fields_ = [("mt", C.c_uint32 * 624), ("idx", C.c_int32)]


class OrcRng(C.Structure):
    _fields_ = [("py", OrcMT), ("np", OrcMT), ("has_gauss", C.c_int32), ("_pad", C.c_int32),
                ("gauss", C.c_double)]


class OrcOutputs(C.Structure):
    _fields_ = [
        ("hist", C.c_void_p), ("grid", C.c_void_p), ("finish", C.c_void_p), ("times", C.c_void_p),
        ("dnf_lap", C.c_void_p), ("draws", C.c_void_p),
        ("log_upy", C.c_void_p), ("cap_upy", C.c_int64), ("log_z", C.c_void_p), ("cap_z", C.c_int64),
        ("log_unp", C.c_void_p), ("cap_unp", C.c_int64),
    ]


_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB_PATH) or any(
            os.path.getmtime(os.path.join(HERE, f)) > os.path.getmtime(LIB_PATH)
            for f in os.listdir(HERE) if f.endswith((".c", ".h"))):
        subprocess.check_call(["make", "-s", "-C", HERE])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_rng_seed.argtypes = [C.POINTER(OrcRng), C.c_uint64]
        L.orc_rng_seed.restype = None
        for f in (L.orc_py_random, L.orc_np_random_sample, L.orc_np_standard_normal):
            f.argtypes = [C.POINTER(OrcRng)]
            f.restype = C.c_double
        L.orc_run_streams.argtypes = [C.POINTER(OrcParams), C.POINTER(OrcRng), C.c_int64, C.POINTER(OrcOutputs)]
        L.orc_run_streams.restype = C.c_int
        L.orc_run_tapes.argtypes = [C.POINTER(OrcParams), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.POINTER(OrcOutputs)]
        L.orc_run_tapes.restype = C.c_int
        L.orc_run_native.argtypes = [C.POINTER(OrcParams), C.c_uint64, C.c_uint32, C.c_uint64, C.c_int64, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_run_native.restype = C.c_int
        L.orc_native_op32_table.argtypes = [C.POINTER(OrcParams), C.c_void_p]
        L.orc_native_op32_table.restype = C.c_int
        L.orc_py_sum.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.orc_py_sum.restype = C.c_double
        _lib = L
    return _lib


def _item_kind(x) -> int:
    if type(x) is float:
        return 1
    if isinstance(x, (int, np.integer)) and not isinstance(x, bool) and x == 0:
        return 0
    return 2  # np.float64 & co: CPython sum() leaves its compensated float loop


def make_params(cfg: dict, mc: dict, pop_no_medium: str = "SOFT", pop_no_soft: str = "MEDIUM",
                drivers: list[str] | None = None) -> OrcParams:
    """cfg = RaceConfig kwargs (src/simulation.py:39-52); mc = run_monte_carlo kwargs (:59-69)."""
    gp = mc["grid_probs"]
    D = list(gp.keys()) if drivers is None else list(drivers)
    n = len(D)
    if not 1 <= n <= MAXD:
        raise ValueError(f"oracle supports 1..{MAXD} drivers, got {n}")
    p = OrcParams()
    p.n_drivers, p.total_laps = n, int(cfg["total_laps"])
    p.track_condition = TRACK[mc.get("track_condition", "dry")]
    p.pop_no_medium = COMPOUNDS.index(pop_no_medium)
    p.pop_no_soft = COMPOUNDS.index(pop_no_soft)
    p.pit_loss, p.overtake_delta = cfg["pit_loss"], cfg["overtake_delta"]
    p.sc_p, p.vsc_p, p.red_p = cfg["sc_probability"], cfg["vsc_probability"], cfg["red_flag_probability"]
    p.drs_delta = cfg["drs_delta"]
    p.dirty_thr = cfg.get("dirty_air_threshold", 2.0)
    p.dirty_pen = cfg.get("dirty_air_penalty", 0.5)
    for k, name in enumerate(COMPOUNDS):
        info = cfg["tire_compounds"].get(name, {})
        p.compound_pace_delta[k] = info.get("pace_delta", 0)
        p.compound_deg_rate[k] = info.get("deg_rate", 0.05)
        p.compound_optimal[k] = info.get("optimal_laps", 30)
    dnf = mc.get("driver_dnf_rates") or {}
    for i, d in enumerate(D):
        team = cfg["driver_teams"].get(d, "Unknown")
        team_rate = cfg["dnf_rates"].get(team, 0.002)
        p.base_pace[i] = mc["base_pace"].get(d, 90.0)
        p.tire_deg[i] = mc["tire_deg"].get(d, 0.05)
        p.tire_deg_pit[i] = mc["tire_deg"].get(d, 0.0)
        p.variance[i] = mc["driver_variance"].get(d, 0.15)
        p.dnf_rate[i] = dnf.get(d, team_rate)
        p.team_rate[i] = team_rate
        row = gp.get(d, [])
        for pos in range(n):
            if pos < len(row):
                p.grid_probs[i][pos] = float(row[pos])
                p.grid_kind[i][pos] = _item_kind(row[pos])
            else:
                p.grid_probs[i][pos] = 0.0
                p.grid_kind[i][pos] = 0
    return p


class Rng:
    """The reference process's two global MT streams (random + np.random), continuing across runs."""

    def __init__(self, seed: int | None = None):
        self.state = OrcRng()
        if seed is not None:
            self.seed(seed)

    def seed(self, seed: int):
        if not 0 <= seed < 2 ** 32:
            raise ValueError("Seed must be between 0 and 2**32 - 1")  # np.random.seed's own limit
        lib().orc_rng_seed(C.byref(self.state), seed)

    def py_random(self) -> float:
        return lib().orc_py_random(C.byref(self.state))

    def np_random_sample(self) -> float:
        return lib().orc_np_random_sample(C.byref(self.state))

    def np_standard_normal(self) -> float:
        return lib().orc_np_standard_normal(C.byref(self.state))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _alloc_outputs(n_sims: int, n: int, detail: bool, hist=None):
    o = {"hist": np.zeros((n, n), np.int64) if hist is None else hist}
    if detail:
        o.update(grid=np.zeros((n_sims, n), np.uint8), finish=np.zeros((n_sims, n), np.uint8),
                 times=np.zeros((n_sims, n), np.float64), dnf_lap=np.zeros((n_sims, n), np.int16),
                 draws=np.zeros((n_sims, 3), np.int64))
    return o


def run_streams(params: OrcParams, rng: Rng, n_sims: int, detail: bool = True, tapes: bool = False,
                hist=None) -> dict:
    """n_sims iterations of the reference's run_monte_carlo loop on the MT streams in `rng`."""
    n = params.n_drivers
    o = _alloc_outputs(n_sims, n, detail, hist)
    out = OrcOutputs()
    for k in ("hist", "grid", "finish", "times", "dnf_lap", "draws"):
        setattr(out, k, _ptr(o.get(k)))
    if tapes:
        # worst case per sim: n-independent bound from SURVEY draw-order spec
        L = params.total_laps
        cap_py = n_sims * (n + (L - 1) * (4 + n + 3 * (n - 1))) + 16
        cap_z = n_sims * (2 * n + (L - 1) * n) + 16
        cap_np = n_sims * n + 16
        o["tape_upy"], o["tape_z"], o["tape_unp"] = (np.zeros(c, np.float64) for c in (cap_py, cap_z, cap_np))
        out.log_upy, out.cap_upy = _ptr(o["tape_upy"]), cap_py
        out.log_z, out.cap_z = _ptr(o["tape_z"]), cap_z
        out.log_unp, out.cap_unp = _ptr(o["tape_unp"]), cap_np
    rc = lib().orc_run_streams(C.byref(params), C.byref(rng.state), n_sims, C.byref(out))
    if rc:
        raise RuntimeError(f"orc_run_streams failed: {rc}")
    if tapes:
        d = o["draws"]
        o["tape_upy"] = o["tape_upy"][: d[-1, 0]]
        o["tape_z"] = o["tape_z"][: d[-1, 1]]
        o["tape_unp"] = o["tape_unp"][: d[-1, 2]]
        off = np.zeros((n_sims + 1, 3), np.int64)
        off[1:] = d
        o["tape_off"] = off
    return o


def run_tapes(params: OrcParams, upy, z, unp, off, detail: bool = True) -> dict:
    """Replay explicit tapes.  `off` is (n_sims+1, 3) int64 offsets into upy / z / unp."""
    off = np.ascontiguousarray(off, np.int64)
    n_sims = off.shape[0] - 1
    n = params.n_drivers
    o = _alloc_outputs(n_sims, n, detail)
    out = OrcOutputs()
    for k in ("hist", "grid", "finish", "times", "dnf_lap", "draws"):
        setattr(out, k, _ptr(o.get(k)))
    upy, z, unp = (np.ascontiguousarray(a, np.float64) for a in (upy, z, unp))
    rc = lib().orc_run_tapes(C.byref(params), n_sims, _ptr(upy), _ptr(z), _ptr(unp), _ptr(off), C.byref(out))
    if rc:
        raise RuntimeError(f"orc_run_tapes failed: {rc}")
    return o


def py_sum(values, kinds) -> tuple[float, int]:
    v = np.ascontiguousarray(values, np.float64)
    k = np.ascontiguousarray(kinds, np.uint8)
    rk = C.c_int(0)
    r = lib().orc_py_sum(_ptr(v), _ptr(k), len(v), C.byref(rk))
    return r, rk.value


def run_monte_carlo(cfg: dict, mc: dict, n_sims: int, seed: int | None, pop_no_medium="SOFT",
                    pop_no_soft="MEDIUM", rng: Rng | None = None, threads: int = 1) -> np.ndarray:
    """Count table hist[driver, pos] of the reference's run_monte_carlo(n_sims, ..., seed).

    threads > 1 splits the sims into independent streams seeded seed+t (as BASELINE.md §3 does for
    the multi-core CPU baseline); that is no longer the single-stream reference result.
    """
    params = make_params(cfg, mc, pop_no_medium, pop_no_soft)
    if threads <= 1:
        rng = rng or Rng()
        if seed is not None:
            rng.seed(seed)
        return run_streams(params, rng, n_sims, detail=False)["hist"]
    from concurrent.futures import ThreadPoolExecutor
    per = [n_sims // threads + (1 if t < n_sims % threads else 0) for t in range(threads)]
    with ThreadPoolExecutor(threads) as ex:
        hs = list(ex.map(lambda t: run_streams(params, Rng((seed or 0) + t), per[t], detail=False)["hist"],
                         range(threads)))
    return np.sum(hs, axis=0)


TRACE_DTYPE = np.dtype([("position", np.uint8), ("compound", np.uint8), ("tire_age", np.uint8), ("flags", np.uint8),
                        ("gap", np.float32)])


def native_op32_table(params: OrcParams) -> np.ndarray:
    """The mirror's overtake-pace table [total_laps + 5, n] (float32); see native_mirror.h."""
    out = np.zeros((params.total_laps + 5, params.n_drivers), np.float32)
    rc = lib().orc_native_op32_table(C.byref(params), _ptr(out))
    if rc:
        raise RuntimeError(f"orc_native_op32_table failed: {rc}")
    return out


def run_native(params: OrcParams, seed: int, n_sims: int, sim_begin: int = 0, stream: int = 0, exact: bool = True,
               detail: bool = False, threads: int = 1, trace: bool = False) -> dict:
    """Scalar mirror of the native-mode kernel (oracle/native_mirror.c)."""
    n = params.n_drivers
    if trace:
        threads = 1

    def one(begin, count):
        o = {"hist": np.zeros((n, n), np.int64)}
        if detail:
            o["finish"] = np.zeros((count, n), np.uint8)
            o["times"] = np.zeros((count, n), np.float32)
        if trace:
            o["trace"] = np.zeros((count, params.total_laps, n), TRACE_DTYPE)
        rc = lib().orc_run_native(C.byref(params), seed & (2 ** 64 - 1), stream, begin, count, int(exact),
                                  _ptr(o["hist"]), _ptr(o.get("finish")), _ptr(o.get("times")), _ptr(o.get("trace")))
        if rc:
            raise RuntimeError(f"orc_run_native failed: {rc}")
        return o

    if threads <= 1 or n_sims < 4 * threads:
        return one(sim_begin, n_sims)
    from concurrent.futures import ThreadPoolExecutor
    bounds = [sim_begin + n_sims * t // threads for t in range(threads + 1)]
    with ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(lambda t: one(bounds[t], bounds[t + 1] - bounds[t]), range(threads)))
    out = {"hist": np.sum([q["hist"] for q in parts], axis=0)}
    if detail:
        out["finish"] = np.concatenate([q["finish"] for q in parts])
        out["times"] = np.concatenate([q["times"] for q in parts])
    return out
