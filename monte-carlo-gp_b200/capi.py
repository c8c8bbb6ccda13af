"""ctypes binding of libmcgp.so (include/mcgp.h) -- the thin layer between the Python host code and CUDA.

No torch here: buffers cross the ABI as raw pointers (numpy arrays for the host-buffer entry points,
``tensor.data_ptr()`` integers for the device-resident ones).  There is no CPU fallback: if the shared
library is missing or no B200 is visible, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCGP_LIB_PATH") or os.path.join(PKG_DIR, "libmcgp.so")  # override: A/B builds of the kernel
MAX_DRIVERS, N_COMPOUNDS = 32, 5
COMPOUNDS = ("SOFT", "MEDIUM", "HARD", "INTERMEDIATE", "WET")
TRACK_CONDITIONS = {"dry": 0, "damp": 1, "wet": 2}
ITEM_INT0, ITEM_FLOAT, ITEM_NPFLOAT = 0, 1, 2
F_EXACT_NORMAL = 1

OK, EINVAL, ENODEVICE, ECUDA, ETAPE, ENOMEM = 0, -1, -2, -3, -4, -5


class McgpRaceParams(C.Structure):
    """mcgp_race_params (include/mcgp.h)."""
    _fields_ = [
        ("n_drivers", C.c_int32), ("total_laps", C.c_int32), ("track_condition", C.c_int32),
        ("pop_no_medium", C.c_int32), ("pop_no_soft", C.c_int32), ("stream", C.c_uint32),
        ("pit_loss", C.c_double), ("overtake_delta", C.c_double),
        ("sc_probability", C.c_double), ("vsc_probability", C.c_double), ("red_flag_probability", C.c_double),
        ("drs_delta", C.c_double), ("dirty_air_threshold", C.c_double), ("dirty_air_penalty", C.c_double),
        ("compound_pace_delta", C.c_double * N_COMPOUNDS), ("compound_deg_rate", C.c_double * N_COMPOUNDS),
        ("compound_optimal_laps", C.c_double * N_COMPOUNDS),
        ("base_pace", C.c_double * MAX_DRIVERS), ("tire_deg", C.c_double * MAX_DRIVERS),
        ("tire_deg_pit", C.c_double * MAX_DRIVERS), ("driver_variance", C.c_double * MAX_DRIVERS),
        ("dnf_rate", C.c_double * MAX_DRIVERS), ("team_dnf_rate", C.c_double * MAX_DRIVERS),
        ("grid_probs", (C.c_double * MAX_DRIVERS) * MAX_DRIVERS),
        ("grid_kind", (C.c_uint8 * MAX_DRIVERS) * MAX_DRIVERS),
    ]


# mcgp_trace_record (include/mcgp.h): one record per (sim, lap, driver)
TRACE_DTYPE = np.dtype([("position", np.uint8), ("compound", np.uint8), ("tire_age", np.uint8), ("flags", np.uint8),
                        ("gap", np.float32)])


class McgpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libmcgp error {code}: {msg}")
        self.code = code
        self.msg = msg


_lib = None

_SIGNATURES = {
    "mcgp_abi_version": (C.c_int, []),
    "mcgp_native_philox_rounds": (C.c_int, []),
    "mcgp_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "mcgp_destroy": (C.c_int, [C.c_void_p]),
    "mcgp_last_error": (C.c_char_p, [C.c_void_p]),
    "mcgp_device_info": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_int)] * 4),
    "mcgp_last_launch_count": (C.c_int, [C.c_void_p]),
    "mcgp_last_upload_bytes": (C.c_uint64, [C.c_void_p]),
    "mcgp_pace_table": (C.c_int, [C.POINTER(McgpRaceParams), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p]),
    "mcgp_upload_races": (C.c_int, [C.c_void_p, C.POINTER(McgpRaceParams), C.c_int]),
    "mcgp_run_native": (C.c_int, [C.c_void_p, C.POINTER(McgpRaceParams), C.c_int, C.c_uint64, C.c_uint64,
                                  C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "mcgp_launch_native": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "mcgp_launch_native_traced": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p,
                                            C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "mcgp_run_native_traced": (C.c_int, [C.c_void_p, C.POINTER(McgpRaceParams), C.c_int, C.c_uint64, C.c_uint64,
                                         C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]),
    "mcgp_lap_histogram_laps": (C.c_int, [C.c_void_p]),
    "mcgp_launch_native_laphist": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p,
                                             C.c_void_p, C.c_void_p]),
    "mcgp_run_native_laphist": (C.c_int, [C.c_void_p, C.POINTER(McgpRaceParams), C.c_int, C.c_uint64, C.c_uint64,
                                          C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p]),
    "mcgp_score_counts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint64] + [C.c_void_p] * 8),
    "mcgp_run_season": (C.c_int, [C.c_void_p, C.POINTER(McgpRaceParams), C.c_int, C.c_uint64, C.c_uint64, C.c_uint32,
                                  C.c_double] + [C.c_void_p] * 14),
    "mcgp_run_replay": (C.c_int, [C.c_void_p, C.POINTER(McgpRaceParams), C.c_uint64] + [C.c_void_p] * 10),
    "mcgp_launch_replay": (C.c_int, [C.c_void_p, C.c_uint64] + [C.c_void_p] * 12),
    "mcgp_replay_serial_grid": (C.c_int, [C.c_void_p, C.c_int]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def load_library():
    """dlopen libmcgp.so; fails loudly -- there is no fallback implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C monte-carlo-gp_b200/csrc`). "
                "This engine has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            if not hasattr(lib, name) and os.environ.get("MCGP_LIB_PATH"):
                continue  # an A/B build of an older kernel may lack newer entry points; the shipped library may not
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.mcgp_abi_version() != 1:
            raise RuntimeError("libmcgp.so ABI version mismatch")
        _lib = lib
    return _lib


def _p(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))  # raw device pointer (e.g. tensor.data_ptr())


class Engine:
    """One mcgp context (= one GPU).  Calls on one Engine are serialised by the caller."""

    def __init__(self, device: int = 0):
        lib = load_library()
        h = C.c_void_p()
        rc = lib.mcgp_create(C.byref(h), int(device))
        if rc:
            raise McgpError(rc, lib.mcgp_last_error(None).decode())
        self._lib, self._h, self.device = lib, h, int(device)
        self.n_races = self.n_drivers = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mcgp_destroy(self._h)
            self._h = None

    __del__ = close

    def _check(self, rc: int):
        if rc:
            raise McgpError(rc, self._lib.mcgp_last_error(self._h).decode())

    def device_info(self) -> dict:
        v = [C.c_int() for _ in range(4)]
        self._check(self._lib.mcgp_device_info(self._h, *[C.byref(x) for x in v]))
        return dict(sm_count=v[0].value, sm_clock_khz=v[1].value, cc=(v[2].value, v[3].value))

    @property
    def last_launch_count(self) -> int:
        return self._lib.mcgp_last_launch_count(self._h)

    def last_upload_bytes(self) -> int:
        return int(self._lib.mcgp_last_upload_bytes(self._h))

    @staticmethod
    def _pack(races) -> tuple:
        races = list(races)
        arr = (McgpRaceParams * len(races))(*races)
        return arr, len(races), races[0].n_drivers

    # ---- native mode ---------------------------------------------------------------------------
    def run_native(self, races, n_sims: int, sim_begin: int = 0, seed: int = 0, flags: int = 0,
                   want_finish: bool = False, hist: np.ndarray | None = None, want_times: bool = False):
        """Host-buffer call (parameters in, counts out, synchronous).  Returns hist[n_races, n, n] uint64;
        with want_finish also finish[n_races, n_sims, n] uint8; with want_times also times[...] float32."""
        arr, n_races, n = self._pack(races)
        if hist is None:
            hist = np.zeros((n_races, n, n), np.uint64)
        assert hist.dtype == np.uint64 and hist.shape == (n_races, n, n) and hist.flags.c_contiguous
        finish = np.zeros((n_races, n_sims, n), np.uint8) if want_finish else None
        times = np.zeros((n_races, n_sims, n), np.float32) if want_times else None
        self._check(self._lib.mcgp_run_native(self._h, arr, n_races, n_sims, sim_begin, seed & (2 ** 64 - 1), flags,
                                              _p(hist), _p(finish), _p(times)))
        self.n_races, self.n_drivers = n_races, n
        if want_finish or want_times:
            return tuple(x for x in (hist, finish, times) if x is not None)
        return hist

    def run_native_traced(self, races, n_sims: int, sim_begin: int = 0, seed: int = 0, flags: int = 0,
                          trace_first: int = 0, trace_count: int | None = None):
        """Host-buffer call with the per-lap trace of sims [trace_first, trace_first+trace_count) of the range.
        Returns (hist, trace) with trace a structured array [n_races, trace_count, laps, n] of TRACE_DTYPE."""
        arr, n_races, n = self._pack(races)
        trace_count = n_sims - trace_first if trace_count is None else trace_count
        laps = races[0].total_laps
        hist = np.zeros((n_races, n, n), np.uint64)
        trace = np.zeros((n_races, trace_count, laps, n), TRACE_DTYPE)
        self._check(self._lib.mcgp_run_native_traced(self._h, arr, n_races, n_sims, sim_begin, seed & (2 ** 64 - 1), flags,
                                                     _p(hist), _p(trace), trace_first, trace_count))
        self.n_races, self.n_drivers = n_races, n
        return hist, trace

    def run_native_laphist(self, races, n_sims: int, sim_begin: int = 0, seed: int = 0, flags: int = 0):
        """Host-buffer call that also returns the per-lap position histogram: (hist [n_races, n, n],
        laphist [n_races, laps, n, n]) with laphist[r, lap-1, d, pos] = sims in which driver d RUNS in position pos
        (0 = leading) after lap `lap`; laps = the longest race of the batch (include/mcgp.h: mcgp_run_native_laphist)."""
        arr, n_races, n = self._pack(races)
        laps = max(r.total_laps for r in races)
        hist = np.zeros((n_races, n, n), np.uint64)
        laphist = np.zeros((n_races, laps, n, n), np.uint64)
        self._check(self._lib.mcgp_run_native_laphist(self._h, arr, n_races, n_sims, sim_begin, seed & (2 ** 64 - 1), flags,
                                                      _p(hist), _p(laphist)))
        self.n_races, self.n_drivers = n_races, n
        assert self._lib.mcgp_lap_histogram_laps(self._h) == laps
        return hist, laphist

    def launch_native_laphist(self, n_sims, sim_begin, seed, hist_ptr, laphist_ptr, flags=0, stream=None):
        self._check(self._lib.mcgp_launch_native_laphist(self._h, n_sims, sim_begin, seed & (2 ** 64 - 1), flags, _p(hist_ptr),
                                                         _p(laphist_ptr), _p(stream)))

    def launch_native_traced(self, n_sims, sim_begin, seed, hist_ptr, trace_ptr, trace_first, trace_count, flags=0,
                             stream=None):
        self._check(self._lib.mcgp_launch_native_traced(self._h, n_sims, sim_begin, seed & (2 ** 64 - 1), flags, _p(hist_ptr),
                                                        _p(trace_ptr), trace_first, trace_count, _p(stream)))

    def upload_races(self, races):
        arr, n_races, n = self._pack(races)
        self._check(self._lib.mcgp_upload_races(self._h, arr, n_races))
        self.n_races, self.n_drivers = n_races, n

    def launch_native(self, n_sims: int, sim_begin: int, seed: int, hist_ptr: int, flags: int = 0,
                      finish_ptr: int | None = None, times_ptr: int | None = None, stream: int | None = None):
        """Asynchronous launch on device-resident buffers (raw device pointers)."""
        self._check(self._lib.mcgp_launch_native(self._h, n_sims, sim_begin, seed & (2 ** 64 - 1), flags, _p(hist_ptr),
                                                 _p(finish_ptr), _p(times_ptr), _p(stream)))

    # ---- scoring / device-resident season ------------------------------------------------------
    def score_counts(self, hist_ptr, n_races: int, n: int, n_sims: int, winners, podiums=None, stream=None) -> dict:
        """Scores count tables that live on the GPU (include/mcgp.h: mcgp_score_counts); only the results come back."""
        winners = np.ascontiguousarray(winners, np.int32)
        podiums = None if podiums is None else np.ascontiguousarray(podiums, np.int32).reshape(n_races, 3)
        out = dict(tallies=np.zeros((n_races, 3, n), np.uint64), brier=np.zeros(n_races, np.float64),
                   podium_hits=np.zeros(n_races, np.int32), calib=np.zeros((3, 10), np.float64), calib_bins=np.zeros(1, np.int32))
        self._check(self._lib.mcgp_score_counts(self._h, _p(hist_ptr), n_races, n, n_sims, _p(winners), _p(podiums), _p(out["tallies"]),
                                                _p(out["brier"]), _p(out["podium_hits"]), _p(out["calib"]), _p(out["calib_bins"]),
                                                _p(stream)))
        return out

    def run_season(self, races, n_sims: int, seed: int, quali0, race0, k_factor: float = 32.0, penalties=None,
                   flags: int = 0) -> dict:
        """The device-resident season loop (include/mcgp.h: mcgp_run_season)."""
        arr, R, n = self._pack(races)
        q0, r0 = np.ascontiguousarray(quali0, np.float64), np.ascontiguousarray(race0, np.float64)
        assert q0.shape == (n,) and r0.shape == (n,)
        pen = None if penalties is None else np.ascontiguousarray(penalties, np.int32).reshape(R, n)
        out = dict(hist=np.zeros((R, n, n), np.uint64), quali=np.zeros((R + 1, n)), race=np.zeros((R + 1, n)),
                   grid_rows=np.zeros((R, n, n)), actual_grid=np.zeros((R, n), np.uint8), actual_finish=np.zeros((R, n), np.uint8),
                   tallies=np.zeros((R, 3, n), np.uint64), brier=np.zeros(R), podium_hits=np.zeros(R, np.int32),
                   calib=np.zeros((3, 10)), calib_bins=np.zeros(1, np.int32))
        self._check(self._lib.mcgp_run_season(
            self._h, arr, R, n_sims, seed & (2 ** 64 - 1), flags, float(k_factor), _p(q0), _p(r0), _p(pen), _p(out["hist"]),
            _p(out["quali"]), _p(out["race"]), _p(out["grid_rows"]), _p(out["actual_grid"]), _p(out["actual_finish"]),
            _p(out["tallies"]), _p(out["brier"]), _p(out["podium_hits"]), _p(out["calib"]), _p(out["calib_bins"])))
        self.n_races = self.n_drivers = 0
        return out

    # ---- replay mode ---------------------------------------------------------------------------
    def run_replay(self, race: McgpRaceParams, u_py, z, u_np, off, detail: bool = True) -> dict:
        """Host-buffer replay of explicit tapes; `off` is (n_sims+1, 3) int64."""
        off = np.ascontiguousarray(off, np.int64)
        n_sims, n = off.shape[0] - 1, race.n_drivers
        u_py, z, u_np = (np.ascontiguousarray(a, np.float64) for a in (u_py, z, u_np))
        out = {"hist": np.zeros((n, n), np.uint64)}
        if detail:
            out.update(finish=np.zeros((n_sims, n), np.uint8), times=np.zeros((n_sims, n), np.float64),
                       dnf_lap=np.zeros((n_sims, n), np.int16), grid=np.zeros((n_sims, n), np.uint8),
                       used=np.zeros((n_sims, 3), np.int64))
        arr = (McgpRaceParams * 1)(race)
        self._check(self._lib.mcgp_run_replay(
            self._h, arr, n_sims, _p(u_py), _p(z), _p(u_np), _p(off), _p(out["hist"]), _p(out.get("finish")),
            _p(out.get("times")), _p(out.get("dnf_lap")), _p(out.get("grid")), _p(out.get("used"))))
        self.n_races, self.n_drivers = 1, n
        return out

    def replay_serial_grid(self, on: bool):
        """Send every _sample_grid position of later replay launches down the serial (reference operation order) path;
        by default the kernel takes it only where its parallel evaluation cannot certify the same selection."""
        self._check(self._lib.mcgp_replay_serial_grid(self._h, 1 if on else 0))

    def launch_replay(self, n_sims, u_py_ptr, z_ptr, u_np_ptr, off_ptr, hist_ptr, finish_ptr=None, times_ptr=None,
                      dnf_lap_ptr=None, grid_ptr=None, used_ptr=None, status_ptr=None, stream=None):
        self._check(self._lib.mcgp_launch_replay(
            self._h, n_sims, _p(u_py_ptr), _p(z_ptr), _p(u_np_ptr), _p(off_ptr), _p(hist_ptr), _p(finish_ptr),
            _p(times_ptr), _p(dnf_lap_ptr), _p(grid_ptr), _p(used_ptr), _p(status_ptr), _p(stream)))


_engines: dict[int, Engine] = {}


def pace_table(race: McgpRaceParams) -> np.ndarray:
    """The overtake pace table the library derives for one race (host-only, needs no GPU): float32
    [rows = laps + 5, stride, 4] with entries {f(P[d][age]), thr_no_drs, thr_drs, 0} (include/mcgp.h: mcgp_pace_table)."""
    lib = load_library()
    rows, stride = C.c_int32(0), C.c_int32(0)
    rc = lib.mcgp_pace_table(C.byref(race), C.byref(rows), C.byref(stride), None)
    if rc:
        raise McgpError(rc, "mcgp_pace_table: invalid race parameters")
    out = np.zeros((rows.value, stride.value, 4), np.float32)
    rc = lib.mcgp_pace_table(C.byref(race), C.byref(rows), C.byref(stride), _p(out))
    if rc:
        raise McgpError(rc, "mcgp_pace_table failed")
    return out


def get_engine(device: int = 0) -> Engine:
    """Process-wide Engine per device (contexts are cheap but their scratch buffers are worth keeping)."""
    if device not in _engines:
        _engines[device] = Engine(device)
    return _engines[device]
