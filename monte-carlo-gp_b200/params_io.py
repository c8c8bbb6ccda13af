"""On-disk format for race parameter blocks.

The reference derives a race's simulation inputs from FastF1 data with pandas (``src/predictor.py:409-569``); that
extraction is out of scope here, but its RESULT -- the dense ``mcgp_race_params`` block the simulator consumes
(include/mcgp.h) plus the driver names -- is worth keeping: a saved block replays a race weekend's
simulation on any box without FastF1, the network or the cache.  Format: one ``.npz`` with every field of the C struct
as an array stacked over races (scalars -> [n_races], vectors -> [n_races, 32], the grid -> [n_races, 32, 32]) and a
JSON ``meta`` entry (format version, driver names per race, free-form labels).
"""
from __future__ import annotations

import ctypes as C
import json

import numpy as np

from . import capi

FORMAT = "mcgp-race-params/1"


def params_to_arrays(params: list) -> dict:
    """Stack the fields of a list of McgpRaceParams into numpy arrays (one leading axis over races)."""
    out = {}
    for name, ctype in capi.McgpRaceParams._fields_:
        vals = [np.ctypeslib.as_array(getattr(p, name)).copy() if issubclass(ctype, C.Array) else getattr(p, name) for p in params]
        out[name] = np.stack([np.asarray(v) for v in vals])
    return out


def arrays_to_params(arrays: dict) -> list:
    n_races = len(arrays["n_drivers"])
    params = []
    for r in range(n_races):
        p = capi.McgpRaceParams()
        for name, ctype in capi.McgpRaceParams._fields_:
            v = arrays[name][r]
            if issubclass(ctype, C.Array):
                dst = np.ctypeslib.as_array(getattr(p, name))
                if dst.shape != np.shape(v):
                    raise ValueError(f"field {name}: expected shape {dst.shape}, file has {np.shape(v)}")
                dst[...] = v
            else:
                setattr(p, name, v.item())
        params.append(p)
    return params


def save_race_params(path: str, params: list, drivers: list[list[str]] | None = None, labels: list[str] | None = None) -> None:
    """Write parameter blocks (e.g. ``RaceSimulator._params(...)`` of each race of a weekend / season) to `path`."""
    meta = {"format": FORMAT, "n_races": len(params), "drivers": drivers, "labels": labels}
    np.savez_compressed(path, meta=json.dumps(meta), **params_to_arrays(list(params)))


def load_race_params(path: str) -> tuple[list, dict]:
    """Returns (list of McgpRaceParams, meta).  Raises ValueError on a foreign or truncated file."""
    with np.load(path, allow_pickle=False) as z:
        if "meta" not in z.files:
            raise ValueError("not a race-parameter file: no meta entry")
        meta = json.loads(str(z["meta"]))
        if meta.get("format") != FORMAT:
            raise ValueError(f"unsupported format {meta.get('format')!r}, expected {FORMAT!r}")
        missing = [n for n, _ in capi.McgpRaceParams._fields_ if n not in z.files]
        if missing:
            raise ValueError(f"race-parameter file lacks fields {missing}")
        arrays = {n: z[n] for n, _ in capi.McgpRaceParams._fields_}
    return arrays_to_params(arrays), meta
