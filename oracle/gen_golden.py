"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the UNMODIFIED reference.

    python oracle/gen_golden.py            # all cases, in parallel, PYTHONHASHSEED pinned per case
    python oracle/gen_golden.py --case sprint19 --hashseed 1

Runs only in the build container (needs /root/reference).  Every fixture stores, for its case
(tests/golden_cases.py): the 20x20 finish-position count table over all sims, and for the first
DETAIL_SIMS sims the grid, finishing order, final cumulative times, DNF laps and cumulative draw
counts; for the first TAPE_SIMS sims also the raw draws.  `meta` records the interpreter's
`set.pop()` choices so the consumers pass the matching knobs.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# (case, PYTHONHASHSEED, fixture suffix, sims override)
PLAN = [(c, 0, "", None) for c in (
    "bahrain_dry", "monaco_sc", "sprint19", "damp", "wet", "onehot", "npfloat_grid", "defaults",
    "attrition", "wipeout", "events", "tight", "small_grids", "single", "one_lap", "canada70")]
# the other `set.pop()` outcomes, as probed from the reference in the generating process:
# PYTHONHASHSEED=0 -> (SOFT, MEDIUM); 1 -> (HARD, MEDIUM); 7 -> (HARD, HARD)
PLAN += [("bahrain_dry", 1, "_h1", 10000), ("sprint19", 7, "_h7", None)]


def run_one(case: str, suffix: str, n_override):
    import numpy as np
    import golden_cases as gc
    from oracle import ref_record

    ref_record.self_check()
    cfg, mc, seed, n_sims = gc.get_case(case)
    if n_override:
        n_sims = n_override
    t0 = time.perf_counter()
    rec = ref_record.record(cfg, mc, seed, n_sims, tape_sims=gc.TAPE_SIMS)
    dt = time.perf_counter() - t0
    pop_no_medium, pop_no_soft = ref_record.reference_pop_choices()
    n = rec["hist"].shape[0]
    canon = json.dumps([[int(rec["hist"][d, p]) for p in range(n)] for d in range(n)])
    meta = dict(case=case, seed=seed, n_sims=n_sims, drivers=list(mc["grid_probs"].keys()),
                pop_no_medium=pop_no_medium, pop_no_soft=pop_no_soft,
                hashseed=os.environ.get("PYTHONHASHSEED"), python=sys.version.split()[0],
                numpy=np.__version__, hist_sha256=hashlib.sha256(canon.encode()).hexdigest(),
                ref_wall_s=round(dt, 2), ref_sims_per_s=round(n_sims / dt, 1))
    k = min(gc.DETAIL_SIMS, n_sims)
    out = {key: rec[key][:k] for key in ("grid", "finish", "times", "dnf_lap", "draws")}
    out["hist"] = rec["hist"]
    for key in ("tape_upy", "tape_z", "tape_unp"):
        out[key] = rec[key]
    out["meta"] = np.array(json.dumps(meta))
    path = os.path.join(ROOT, "tests", "golden", f"{case}{suffix}.npz")
    np.savez_compressed(path, **out)
    print(f"{case}{suffix}: {n_sims} sims {dt:.1f}s sha={meta['hist_sha256'][:16]} pop=({pop_no_medium},{pop_no_soft})",
          flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case")
    ap.add_argument("--suffix", default="")
    ap.add_argument("--sims", type=int, default=0)
    ap.add_argument("--jobs", type=int, default=os.cpu_count())
    args = ap.parse_args()
    if args.case:
        run_one(args.case, args.suffix, args.sims or None)
        return
    procs = []
    pending = list(PLAN)
    while pending or procs:
        while pending and len(procs) < args.jobs:
            case, hs, suffix, n = pending.pop(0)
            env = dict(os.environ, PYTHONHASHSEED=str(hs))
            cmd = [sys.executable, os.path.abspath(__file__), "--case", case, "--suffix", suffix, "--sims", str(n or 0)]
            procs.append(subprocess.Popen(cmd, env=env, cwd=ROOT))
        for p in list(procs):
            if p.poll() is not None:
                if p.returncode:
                    raise SystemExit(f"golden generation failed: {p.args}")
                procs.remove(p)
        time.sleep(0.2)


if __name__ == "__main__":
    main()
