#!/usr/bin/env python
"""A/B harness for kernel builds: times the headline launch (Bahrain-57, 20 drivers, native mode) of the library
selected with MCGP_LIB_PATH and, with --check, holds it bit for bit against the scalar mirror (exact-normal mode).

    MCGP_LIB_PATH=monte-carlo-gp_b200/libmcgp_v1.so python tools/ab_bench.py [--sims 4000000] [--reps 3] [--check] [--tag v1]

Prints one JSON line.  Test infrastructure (imports oracle/ for --check); not part of the product path."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sims", type=int, default=4_000_000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--workload", default="bahrain")
    ap.add_argument("--mode", default="counts", choices=["counts", "trace", "laphist"],
                    help="which kernel variant the timed launches use: count table only, + per-lap trace, + lap histogram")
    ap.add_argument("--tag", default=os.path.basename(os.environ.get("MCGP_LIB_PATH", "libmcgp.so")))
    args = ap.parse_args()
    import numpy as np
    import torch
    import mcgp_b200
    cfg, mc = mcgp_b200.workloads.workload(args.workload)
    sim = mcgp_b200.simulation.RaceSimulator(mcgp_b200.simulation.RaceConfig(**cfg), device=0, pop_no_medium="SOFT", pop_no_soft="MEDIUM")
    mc_args = (mc["grid_probs"], mc["base_pace"], mc["tire_deg"], mc["driver_variance"], mc["driver_dnf_rates"])
    p = sim._params(*mc_args, mc["track_condition"])
    eng = mcgp_b200.capi.Engine(0)
    out = {"tag": args.tag}
    if args.check:
        from oracle import pyoracle as po
        op = po.make_params(cfg, mc, "SOFT", "MEDIUM")
        n_chk = 4096
        hist, finish, times = eng.run_native([p], n_chk, 0, 42, flags=mcgp_b200.capi.F_EXACT_NORMAL, want_finish=True, want_times=True)
        m = po.run_native(op, 42, n_chk, exact=True, detail=True)
        out["mirror_orders_equal"] = bool(np.array_equal(finish[0], m["finish"]))
        out["mirror_times_equal"] = bool(np.array_equal(times[0].view(np.uint32), m["times"].view(np.uint32)))
        out["mirror_mismatching_races"] = int((finish[0] != m["finish"]).any(axis=1).sum())
    eng.upload_races([p])
    n = p.n_drivers
    hist = torch.zeros((1, n, n), dtype=torch.int64, device="cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    laps = p.total_laps
    extra = None
    if args.mode == "trace":
        extra = torch.empty(args.sims * laps * n * 8, dtype=torch.uint8, device="cuda:0")
    elif args.mode == "laphist":
        extra = torch.zeros((1, laps, n, n), dtype=torch.int64, device="cuda:0")

    def launch(begin):
        if args.mode == "trace":
            eng.launch_native_traced(args.sims, begin, 42, hist.data_ptr(), extra.data_ptr(), 0, args.sims, stream=st)
        elif args.mode == "laphist":
            eng.launch_native_laphist(args.sims, begin, 42, hist.data_ptr(), extra.data_ptr(), stream=st)
        else:
            eng.launch_native(args.sims, begin, 42, hist.data_ptr(), stream=st)
    for _ in range(2):
        launch(0)
    torch.cuda.synchronize()
    best = 1e30
    for r in range(args.reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        launch((r + 3) * args.sims)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    out.update(mode=args.mode, races_per_s=args.sims / (best * 1e-3), ms=best, sims=args.sims,
               hist_ok=int(hist.sum().item()) == (args.reps + 2) * args.sims * n)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
