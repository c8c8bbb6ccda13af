#!/usr/bin/env python
"""Replay-kernel timing on synthetic worst-case-sized tapes (the same set-up as bench.py's replay_mode leg), for
A/B builds (MCGP_LIB_PATH) and ncu captures:  python tests/checkers/replay_bench.py [--sims 100000] [--reps 3]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sims", type=int, default=100_000)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import numpy as np
    import torch
    import mcgp_b200
    N, LAPS = 20, 57
    cfg, mc = mcgp_b200.workloads.workload("bahrain")
    sim = mcgp_b200.simulation.RaceSimulator(mcgp_b200.simulation.RaceConfig(**cfg), device=0, pop_no_medium="SOFT", pop_no_soft="MEDIUM")
    p = sim._params(mc["grid_probs"], mc["base_pace"], mc["tire_deg"], mc["driver_variance"], mc["driver_dnf_rates"], mc["track_condition"])
    eng = mcgp_b200.capi.Engine(0)
    eng.upload_races([p])
    dev = torch.device("cuda:0")
    n_rep = args.sims
    n_py = N + (LAPS - 1) * (4 + N + 3 * (N - 1))
    n_z = 2 * N + (LAPS - 1) * N
    g = torch.Generator(device=dev).manual_seed(42)
    tapes = [torch.rand(n_rep * n_py, dtype=torch.float64, device=dev, generator=g),
             torch.randn(n_rep * n_z, dtype=torch.float64, device=dev, generator=g),
             torch.rand(n_rep * N, dtype=torch.float64, device=dev, generator=g)]
    off = torch.from_numpy(np.arange(n_rep + 1, dtype=np.int64)[:, None] * np.array([n_py, n_z, N], np.int64)).contiguous().to(dev)
    rh = torch.zeros((N, N), dtype=torch.int64, device=dev)
    used = torch.zeros((n_rep, 3), dtype=torch.int64, device=dev)
    status = torch.zeros(4, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def go():
        eng.launch_replay(n_rep, tapes[0].data_ptr(), tapes[1].data_ptr(), tapes[2].data_ptr(), off.data_ptr(),
                          rh.data_ptr(), used_ptr=used.data_ptr(), status_ptr=status.data_ptr(), stream=st)
    for _ in range(2):
        go()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.reps):
        go()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.reps
    consumed = int(used.sum().item()) * 8
    print(json.dumps({"tag": os.path.basename(os.environ.get("MCGP_LIB_PATH", "libmcgp.so")), "replay_races_per_s": n_rep / (ms * 1e-3),
                      "ms": ms, "sims": n_rep, "tape_bytes_per_race": consumed / n_rep, "tape_gb_per_s": consumed / (ms * 1e-3) / 1e9,
                      "status": int(status[0].item()), "count_table_ok": int(rh.sum().item()) == (args.reps + 2) * n_rep * N,
                      "win_counts_head": rh[:3, 0].tolist()}))


if __name__ == "__main__":
    main()
