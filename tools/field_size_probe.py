#!/usr/bin/env python
"""Throughput of the native kernel for other field sizes (the BASELINE metric is quoted on 20 cars): races/s and
driver-laps/s for n = 10, 16, 20, 24, 32 cars on the Bahrain-like 57-lap race.  One warp simulates one race whatever n
is, so driver-laps/s shows what the 12 idle lanes of a 20-car race cost.  usage: python tools/field_size_probe.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mcgp_b200 as m

wl = m.workloads
out = {}
for n in (10, 16, 20, 24, 32):
    D = [f"D{i:02d}" for i in range(n)]
    cfg, _ = wl.workload("bahrain")
    cfg["driver_teams"] = {d: "Unknown" for d in D}
    mc = dict(grid_probs=wl.gaussian_grid_probs(D), base_pace={d: 92.0 + 0.07 * k for k, d in enumerate(D)},
              tire_deg={d: 0.015 + 0.003 * (k % 20) for k, d in enumerate(D)},
              driver_variance={d: 0.12 + 0.005 * (k % 5) for k, d in enumerate(D)},
              driver_dnf_rates={d: 0.05 / 57 for d in D})
    sim = m.simulation.RaceSimulator(m.simulation.RaceConfig(**cfg), device=0, pop_no_medium="SOFT", pop_no_soft="MEDIUM")
    p = sim._params(mc["grid_probs"], mc["base_pace"], mc["tire_deg"], mc["driver_variance"], mc["driver_dnf_rates"], "dry")
    eng = m.capi.get_engine(0)
    eng.upload_races([p])
    S = 4_000_000
    h = torch.zeros((1, n, n), dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    eng.launch_native(S, 0, 1, h.data_ptr(), stream=st); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(3):
        eng.launch_native(S, (k + 1) * S, 1, h.data_ptr(), stream=st)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    assert int(h.sum()) == 4 * S * n
    out[n] = {"races_per_s": S / (ms * 1e-3), "driver_laps_per_s": S * n * 57 / (ms * 1e-3)}
    print(n, "cars:", f"{out[n]['races_per_s'] / 1e6:.1f} M races/s, {out[n]['driver_laps_per_s'] / 1e9:.1f} G driver-laps/s", flush=True)
print(json.dumps(out))
