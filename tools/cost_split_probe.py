#!/usr/bin/env python
"""Where the native kernel's time goes, measured by switching model features off in the INPUTS (no code changes):
races/s of the Bahrain-57 workload with (a) everything on, (b) no pair ever eligible to overtake, (c) no pit stops,
(d) neither, (e) additionally no events / retirements, (f) a fixed grid (no grid sampling).  usage: python tools/cost_split_probe.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mcgp_b200 as m

wl = m.workloads


def rate(cfg, mc, S=4_000_000):
    sim = m.simulation.RaceSimulator(m.simulation.RaceConfig(**cfg), device=0, pop_no_medium="SOFT", pop_no_soft="MEDIUM")
    p = sim._params(mc["grid_probs"], mc["base_pace"], mc["tire_deg"], mc["driver_variance"], mc["driver_dnf_rates"], "dry")
    eng = m.capi.get_engine(0)
    eng.upload_races([p])
    h = torch.zeros((1, 20, 20), dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    eng.launch_native(S, 0, 1, h.data_ptr(), stream=st); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(3):
        eng.launch_native(S, (k + 1) * S, 1, h.data_ptr(), stream=st)
    b.record(); torch.cuda.synchronize()
    return S / (a.elapsed_time(b) / 3 * 1e-3)


def variant(no_ovt=False, no_pit=False, no_ev=False, fixed=False):
    cfg, mc = wl.workload("bahrain")
    if no_ovt:
        cfg["overtake_delta"] = 1e9
    if no_pit:
        cfg["tire_compounds"] = {k: dict(v, optimal_laps=1000) for k, v in cfg["tire_compounds"].items()}
    if no_ev:
        cfg["sc_probability"] = cfg["vsc_probability"] = cfg["red_flag_probability"] = 0.0
        cfg["dnf_rates"] = {t: 0.0 for t in cfg["dnf_rates"]}
        mc["driver_dnf_rates"] = {d: 0.0 for d in mc["driver_dnf_rates"]}
    if fixed:
        mc["grid_probs"] = wl.onehot_grid_probs(list(mc["grid_probs"]))
    return rate(cfg, mc)


out = {"all on": variant(), "no overtakes": variant(no_ovt=True), "no pit stops": variant(no_pit=True),
       "no overtakes, no pits": variant(no_ovt=True, no_pit=True),
       "no overtakes, pits, events, retirements": variant(no_ovt=True, no_pit=True, no_ev=True),
       "fixed grid": variant(fixed=True),
       "fixed grid, nothing else either": variant(no_ovt=True, no_pit=True, no_ev=True, fixed=True)}
for k, v in out.items():
    print(f"{k:45s} {v / 1e6:7.1f} M races/s   {1e9 * 148 * 4 * 1.965 / v / 57:6.0f} issue slots per race-lap")
print(json.dumps(out))
