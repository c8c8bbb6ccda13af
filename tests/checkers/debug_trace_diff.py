"""Debug helper (GPU box): first per-lap trace record where the native kernel and the CPU mirror differ."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import golden_cases as gc
import mcgp_b200 as mcgp
from oracle import pyoracle as oracle

case = sys.argv[1] if len(sys.argv) > 1 else "bahrain_dry"
POP = ("SOFT", "MEDIUM")
MC_KEYS = ("grid_probs", "base_pace", "tire_deg", "driver_variance", "driver_dnf_rates")
cfg, mc, seed, _ = gc.get_case(case)
sim = mcgp.simulation.RaceSimulator(mcgp.simulation.RaceConfig(**cfg), pop_no_medium=POP[0], pop_no_soft=POP[1])
p = sim._params(*[mc.get(k) for k in MC_KEYS], mc.get("track_condition", "dry"), stream=0)
eng = mcgp.capi.get_engine(0)
n = 200
hist, trace = eng.run_native_traced([p], n, sim_begin=0, seed=seed, flags=mcgp.capi.F_EXACT_NORMAL, trace_first=0, trace_count=n)
ref = oracle.run_native(oracle.make_params(cfg, mc, *POP), seed, n, sim_begin=0, exact=True, trace=True)
tr, rr = trace[0], ref["trace"]
for f in ("position", "compound", "tire_age", "flags", "gap"):
    a, b = tr[f], rr[f]
    if f == "gap":
        a, b = a.view(np.uint32), b.view(np.uint32)
    d = np.argwhere(a != b)
    print(f, "mismatches:", len(d))
    if len(d):
        # earliest by (sim, lap)
        s, l, dr = d[0]
        print("  first: sim", s, "lap", l + 1, "driver", dr, "gpu", tr[f][s, l, dr], "cpu", rr[f][s, l, dr])
bad = np.argwhere((tr["gap"].view(np.uint32) != rr["gap"].view(np.uint32)) | (tr["position"] != rr["position"]) | (tr["flags"] != rr["flags"]) | (tr["tire_age"] != rr["tire_age"]))
if len(bad):
    s, l, _ = bad[0]
    for ll in (l - 1, l):
        if ll < 0: continue
        print(f"sim {s} lap {ll + 1}")
        for dr in range(tr.shape[2]):
            g, c = tr[s, ll, dr], rr[s, ll, dr]
            flag = " <<<" if (g["gap"].view(np.uint32) != c["gap"].view(np.uint32) or g["position"] != c["position"] or g["flags"] != c["flags"] or g["tire_age"] != c["tire_age"]) else ""
            print(f"  d{dr:2d} gpu pos {g['position']:2d} c{g['compound']} age {g['tire_age']:2d} fl {g['flags']:02x} gap {g['gap']:.6f} | cpu pos {c['position']:2d} c{c['compound']} age {c['tire_age']:2d} fl {c['flags']:02x} gap {c['gap']:.6f}{flag}")
