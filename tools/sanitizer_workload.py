#!/usr/bin/env python
"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): every kernel variant once.

    compute-sanitizer --tool memcheck python tools/sanitizer_workload.py

Native kernel: count-table, finish/times and trace variants, a 3-race batch (blocks hop between races), a 32-car
field (the 8-key-vector variant) and the exact-normal build; replay kernel: tapes made on the host with uniform /
normal draws (bit-exactness is tests/; this run only exercises the memory and synchronisation behaviour)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcgp_b200 as m  # noqa: E402

eng = m.capi.get_engine(0)
wl = m.workloads


def params(name, stream=0, **kw):
    cfg, mc = wl.workload(name, **kw)
    sim = m.simulation.RaceSimulator(m.simulation.RaceConfig(**cfg), device=0, pop_no_medium="SOFT", pop_no_soft="MEDIUM")
    return sim._params(mc["grid_probs"], mc["base_pace"], mc["tire_deg"], mc["driver_variance"], mc["driver_dnf_rates"],
                       mc["track_condition"], stream=stream), cfg, mc


n = int(os.environ.get("MCGP_SANITIZE_SIMS", "3000"))
p, cfg, mc = params("bahrain")
h = eng.run_native([p], n, 0, 42)
assert int(h.sum()) == n * 20
h, fin, tim = eng.run_native([p], n, 5, 42, flags=m.capi.F_EXACT_NORMAL, want_finish=True, want_times=True)
assert fin.shape[1] == n
h, tr = eng.run_native_traced([p], n, 0, 42, trace_first=10, trace_count=100)
batch = [params(f"season:{r}", stream=r)[0] for r in (0, 5, 11)]
h = eng.run_native(batch, n, 0, 7)
assert h.shape[0] == 3 and int(h.sum()) == 3 * n * 20
# 32 cars: the NV4 = 8 variant
D = [f"D{i:02d}" for i in range(32)]
cfg32, mc32 = wl.workload("bahrain")
cfg32["driver_teams"] = {d: "Unknown" for d in D}
mc32 = dict(grid_probs=wl.gaussian_grid_probs(D), base_pace={d: 92.0 + 0.05 * k for k, d in enumerate(D)},
            tire_deg={d: 0.02 + 0.002 * k for k, d in enumerate(D)}, driver_variance={d: 0.15 for d in D},
            driver_dnf_rates={d: 0.001 for d in D}, track_condition="dry")
sim32 = m.simulation.RaceSimulator(m.simulation.RaceConfig(**cfg32), device=0, pop_no_medium="SOFT", pop_no_soft="MEDIUM")
h = sim32.run_monte_carlo_counts(n, **{k: mc32[k] for k in ("grid_probs", "base_pace", "tire_deg", "driver_variance",
                                                           "driver_dnf_rates")}, seed=3)
assert h.shape == (32, 32) and int(h.sum()) == n * 32
# replay: synthetic tapes of the worst-case length per sim
rng = np.random.default_rng(1)
ns = max(64, n // 20)
per = (4556, 1160, 20)
off = np.arange(ns + 1)[:, None] * np.array(per)[None, :]
sim = m.simulation.RaceSimulator(m.simulation.RaceConfig(**cfg), device=0, pop_no_medium="SOFT", pop_no_soft="MEDIUM")
out = sim.replay(mc["grid_probs"], mc["base_pace"], mc["tire_deg"], mc["driver_variance"], mc["driver_dnf_rates"],
                 u_py=rng.random(ns * per[0]), z=rng.standard_normal(ns * per[1]), u_np=rng.random(ns * per[2]),
                 offsets=off.astype(np.int64))
assert out["finish"].shape == (ns, 20)
# lap-histogram variant (one 32-warp block per SM), on-device scoring, the device-resident season loop
h, lh = eng.run_native_laphist([p], n, 0, 42)
assert int(lh[0, 0].sum()) <= n * 20
season = m.season.run_device_season(max(200, n // 10), 5, races=[0, 1, 2], pop_no_medium="SOFT", pop_no_soft="MEDIUM")
assert season["hist"].shape == (3, 20, 20)
print("sanitizer workload ok:", n, "native sims per variant,", ns, "replay sims")
