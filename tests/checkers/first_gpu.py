import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import mcgp_b200 as m
import golden_cases as gc
from oracle import pyoracle as po
cfg, mc, seed, n = gc.get_case('bahrain_dry')
sim = m.simulation.RaceSimulator(m.simulation.RaceConfig(**cfg), pop_no_medium='SOFT', pop_no_soft='MEDIUM')
kw = {k: mc.get(k) for k in ("grid_probs", "base_pace", "tire_deg", "driver_variance", "driver_dnf_rates")}
for N in (10000, 1000000, 10000000):
    t = time.perf_counter(); h = sim.run_monte_carlo_counts(N, **kw, seed=42); dt = time.perf_counter() - t
    print(N, 'sims', dt, 's', N/dt, 'races/s; VER win', h[0,0]/N, 'podium', h[0,:3].sum()/N, 'sum', h.sum()/N)
ref = po.run_monte_carlo(cfg, mc, 400000, 1, threads=8)
N=10000000
print('ref VER win', ref[0,0]/400000, 'GPU', h[0,0]/N)
p_ref = ref/400000; p_gpu = h.astype(np.float64)/N
sig = np.sqrt(np.maximum(p_ref*(1-p_ref), 1e-9)*(1/400000+1/N))
zs = (p_gpu-p_ref)/sig
print('max |z|', np.abs(zs).max(), 'cells >3', (np.abs(zs)>3).sum(), 'of', zs.size)
np.set_printoptions(linewidth=250, precision=1, suppress=True)
print(zs[:8,:8])
