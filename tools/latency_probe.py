import time, sys
sys.path.insert(0, '/root/repo')
import mcgp_b200
cfg, mc = mcgp_b200.workloads.workload("bahrain")
sim = mcgp_b200.simulation.RaceSimulator(mcgp_b200.simulation.RaceConfig(**cfg), pop_no_medium="SOFT", pop_no_soft="MEDIUM")
for n in (10000, 100000, 1000000):
    sim.run_monte_carlo(n, seed=1, **mc)
    t0 = time.perf_counter()
    reps = 20
    for i in range(reps):
        r = sim.run_monte_carlo(n, seed=i, **mc)
    dt = (time.perf_counter() - t0) / reps
    print(f"run_monte_carlo({n}): {dt*1e3:.3f} ms per call -> {n/dt/1e6:.2f} M races/s")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(50): sim.run_monte_carlo(10000, seed=i, **mc)
pr.disable(); pstats.Stats(pr).sort_stats('cumulative').print_stats(14)
