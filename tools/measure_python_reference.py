#!/usr/bin/env python
"""CPU baseline leg of SURVEY 8(d) / BASELINE.md section 3: the UNMODIFIED reference `RaceSimulator.run_monte_carlo`
(/root/reference/src/simulation.py:59-100) on every host core of THIS (build) box -- a multiprocessing pool of
os.cpu_count() workers, worker k running run_monte_carlo(n_per_worker, ..., seed=42+k) on the bench workload
(Bahrain-57, 20 drivers); wall clock of the slowest worker.  The reference tree does not exist on the GPU box, so
bench.py cannot run this there: it quotes the committed result (profiles/r2_python_reference.json), labelled as
measured on the build box.     usage: python tools/measure_python_reference.py [sims_per_worker]"""
import json
import multiprocessing as mp
import os
import platform
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def worker(args):
    k, n = args
    sys.dont_write_bytecode = True
    sys.path.insert(0, "/root/reference")
    from src.simulation import RaceSimulator, RaceConfig      # the unmodified reference
    import mcgp_b200
    cfg, mc = mcgp_b200.workloads.workload("bahrain")
    sim = RaceSimulator(RaceConfig(**cfg))
    t = time.perf_counter()
    res = sim.run_monte_carlo(n, mc["grid_probs"], mc["base_pace"], mc["tire_deg"], mc["driver_variance"],
                              mc["driver_dnf_rates"], seed=42 + k, track_condition=mc["track_condition"])
    return time.perf_counter() - t, float(res["VER"].get(1, 0))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    with mp.Pool(cores) as pool:
        out = pool.map(worker, [(k, n) for k in range(cores)])
    wall = time.perf_counter() - t0
    slowest = max(t for t, _ in out)
    res = {"what": "unmodified /root/reference RaceSimulator.run_monte_carlo, multiprocessing pool, one worker per core",
           "where": "build box (no GPU): " + platform.processor() + " / " + platform.machine(), "cores": cores,
           "sims_per_worker": n, "seed": "42 + worker", "slowest_worker_s": slowest, "pool_wall_s": wall,
           "races_per_s_total": cores * n / slowest, "races_per_s_per_core": n / slowest,
           "p_ver_wins_mean": sum(w for _, w in out) / cores, "python": sys.version.split()[0]}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "r2_python_reference.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
