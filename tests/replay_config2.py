#!/usr/bin/env python
"""BASELINE config 2 at size: the Bahrain-57 race, N sims in replay mode, bit-exact against the reference's draws.

Chunk c is the reference's run_monte_carlo(n_chunk, ..., seed=42+c): the CPU oracle (pinned bit-exact to
the unmodified reference on the golden fixtures) produces that chunk's MT19937 / legacy-Gaussian draw tapes and its
finishing orders / race times; the GPU replays the tapes through the C ABI and every sim is compared.

    python tests/replay_config2.py [--chunks 1000] [--chunk-sims 10000] [--threads 16] [--out profiles/x.json]
This is a checker script (it runs the oracle), not a product path.
"""
import argparse, json, os, sys, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def run(chunks: int, chunk_sims: int, threads: int, seed0: int = 42, workload: str = "bahrain") -> dict:
    import mcgp_b200 as mcgp
    from oracle import pyoracle as oracle
    cfg, mc = mcgp.workloads.workload(workload)
    pop = ("SOFT", "MEDIUM")
    oparams = oracle.make_params(cfg, mc, *pop)
    sim = mcgp.simulation.RaceSimulator(mcgp.simulation.RaceConfig(**cfg), pop_no_medium=pop[0], pop_no_soft=pop[1])
    kw = {k: mc.get(k) for k in ("grid_probs", "base_pace", "tire_deg", "driver_variance", "driver_dnf_rates")}
    n = len(mc["grid_probs"])
    total_hist = np.zeros((n, n), np.int64)
    res = dict(sims=0, order_mismatches=0, time_mismatches=0, max_rel_time_err=0.0, draws_mismatches=0, tape_bytes=0,
               t_oracle=0.0, t_gpu=0.0)

    def make(c):
        t0 = time.perf_counter()
        ref = oracle.run_streams(oparams, oracle.Rng(seed0 + c), chunk_sims, detail=True, tapes=True)
        return ref, time.perf_counter() - t0

    with ThreadPoolExecutor(threads) as ex:
        pending = [ex.submit(make, c) for c in range(min(threads * 2, chunks))]
        nxt = len(pending)
        for c in range(chunks):
            ref, dt = pending[c].result()
            pending[c] = None
            if nxt < chunks:
                pending.append(ex.submit(make, nxt)); nxt += 1
            res["t_oracle"] += dt
            t0 = time.perf_counter()
            got = sim.replay(**kw, track_condition=mc.get("track_condition", "dry"), u_py=ref["tape_upy"], z=ref["tape_z"],
                             u_np=ref["tape_unp"], offsets=ref["tape_off"])
            res["t_gpu"] += time.perf_counter() - t0
            res["sims"] += chunk_sims
            res["tape_bytes"] += 8 * int(ref["draws"][-1].sum())
            res["order_mismatches"] += int((got["finish"] != ref["finish"]).any(1).sum()) + int((got["grid"] != ref["grid"]).any(1).sum())
            tb = got["times"].view(np.uint64) != ref["times"].view(np.uint64)
            res["time_mismatches"] += int(tb.any(1).sum())
            if tb.any():
                denom = np.maximum(np.abs(ref["times"]), 1e-300)
                res["max_rel_time_err"] = max(res["max_rel_time_err"], float((np.abs(got["times"] - ref["times"]) / denom).max()))
            res["draws_mismatches"] += int((got["used"].cumsum(0) != ref["draws"]).any(1).sum())
            assert np.array_equal(got["hist"].astype(np.int64), ref["hist"])
            total_hist += ref["hist"]
    res["gpu_replay_races_per_s_e2e"] = res["sims"] / res["t_gpu"]
    res["oracle_races_per_s_per_thread"] = res["sims"] / res["t_oracle"]
    res["workload"] = f"{workload}: {chunks} chunks x {chunk_sims} sims, chunk c = run_monte_carlo(seed={seed0}+c)"
    res["win_counts"] = total_hist[:, 0].tolist()
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=1000)
    ap.add_argument("--chunk-sims", type=int, default=10000)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    r = run(a.chunks, a.chunk_sims, a.threads)
    print(json.dumps(r))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(r, f, indent=1)
    sys.exit(0 if r["order_mismatches"] == 0 and r["time_mismatches"] == 0 and r["draws_mismatches"] == 0 else 1)
