#!/usr/bin/env python
"""Headline benchmark: simulated races per second of the native (Philox / FP32) race kernel.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--sims-per-step S] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json metric "race-sims/sec (20 drv x 57 laps)"): the synthetic Bahrain-like race of
SURVEY.md §8(d) -- 20 drivers, 57 laps, product event probabilities -- S sims per GPU per step (weak scaling).
A step = one pass of the hot path over one batch: ONE kernel launch simulating S races per GPU and, at N > 1,
the ONE all-reduce of the 20x20 int64 count table.

  value   races/s over all GPUs, device-timed (CUDA events on the launch stream, per step, max over ranks), with
          the race parameters already resident in HBM.
  e2e     the same metric through the reference-facing API (RaceSimulator.run_monte_carlo_counts: host dicts ->
          mcgp_race_params -> H2D -> kernel -> D2H count table), wall clock, copies inside the timed region.
  roofline  ALU-issue bound (there is no dense contraction and ~no HBM traffic): algorithmic
          warp-instructions (W_race(57) = 17 100 per race) / s against N_SM x 4 schedulers x f_SM.
  cpu_baseline  the CPU restatement of the reference (oracle/, C, bit-exact to the Python reference) on all host
          cores, bounded sample.  `--impl reference` times that CPU implementation as its own arm.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W_RACE = {57: 17100, 78: 23190}  # algorithmic warp-instructions per race, SURVEY.md §8(d)
# The ncu figures of the headline kernel come from the committed summary of its capture (tools/ncu_summary.py writes
# it together with the SHA-256 of the kernel sources); nothing from a profiler is hard-coded here, and the line says
# whether the capture belongs to the build that ran (roofline.capture_matches_build).
NCU_SUMMARY = os.path.join("profiles", "r2_native_ncu_summary.json")
PYTHON_REFERENCE = os.path.join("profiles", "r2_python_reference.json")
KERNEL_SOURCES = ("monte-carlo-gp_b200/csrc/native_kernel.cu", "monte-carlo-gp_b200/csrc/native_math.cuh",
                  "monte-carlo-gp_b200/csrc/device_params.h")
N_DRIVERS, LAPS = 20, 57
WORKLOAD = "bahrain57: 20 drivers x 57 laps, native Philox4x32-7/FP32, synthetic inputs of SURVEY 8(d)"
HASH_SIMS, HASH_SEED = 8_000_000, 42   # the fixed global sim range whose count table is hashed (any N must agree)


def load_json(rel):
    try:
        with open(os.path.join(ROOT, rel)) as f:
            return json.load(f)
    except (OSError, ValueError):
        return None


def kernel_source_sha256() -> str:
    import hashlib
    h = hashlib.sha256()
    for rel in KERNEL_SOURCES:
        with open(os.path.join(ROOT, rel), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def w_race(laps: int) -> int:
    return W_RACE.get(laps, 600 + 110 + 290 * (laps - 1) + 150)


def inputs():
    import mcgp_b200
    cfg, mc = mcgp_b200.workloads.workload("bahrain")
    return mcgp_b200, cfg, mc


# ---- clocks / throttle reasons sampled during the timed region ---------------------------------
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.samples, self.proc, self.thread, self.idx = [], None, None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], 0, set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ts, line in self.samples:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7 or not (t0 <= ts <= t1 + 0.2):
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "samples": len(sm),
                "power_w_max": max(power) if power else None, "reasons": sorted(reasons)}


def physical_gpu_index(local: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except (ValueError, IndexError):
            return local
    return local


# ---- CPU implementation of the path (the oracle port of the reference) ---------------------------
def cpu_run(n_sims: int, seed: int, threads: int) -> tuple[float, int]:
    """Times the CPU restatement of the reference race loop on `threads` host threads; returns (seconds, sims)."""
    from oracle import pyoracle as po          # allowed here: cpu_baseline / --impl reference legs only
    import mcgp_b200
    cfg, mc = mcgp_b200.workloads.workload("bahrain")
    po.lib()
    t0 = time.perf_counter()
    hist = po.run_monte_carlo(cfg, mc, n_sims, seed, threads=threads)
    dt = time.perf_counter() - t0
    assert int(hist.sum()) == n_sims * N_DRIVERS
    return dt, n_sims


def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the Python original cannot
    travel to the GPU box and has no compiled form), all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = args.cpu_sims or 40000 * cores
    for _ in range(args.warmup):
        cpu_run(max(per_step // 10, cores), 1, cores)
    t = 0.0
    for k in range(args.steps):
        dt, _ = cpu_run(per_step, 42 + k, cores)
        t += dt
    value = per_step * args.steps / t
    sample = f"{per_step} sims/step x {args.steps} steps, {cores} threads, seed 42+step (C port of src/simulation.py, bit-exact to it)"
    print(json.dumps({
        "impl": "reference", "metric": "race-sims/sec (20 drv x 57 laps)", "value": value, "unit": "races/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD.replace("native Philox4x32-7/FP32", "reference MT19937/FP64 on CPU")},
        "driver_laps_per_s": value * N_DRIVERS * LAPS,
        "cpu_baseline": {"value": value, "unit": "races/s", "cores": cores, "kind": "port", "sample": sample,
                         "python_reference": load_json(PYTHON_REFERENCE)},
        "e2e": {"value": value, "unit": "races/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--sims-per-step", type=int, default=10_000_000, help="sims per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sims", type=int, default=0, help="CPU baseline sample size (0 = auto, ~10-20 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    mcgp, cfg, mc = inputs()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the race engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_gpus = world
    S = args.sims_per_step
    seed = 42

    sim = mcgp.simulation.RaceSimulator(mcgp.simulation.RaceConfig(**cfg), device=local, pop_no_medium="SOFT", pop_no_soft="MEDIUM")
    mc_args = (mc["grid_probs"], mc["base_pace"], mc["tire_deg"], mc["driver_variance"], mc["driver_dnf_rates"])
    params = sim._params(*mc_args, mc["track_condition"])
    sharded = mcgp.distributed.ShardedSimulator([params], device=local)   # parameters resident in HBM from here on
    dev = torch.device("cuda", local)
    hist = torch.zeros((1, N_DRIVERS, N_DRIVERS), dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)        # > 126 MB L2
    info = sharded.engine.device_info()

    def step(k: int):
        """one pass: this rank's S sims of global step k (global sim ids never repeat), then the one all-reduce"""
        begin = (k * n_gpus + rank) * S
        hist.zero_()
        sharded.launch(begin, S, seed, hist)
        if world > 1:
            dist.all_reduce(hist)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(args.warmup):
        step(k)
    barrier()
    assert int(hist.sum().item()) == S * n_gpus * N_DRIVERS, "count table does not add up"

    sampler = ClockSampler(physical_gpu_index(local)) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    launches = 0
    for k in range(args.steps):
        flush.fill_(k & 0xff)                       # L2 flush between timed iterations, outside the step's events
        ev[k][0].record()
        step(args.warmup + k)
        ev[k][1].record()
        launches += sharded.engine.last_launch_count     # the library's own count for this step (counter reset + race kernel)
    barrier()
    t_wall1 = time.perf_counter()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    tot = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    dev_ms = float(tot.item())
    value = S * n_gpus * args.steps / (dev_ms * 1e-3)

    # ---- e2e: the public drop-in API with host buffers, copies inside the timed region --------------
    # (every step passes a different `stream` id: the library skips derivation + upload for a batch identical to the
    # resident one, and an end-to-end step has to pay for its inputs)
    e2e_steps = max(2, min(args.steps, 5))
    sim.run_monte_carlo_counts(S, *mc_args, seed=seed, sim_begin=rank * S)
    barrier()
    h2d_params = 0
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        h = sim.run_monte_carlo_counts(S, *mc_args, seed=seed, sim_begin=((100 + k) * n_gpus + rank) * S, stream=1 + k)
        h2d_params = mcgp.capi.get_engine(local).last_upload_bytes()
        if world > 1:
            ht = torch.from_numpy(h.astype("int64")).to(dev)
            dist.all_reduce(ht)
            h = ht.cpu().numpy()
    barrier()
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = S * n_gpus * e2e_steps / float(e2e_t.item())
    # bytes per step: the += count table both ways, plus what the library uploads on every host-buffer call (the
    # derived parameter blocks and the overtake pace table, counted by the library itself)
    h2d = N_DRIVERS * N_DRIVERS * 8 + h2d_params
    d2h = N_DRIVERS * N_DRIVERS * 8

    # ---- bit-identity across GPU counts: SHA-256 of the count table of a FIXED global sim range -----
    # sims [0, HASH_SIMS) with seed HASH_SEED, sharded over the N ranks, one all-reduce: BENCH / SCALE lines of any N
    # must carry the same hash (draws are keyed by the global sim index).
    import hashlib
    hh = sharded.run(HASH_SIMS, HASH_SEED)
    torch.cuda.synchronize()
    hist_sha256 = hashlib.sha256(hh.cpu().numpy().astype("<i8").tobytes()).hexdigest()

    # ---- the product-sized call: run_monte_carlo(10 000) as src/predictor.py:283-291 makes it ----------
    product = None
    if rank == 0:
        import statistics
        lat = {}
        for label, vary in (("fresh_params", True), ("resident_params", False)):
            ts = []
            for i in range(60):
                t1 = time.perf_counter()
                sim.run_monte_carlo_counts(10_000, *mc_args, seed=1000 + i, stream=(500 + i) if vary else 0)
                ts.append(time.perf_counter() - t1)
            lat[label] = 1e3 * statistics.median(ts[10:])
        ts = []
        for i in range(60):
            t1 = time.perf_counter()
            sim.run_monte_carlo(10_000, *mc_args, seed=2000 + i)
            ts.append(time.perf_counter() - t1)
        lat["run_monte_carlo_dicts_in_dict_out"] = 1e3 * statistics.median(ts[10:])
        product = {"ms": lat["run_monte_carlo_dicts_in_dict_out"], "sims": 10_000, "calls": 50, "statistic": "median",
                   "counts_api_fresh_params_ms": lat["fresh_params"], "counts_api_resident_params_ms": lat["resident_params"],
                   "what": "RaceSimulator.run_monte_carlo(10 000, host dicts) -> {driver: {pos: p}} as src/predictor.py:283-291 calls it; "
                           "fresh = parameters differ from the previous call (derived + uploaded), resident = same race again"}

    # ---- replay mode (BASELINE config 2), reported beside the headline ----------------------------
    # Throughput only (bit-exactness is tests/ and tests/replay_config2.py): synthetic uniform / normal tapes made
    # on the device, cut into per-sim slices of the worst-case draw count (4556 U_py / 1160 Z / 20 U_np).
    replay = None
    if rank == 0:
        try:
            import numpy as np
            n_rep = 100_000   # enough races to fill the GPU (21 per resident warp); ~3.7 GB of synthetic tapes
            n_py = N_DRIVERS + (LAPS - 1) * (4 + N_DRIVERS + 3 * (N_DRIVERS - 1))
            n_z = 2 * N_DRIVERS + (LAPS - 1) * N_DRIVERS
            g = torch.Generator(device=dev).manual_seed(42)
            tapes = [torch.rand(n_rep * n_py, dtype=torch.float64, device=dev, generator=g),
                     torch.randn(n_rep * n_z, dtype=torch.float64, device=dev, generator=g),
                     torch.rand(n_rep * N_DRIVERS, dtype=torch.float64, device=dev, generator=g)]
            off = torch.from_numpy(np.arange(n_rep + 1, dtype=np.int64)[:, None] * np.array([n_py, n_z, N_DRIVERS], np.int64)).contiguous().to(dev)
            rh = torch.zeros((N_DRIVERS, N_DRIVERS), dtype=torch.int64, device=dev)
            used = torch.zeros((n_rep, 3), dtype=torch.int64, device=dev)
            status = torch.zeros(4, dtype=torch.int32, device=dev)
            eng = sharded.engine
            st = torch.cuda.current_stream().cuda_stream
            reps = 3

            def go():
                eng.launch_replay(n_rep, tapes[0].data_ptr(), tapes[1].data_ptr(), tapes[2].data_ptr(), off.data_ptr(),
                                  rh.data_ptr(), used_ptr=used.data_ptr(), status_ptr=status.data_ptr(), stream=st)
            for _ in range(2):
                go()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                go()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / reps
            consumed = int(used.sum().item()) * 8
            replay = {"races_per_s": n_rep / (ms * 1e-3), "sims": n_rep, "dtype": "f64", "status": int(status[0].item()),
                      "tape_bytes_consumed_per_race": consumed / n_rep, "tape_gb_per_s": consumed / (ms * 1e-3) / 1e9,
                      "count_table_ok": int(rh.sum().item()) == (reps + 2) * n_rep * N_DRIVERS}
        except Exception as e:  # the replay figure is auxiliary; never lose the headline line over it
            replay = {"error": repr(e)}

    # ---- BASELINE configs 4 and 5, reported beside the headline (rank 0, same device-timed method) -----
    aux = {}
    if rank == 0:
        try:
            eng = sharded.engine
            st = torch.cuda.current_stream().cuda_stream

            def timed(fn, reps=3):
                fn(); torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(reps):
                    fn()
                b.record(); torch.cuda.synchronize()
                return a.elapsed_time(b) / reps
            # config 4: the 24-race season in ONE launch (one block column per race, 24 count tables)
            wl = mcgp.workloads
            plist = []
            for r in range(wl.N_SEASON_RACES):
                c4, m4 = wl.workload(f"season:{r}")
                s4 = mcgp.simulation.RaceSimulator(mcgp.simulation.RaceConfig(**c4), device=local, pop_no_medium="SOFT", pop_no_soft="MEDIUM")
                plist.append(s4._params(m4["grid_probs"], m4["base_pace"], m4["tire_deg"], m4["driver_variance"],
                                        m4["driver_dnf_rates"], m4["track_condition"], stream=r))
            eng.upload_races(plist)
            n4 = max(1, S // 24)
            h4 = torch.zeros((len(plist), N_DRIVERS, N_DRIVERS), dtype=torch.int64, device=dev)
            ms = timed(lambda: eng.launch_native(n4, 0, seed, h4.data_ptr(), stream=st))
            laps4 = sum(p.total_laps for p in plist)
            aux["season_batch"] = {"races_in_batch": len(plist), "sims_per_race": n4, "races_per_s": len(plist) * n4 / (ms * 1e-3),
                                   "driver_laps_per_s": n4 * laps4 * N_DRIVERS / (ms * 1e-3), "ms": ms}
            # config 5: per-lap trace of every sim (8 B per driver-lap) -- the variant that writes to HBM
            eng.upload_races([params])
            n5 = min(S, 4_000_000)
            tr = torch.empty(n5 * LAPS * N_DRIVERS * 8, dtype=torch.uint8, device=dev)
            h5 = torch.zeros((1, N_DRIVERS, N_DRIVERS), dtype=torch.int64, device=dev)
            ms = timed(lambda: eng.launch_native_traced(n5, 0, seed, h5.data_ptr(), tr.data_ptr(), 0, n5, stream=st))
            aux["trace_mode"] = {"sims": n5, "races_per_s": n5 / (ms * 1e-3), "trace_bytes_per_race": LAPS * N_DRIVERS * 8,
                                 "hbm_write_gb_per_s": n5 * LAPS * N_DRIVERS * 8 / (ms * 1e-3) / 1e9, "ms": ms}
            del tr
            # config 5's alternative output: the trace reduced on-chip to per-lap position histograms (57 x 20 x 20 counters)
            lh = torch.zeros((1, LAPS, N_DRIVERS, N_DRIVERS), dtype=torch.int64, device=dev)
            h6 = torch.zeros((1, N_DRIVERS, N_DRIVERS), dtype=torch.int64, device=dev)
            ms = timed(lambda: eng.launch_native_laphist(S, 0, seed, h6.data_ptr(), lh.data_ptr(), stream=st))
            aux["lap_histogram_mode"] = {"sims": S, "races_per_s": S / (ms * 1e-3), "counters": LAPS * N_DRIVERS * N_DRIVERS, "ms": ms}
            eng.upload_races([params])
        except Exception as e:  # auxiliary figures; never lose the headline line over them
            aux["error"] = repr(e)

    # ---- CPU baseline on this box's host cores (rank 0, N == 1 only) --------------------------------
    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_cpu = args.cpu_sims or 120000 * cores   # ~12 s of CPU work
        dt, n_done = cpu_run(n_cpu, 42, cores)
        pyref = load_json(PYTHON_REFERENCE)   # the unmodified Python reference, measured where /root/reference exists
        cpu = {"value": n_done / dt, "unit": "races/s", "cores": cores, "kind": "port",
               "sample": f"{n_done} sims of the same race, {cores} threads, {dt:.1f} s; C restatement of src/simulation.py "
                         f"(bit-exact to the Python reference)",
               "python_reference": None if pyref is None else {
                   "races_per_s_per_core": pyref["races_per_s_per_core"], "races_per_s_total": pyref["races_per_s_total"],
                   "cores": pyref["cores"], "measured_on": "build box (not this GPU box): " + pyref["where"],
                   "how": pyref["what"] + f", {pyref['sims_per_worker']} sims per worker (tools/measure_python_reference.py)"}}

    if world > 1:
        dist.barrier()
    if rank == 0:
        sm_count = info["sm_count"]
        f_max = (clocks or {}).get("sm_max_mhz") or info["sm_clock_khz"] / 1e3
        f_run = (clocks or {}).get("sm_mhz") or f_max
        per_gpu = value / n_gpus
        achieved = per_gpu * w_race(LAPS) / 1e12                 # Twarp-instr/s per GPU (algorithmic)
        peak = sm_count * 4 * f_max * 1e6 / 1e12
        peak_run = sm_count * 4 * f_run * 1e6 / 1e12
        ncu = load_json(NCU_SUMMARY) or {}
        executed = ncu.get("executed_warp_instr_per_unit")
        traffic = None if ncu.get("dram_bytes_read") is None else ncu["dram_bytes_read"] + (ncu.get("dram_bytes_written") or 0)
        build_sha = kernel_source_sha256()
        line = {
            "metric": "race-sims/sec (20 drv x 57 laps)", "value": value, "unit": "races/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sims_per_gpu_per_step": S, "n_drivers": N_DRIVERS, "laps": LAPS, "seed": seed,
                       "parallelism": f"sim-sharded x{n_gpus}, one int64 all-reduce of the 20x20 count table per step",
                       "l2": "256 MiB fill between timed steps (outside the per-step events); the kernel's inputs are a 6.9 KB parameter block + a 19.8 KB pace table"},
            "driver_laps_per_s": value * N_DRIVERS * LAPS,
            "e2e": {"value": e2e_value, "unit": "races/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "RaceSimulator.run_monte_carlo_counts (host dicts in, count table out)", "steps": e2e_steps},
            "gpu_launches": launches * n_gpus,
            "clocks": clocks,
            "roofline": {"bound": "alu_issue", "achieved": achieved, "peak": peak, "unit": "Twarp-instr/s per GPU",
                         "frac": achieved / peak, "frac_at_sampled_clock": achieved / peak_run,
                         "algorithmic_warp_instr_per_race": w_race(LAPS), "sm_count": sm_count, "f_sm_mhz_max": f_max,
                         "f_sm_mhz_sampled": f_run, "traffic": traffic,
                         "executed_warp_instr_per_race": executed,
                         "issue_slot_utilisation": None if executed is None else per_gpu * executed / 1e12 / peak,
                         "ncu_issue_active_pct": ncu.get("issue_active_pct"),
                         "capture": {"summary": NCU_SUMMARY if ncu else None, "kernel": ncu.get("kernel"),
                                     "source_sha256": ncu.get("source_sha256"), "build_source_sha256": build_sha},
                         "capture_matches_build": bool(ncu) and ncu.get("source_sha256") == build_sha,
                         "note": "no dense contraction and ~0 HBM traffic (SURVEY 8(d)): the bound is warp-instruction issue, "
                                 "N_SM x 4 x f_SM.  frac = ALGORITHMIC work (17 100 warp-instr per race, SURVEY 8(d)) / peak; "
                                 "issue_slot_utilisation = instructions this kernel actually executes per race (ncu capture "
                                 "named in `capture`, same sources iff capture_matches_build) x races/s / peak; traffic = dram "
                                 "bytes read + written by the captured launch (2 M races, no L2 fill before it)"},
            "cpu_baseline": cpu,
            "hist_sha256": {"value": hist_sha256, "sims": HASH_SIMS, "seed": HASH_SEED,
                            "what": "SHA-256 of the int64 count table of global sims [0, sims), sharded over the N ranks + all-reduce"},
            "e2e_product_call_ms": product["ms"] if product else None, "product_call": product,
            "replay_mode": replay,
            "season_batch": aux.get("season_batch"), "trace_mode": aux.get("trace_mode"),
            "lap_histogram_mode": aux.get("lap_histogram_mode"),
        }
        if "error" in aux:
            line["aux_error"] = aux["error"]
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
