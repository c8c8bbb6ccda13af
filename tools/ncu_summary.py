#!/usr/bin/env python
"""Turn an `ncu --set full` capture into the small JSON bench.py reads for its roofline numbers.

    python tools/ncu_summary.py <file.ncu-rep> <tag> <units-in-the-profiled-launch> [--kernel native|replay] [--note "..."]

Writes profiles/<tag>_ncu_summary.json: the kernel's demangled name, executed warp-instructions per unit (race),
DRAM bytes read / written by the launch, issue / pipe utilisation, stall reasons per issue, divergence and occupancy
figures, and the SHA-256 of the kernel's source files at the time of the capture -- bench.py compares that hash with
the tree it runs from and prints `roofline.capture_matches_build`, so a number from an older build cannot pass as
the shipped kernel's.  Also dumps the raw metric page next to it (profiles/<tag>_ncu_raw.csv)."""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCES = {"native": ["monte-carlo-gp_b200/csrc/native_kernel.cu", "monte-carlo-gp_b200/csrc/native_math.cuh",
                      "monte-carlo-gp_b200/csrc/device_params.h"],
           "replay": ["monte-carlo-gp_b200/csrc/replay_kernel.cu", "monte-carlo-gp_b200/csrc/device_params.h"]}


def source_sha256(kind: str) -> str:
    h = hashlib.sha256()
    for rel in SOURCES[kind]:
        with open(os.path.join(ROOT, rel), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def main():
    rep, tag, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
    kind = sys.argv[sys.argv.index("--kernel") + 1] if "--kernel" in sys.argv else "native"
    note = sys.argv[sys.argv.index("--note") + 1] if "--note" in sys.argv else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, vals = rows[0], rows[-1]          # (row 1 holds the units; one profiled launch per capture)
    d = dict(zip(hdr, vals))

    def f(key, default=None):
        try:
            return float(d[key].replace(",", ""))
        except (KeyError, ValueError):
            return default

    inst = f("smsp__inst_executed.sum")
    unit_of = dict(zip(hdr, rows[1]))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}

    def bytes_of(key):
        v = f(key)
        return None if v is None else v * scale.get(unit_of.get(key, "byte"), 1.0)

    stalls = {k.split("stalled_")[1].split("_per_issue")[0]: round(float(v), 3) for k, v in d.items()
              if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and float(v) >= 0.05 and "selected_per" not in k[-40:]}
    out = {
        "tag": tag, "kernel": d.get("Kernel Name"), "kind": kind, "capture": os.path.basename(rep),
        "units_in_launch": units, "executed_warp_instr": inst, "executed_warp_instr_per_unit": inst / units,
        "duration_ms": f("gpu__time_duration.sum"), "sm_cycles": f("sm__cycles_elapsed.max"),
        "dram_bytes_read": bytes_of("dram__bytes_read.sum"), "dram_bytes_written": bytes_of("dram__bytes_write.sum"),
        "issue_active_pct": f("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
        "pipe_alu_pct": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "pipe_fma_pct": f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "pipe_fp64_pct": f("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "pipe_lsu_pct": f("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        "pipe_xu_pct": f("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        "lsu_data_pipe_wavefronts_pct": f("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        "shared_bank_conflict_wavefronts": f("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        "shared_wavefronts": f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        "threads_per_instruction": f("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "branch_targets_uniform_pct": f("smsp__sass_average_branch_targets_threads_uniform.pct"),
        "warps_active_pct_of_64": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": f("launch__registers_per_thread"),
        "stall_cycles_per_issue": stalls,
        "source_sha256": source_sha256(kind), "source_files": SOURCES[kind], "note": note,
    }
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.json"), "w") as fo:
        json.dump(out, fo, indent=1)
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_raw.csv"), "w") as fo:
        fo.write(raw)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
