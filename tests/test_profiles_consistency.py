"""The ncu evidence under profiles/ belongs to the sources in the tree: every `*_ncu_summary.json` of the shipped kernels
carries the SHA-256 of the kernel sources it was captured from (tools/ncu_summary.py), and bench.py prints
`roofline.capture_matches_build` from the same comparison.  A kernel edit without a new capture fails here first."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.parametrize("tag,kind", [("r2_native", "native"), ("r2_native_trace", "native"), ("r2_native_laphist", "native"),
                                       ("r2_replay", "replay")])
def test_capture_matches_sources(tag, kind):
    import ncu_summary
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.json")) as f:
        d = json.load(f)
    assert d["kind"] == kind
    assert d["source_sha256"] == ncu_summary.source_sha256(kind), f"profiles/{tag}_ncu_summary.json was captured from other sources"
    assert d["executed_warp_instr_per_unit"] > 0 and 0 < d["issue_active_pct"] <= 100


def test_bench_reads_the_headline_capture():
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert "r2_native_ncu_summary.json" in src
    line = json.loads(open(os.path.join(ROOT, "profiles", "r2_bench_line.json")).read().strip().splitlines()[-1])
    assert line["roofline"]["capture_matches_build"] is True
    with open(os.path.join(ROOT, "profiles", "r2_native_ncu_summary.json")) as f:
        cap = json.load(f)
    assert abs(line["roofline"]["executed_warp_instr_per_race"] - cap["executed_warp_instr_per_unit"]) < 1e-6
