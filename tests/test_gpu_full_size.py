"""BASELINE configs 3 / 4 / 5 at their FULL sizes, driver-run (one GPU; tests/full_size_configs.py is the script the
1- and 8-GPU results under profiles/ come from): 1e9 Monaco-78 sims, the 24 x 1e6-sim season in one launch, 10 x 1e8
sims over the five prediction points x {57, 19} laps, a 2e6-sim trace window.  Gates: every count table adds up; win /
podium / position cells agree with the reference (its bit-exact C port, 4e5 then 1.6e6 sims) within 3 sigma, two-stage;
and every table is bit-identical to the one committed under profiles/ (`sha256_count_table`: the native mode is a pure
function of (seed, sim index, inputs), on any number of GPUs -- the 8-GPU file holds the same hashes)."""
import json
import os
import types

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_baseline_configs_3_4_5_at_full_size():
    import full_size_configs
    res = full_size_configs.run(types.SimpleNamespace(scale=1.0, ref_sims=400_000, out=""))
    assert res["config3"]["sims"] == 1_000_000_000 and res["config4"]["sims_per_race"] == 1_000_000
    assert res["config5"]["total_sims"] == 1_000_000_000
    gates = [res["config3"]["vs_reference"]] + [e["vs_reference"] for e in res["config5"]["points"].values() if "vs_reference" in e]
    assert len(gates) == 5
    for g in gates:
        assert g["violations"] == [], g
    pinned = {}
    for name in ("r2_full_size_configs_1gpu.json", "r2_full_size_configs_8gpu.json"):
        with open(os.path.join(ROOT, "profiles", name)) as f:
            pinned[name] = json.loads(f.read())
    for name, ref in pinned.items():
        assert res["config3"]["sha256_count_table"] == ref["config3"]["sha256_count_table"], name
        assert res["config4"]["sha256_count_table"] == ref["config4"]["sha256_count_table"], name
        for point, e in res["config5"]["points"].items():
            assert e["sha256_count_table"] == ref["config5"]["points"][point]["sha256_count_table"], (name, point)
