"""The C-ABI library loads on a CPU-only box, exports every symbol include/mcgp.h declares, its structs have
the layout the ctypes binding assumes, and it fails loudly (no fallback) when there is no GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mcgp.h")


@pytest.fixture(scope="module")
def mcgp():
    import mcgp_b200
    return mcgp_b200


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mcgp_[a-z_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(mcgp):
    lib = mcgp.capi.load_library()
    names = _declared_functions()
    assert len(names) >= 11
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/mcgp.h but not exported by libmcgp.so"
    assert set(names) == set(mcgp.capi.EXPORTED_SYMBOLS), "ctypes binding and header disagree on the entry points"
    assert lib.mcgp_abi_version() == 1


def test_struct_layout_matches_header(mcgp, tmp_path):
    prog = tmp_path / "layout.c"
    fields = [f[0] for f in mcgp.capi.McgpRaceParams._fields_]
    lines = "\n".join(f'printf("{f} %zu\\n", offsetof(mcgp_race_params, {f}));' for f in fields)
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "mcgp.h"\nint main(void){\n'
                    'printf("sizeof %zu\\n", sizeof(mcgp_race_params));\n' + lines + "\nreturn 0;}\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    assert int(out["sizeof"]) == C.sizeof(mcgp.capi.McgpRaceParams)
    for f in fields:
        assert int(out[f]) == getattr(mcgp.capi.McgpRaceParams, f).offset, f


def test_header_is_plain_c(tmp_path):
    prog = tmp_path / "c89.c"
    prog.write_text('#include "mcgp.h"\nint main(void){return mcgp_abi_version == 0;}\n')
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(prog)])


def test_no_cpu_fallback(mcgp):
    """Without a GPU the product path must fail loudly, not fall back to a CPU implementation."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(mcgp.capi.McgpError) as e:
        mcgp.capi.Engine(0)
    assert e.value.code == mcgp.capi.ENODEVICE and "no CPU fallback" in str(e.value)
    cfg, mc = mcgp.workloads.workload("bahrain")
    sim = mcgp.simulation.RaceSimulator(mcgp.simulation.RaceConfig(**cfg))
    with pytest.raises(mcgp.capi.McgpError):
        sim.run_monte_carlo(10, mc["grid_probs"], mc["base_pace"], mc["tire_deg"], mc["driver_variance"])


def test_product_does_not_import_the_oracle():
    """Only tests/, smoke() and bench.py's CPU legs may touch oracle/ (task rule)."""
    pkg = os.path.join(ROOT, "monte-carlo-gp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "race_oracle" not in txt, f
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
