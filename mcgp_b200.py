"""Importable alias of the hyphenated package directory ``monte-carlo-gp_b200/``:
``import mcgp_b200`` == ``importlib.import_module("monte-carlo-gp_b200")``."""
import importlib
import sys

_pkg = importlib.import_module("monte-carlo-gp_b200")
for _name in ("workloads", "capi", "simulation", "distributed", "scoring", "grid_model", "ratings", "params_io", "season"):
    importlib.import_module(f"monte-carlo-gp_b200.{_name}")
sys.modules[__name__] = _pkg
