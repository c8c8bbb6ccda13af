"""B200-native Monte Carlo race engine: drop-in for the reference's src/simulation.py hot path.

The directory name carries a hyphen (it mirrors the upstream repository name), so import it with
``importlib.import_module("monte-carlo-gp_b200")`` or through the top-level alias ``mcgp_b200``.
Submodules are imported lazily; nothing here needs a GPU until a simulation is launched.
"""
__all__ = ["simulation", "workloads", "capi", "distributed", "scoring", "grid_model", "ratings", "params_io", "season"]
