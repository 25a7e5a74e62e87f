"""Small target for ncu: one window resident in HBM, a few LM iterations (never a bench value)."""
import argparse
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("ctrl-vio_b200")
syn = pkg.synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2", choices=["c2", "c4", "c2-ldfree"])
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--solves", type=int, default=2)
a = ap.parse_args()
w = {"c2": syn.config_c2, "c4": syn.config_c4, "c2-ldfree": lambda: syn.config_c2(fix_ld=False)}[a.workload]()
est = pkg.setup_estimator(pkg.load(), w)
est.SaveState()
for _ in range(a.solves):
    est.RestoreState()
    s = est.Solve(a.iters)
print(a.workload, s.as_dict())
