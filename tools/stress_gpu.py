"""debug: soak test - repeated dense-solver self-checks and solve parity on several seeds (GPU vs oracle)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
pkg = importlib.import_module("ctrl-vio_b200"); syn = pkg.synthetic
from helpers import rot_angle_between
lib = pkg.load()
olib = pkg.CtvioLib(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "liboracle.so"), "ctvo_",
                    optional=pkg.binding.DEVICE_ONLY_SYMBOLS)
bad = 0
for name, mk, reps in (("c2", syn.config_c2, 2000), ("c4", syn.config_c4, 300)):
    est = pkg.setup_estimator(lib, mk())
    mm, res = est.SelfcheckSolver(reps=reps)
    print(f"{name}: {reps} repeated solves, bitwise mismatches {mm}, rel residual {res:.2e}", flush=True)
    bad += mm != 0 or res > 1e-9
for seed in range(8):
    w = syn.config_c2(seed=syn.SEED0 + 100 + seed, fix_ld=bool(seed & 1))
    g = pkg.setup_estimator(lib, w); o = pkg.setup_estimator(olib, w)
    sg, so = g.Solve(15), o.Solve(15)
    (qg, pg), (qo, po) = g.GetKnots(), o.GetKnots()
    rel = np.abs(pg - po).max() / np.abs(po).max(); ang = rot_angle_between(qg, qo).max()
    ok = sg.iterations == so.iterations and rel < 1e-5 and ang < 1e-4 and abs(sg.final_cost / so.final_cost - 1) < 1e-6
    print(f"seed {seed} fix_ld={bool(seed & 1)}: it {sg.iterations}/{so.iterations} cost rel {abs(sg.final_cost/so.final_cost-1):.1e} "
          f"rel_t {rel:.1e} ang {ang:.1e} {'OK' if ok else 'MISMATCH'}", flush=True)
    bad += not ok
print("FAILURES:", bad)
sys.exit(1 if bad else 0)
