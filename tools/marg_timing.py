"""Wall-clock of ctvio_marginalize (K7) on the C3 window (prior dim ~190) and on a C5 window (prior dim ~85), GPU only."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from helpers import c3_window_a, pkg
st = importlib.import_module("ctrl-vio_b200.streaming")
lib = pkg.load()
g, seq, wa, nowk = c3_window_a(lib)
g.Solve(3)
for rep in range(3):
    t = time.perf_counter(); pr = g.SaveMarginalizationInfo(); dt = time.perf_counter() - t
    import ctypes as C
    dbg = (C.c_int * 8)(); lib.lib.ctvio_debug_jacobi(dbg)
    print(f"C3 marginalize: {1e3*dt:.3f} ms  (kept dim {pr.n}); last Jacobi: sweeps {dbg[0]} n {dbg[1]} -log10(off/diag) {dbg[2]}")
r = st.StreamingRunner(lib, st.config_c5_sequence(12))
r.run(10)
e = r.est
for rep in range(3):
    t = time.perf_counter(); pr = e.SaveMarginalizationInfo(); dt = time.perf_counter() - t
    dbg = (C.c_int * 8)(); lib.lib.ctvio_debug_jacobi(dbg)
    print(f"C5 marginalize: {1e3*dt:.3f} ms  (kept dim {pr.n}); last Jacobi: sweeps {dbg[0]} n {dbg[1]} -log10(off/diag) {dbg[2]}")
print("C5 per-window ms:", [round(x["ms"], 2) for x in r.records], "solve device ms", [round(x["device_ms"], 2) for x in r.records])
