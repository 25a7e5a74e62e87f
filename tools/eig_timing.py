"""Device time of the eigen-solvers behind marginalize() on synthetic priors (CUDA-event free: wall clock around a
synchronising debug call, minus an empty call), plus the C3 / C5 marginalization wall clock (tools/marg_timing.py)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from helpers import pkg
lib = pkg.load()
f = lib.lib.ctvio_debug_eig
f.restype = C.c_int
f.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
dbg = (C.c_int * 8)()
rng = np.random.default_rng(5)
for n in (19, 66, 85, 112):
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    ev = 10.0 ** rng.uniform(-6, 6, n); ev[:6] = 10.0 ** rng.uniform(-14, -10, 6)
    a = np.ascontiguousarray((q * ev) @ q.T); a = 0.5 * (a + a.T)
    v, w = np.zeros((n, n)), np.zeros(n)
    for mode in ("blocked", "elementwise"):
        os.environ["CTVIO_JACOBI"] = mode
        best = 1e9
        for _ in range(5):
            t = time.perf_counter(); rc = f(n, a.ctypes.data, v.ctypes.data, w.ctypes.data, 0); assert rc == 0, rc; best = min(best, time.perf_counter() - t)
        (lib.lib.ctvio_debug_jacobi_blocked if mode == "blocked" else lib.lib.ctvio_debug_jacobi)(dbg)
        err = np.max(np.abs(np.sort(w) - np.linalg.eigvalsh(a))) / np.linalg.norm(a, 2)
        print(f"n={n:4d} {mode:12s} wall {1e3*best:7.3f} ms (incl. ~0.3 ms malloc/copies)  sweeps {dbg[0]}  -log10(off/diag) {dbg[2]}  ev err {err:.1e}" + (f"  rounds {dbg[3]} cycles/round: inner {dbg[4]} wait {dbg[5]} update {dbg[6]} wait {dbg[7]}" if mode == "blocked" else ""))
os.environ.pop("CTVIO_JACOBI")
