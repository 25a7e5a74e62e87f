"""debug: phase timing of CTA 0 inside chol_coop_kernel (needs a CTVIO_CHOL_TIMING build)."""
import ctypes as C, importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("ctrl-vio_b200"); syn = pkg.synthetic
LIB = pkg.CtvioLib(os.path.join(os.path.dirname(pkg.LIB_PATH), "libctvio_b200_timing.so"), "ctvio_")  # tools/build_timing.sh
for nt in (32, 256, 512):
    o = (C.c_longlong * 6)()
    LIB.lib.ctvio_debug_latency(nt, o)
    o = np.array(o[:], dtype=np.float64)
    print(f"latency probe, {nt} threads/CTA: dep DFMA {o[0]/256:.1f} cyc, dep DMUL {o[1]/256:.1f}, dep rsqrt(double) {o[2]/64:.1f}, "
          f"STS+bar+LDS+DADD+bar {o[3]/64:.1f}, 16 indep DFMA chains {o[4]/1024:.2f} cyc/DFMA, dep FFMA {o[5]/256:.1f}")
for name in ("c2", "c4"):
    w = syn.config_c2() if name == "c2" else syn.config_c4()
    est = pkg.setup_estimator(LIB, w)
    est.Solve(2)
    buf = (C.c_ulonglong * 512)()
    LIB.lib.ctvio_debug_chol_stamps(buf, 512)
    t = np.array(buf[:], dtype=np.float64)
    nb = (6 * w.n_knots + 6 * len(w.kf_times) + 1 + 63) // 64
    print(name, "nb", nb)
    i = 0
    print(" init+sync %.1f us" % ((t[1] - t[0]) / 1e3)); i = 1
    for k in range(nb):
        load = t[i + 1] - t[i]
        fact = t[i + 2] - t[i + 1]; slab = t[i + 3] - t[i + 2]
        if k == nb - 1:
            print(f" step {k}: load {load/1e3:.1f} factor+inv {fact/1e3:.1f} slab/xk {slab/1e3:.1f}"); i += 3; break
        s1 = t[i + 4] - t[i + 3]; upd = t[i + 5] - t[i + 4]; s2 = t[i + 6] - t[i + 5]
        print(f" step {k}: load {load/1e3:.1f} factor+inv {fact/1e3:.1f} slab {slab/1e3:.1f} sync1 {s1/1e3:.1f} update {upd/1e3:.1f} sync2 {s2/1e3:.1f}")
        i += 6
    print(" backward %.1f us, total %.1f us" % ((t[i + 1] - t[i]) / 1e3, (t[i + 1] - t[0]) / 1e3))
