"""debug: per-column-CTA timeline of chol_dag_kernel (needs tools/build_timing.sh)."""
import ctypes as C, importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("ctrl-vio_b200"); syn = pkg.synthetic
LIB = pkg.CtvioLib(os.path.join(os.path.dirname(pkg.LIB_PATH), "libctvio_b200_timing.so"), "ctvio_")
for name in ("c2", "c4"):
    w = syn.config_c2() if name == "c2" else syn.config_c4()
    est = pkg.setup_estimator(LIB, w)
    est.Solve(2)
    buf = (C.c_ulonglong * 512)()
    LIB.lib.ctvio_debug_dag_stamps(buf)
    t = np.array(buf[:], dtype=np.float64).reshape(32, 16)
    nb = (6 * w.n_knots + 6 * len(w.kf_times) + 1 + 63) // 64
    t0 = t[:nb, 0].min()
    print(name, "nb", nb, " columns (us from kernel start): start | old updates done | packets 0 1 2 3 of col j-1 seen | trsm done | D stored | "
          "factored (last packet posted) | fwd_ready posted | x_ready posted")
    for j in range(nb):
        v = t[j]
        cols = [v[0], v[1], v[8], v[9], v[10], v[11], v[3], v[4], v[5], v[6], v[7]]
        print("  col %2d: " % j + " ".join("%7.1f" % ((x - t0) / 1e3) if x > 0 else "      -" for x in cols))
    for j in (1, 3):
        v = t[16 + j]
        print("  col %d trsm steps (us from kernel start): [step a done | commit done | c done | d done] x 4: " % j + " | ".join(" ".join("%6.2f" % ((v[4 * s_ + k] - t0) / 1e3) for k in range(4)) for s_ in range(4)))
    fc = (C.c_longlong * 256)()
    LIB.lib.ctvio_debug_fac_clk(fc)
    fc = np.array(fc[:], dtype=np.int64).reshape(8, 4, 8)
    t00 = fc[0, 0, 0]
    print("  diag factor of column 0 (cycles since loop start), per 16-col step: loop top | P2a done | after barrier | before end barrier ; pub done (warps>0)")
    for s_ in range(4):
        print("   step %d: " % s_ + "  ".join("w%d %5d %5d %5d %5d p%5d" % (w_, fc[w_, s_, 0] - t00, fc[w_, s_, 1] - t00, fc[w_, s_, 2] - t00, fc[w_, s_, 3] - t00, fc[w_, s_, 4] - t00) for w_ in (0, 1, 2, 7)))
try:
    f = LIB.lib.ctvio_debug_chol_cluster_launches
    f.restype = C.c_longlong
    print("tile-DAG launches that used thread-block clusters:", f())
except AttributeError:
    pass
