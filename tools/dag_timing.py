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
    print(name, "nb", nb, " columns: start | old-updates done | Linv(j-1) seen | diag-update gemm done | D stored + rhs | factored | diag_ready posted | x_ready posted  (us from kernel start)")
    for j in range(nb):
        print("  col %2d: " % j + " ".join("%7.1f" % ((v - t0) / 1e3) for v in t[j, :8]))
    j = min(3, nb - 1)
    d = t[j]
    print("  col %d slab detail (us): Linv^T+x loaded %.2f | slab gemm %.2f | smem stores %.2f | diagonal update gemm %.2f" % (
        j, (d[8] - d[2]) / 1e3, (d[9] - d[8]) / 1e3, (d[10] - d[9]) / 1e3, (d[3] - d[10]) / 1e3))
    fc = (C.c_longlong * 256)()
    LIB.lib.ctvio_debug_fac_clk(fc)
    fc = np.array(fc[:], dtype=np.int64).reshape(8, 4, 8)
    t00 = fc[0, 0, 0]
    print("  diag factor of column 0: arrival (cycles since start) of each warp's lane 0 BEFORE the barrier ending P1 | P2 | P3, per 16-col step")
    for s_ in range(4):
        print("   step %d: " % s_ + "  ".join("w%d %5d %5d %5d" % (w_, fc[w_, s_, 1] - t00, fc[w_, s_, 3] - t00, fc[w_, s_, 5] - t00) for w_ in (0, 1, 2, 3, 7)))
