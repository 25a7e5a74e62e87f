"""debug: C5 streaming, GPU vs oracle window by window + time split of one window."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("ctrl-vio_b200"); syn = pkg.synthetic
st = importlib.import_module("ctrl-vio_b200.streaming")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
seq = st.config_c5_sequence(n + 1)
olib = pkg.CtvioLib(os.path.join(os.path.dirname(os.path.dirname(pkg.LIB_PATH)), "..", "oracle", "liboracle.so"), "ctvo_",
                    optional=pkg.binding.DEVICE_ONLY_SYMBOLS)
g = st.StreamingRunner(pkg.load(), seq, iters=8); o = st.StreamingRunner(olib, seq, iters=8)
for k in range(n):
    rg, ro = g.step(k), o.step(k)
    d = syn.qmul(syn.qconj(o.q), g.q); ang = (2 * np.arctan2(np.linalg.norm(d[:, :3], axis=1), np.abs(d[:, 3]))).max()
    print(f"win {k}: it {rg['iterations']}/{ro['iterations']} cost {rg['final_cost']:.6f}/{ro['final_cost']:.6f} "
          f"rel {abs(rg['final_cost']/ro['final_cost']-1):.1e} dp {np.abs(g.p-o.p).max():.2e} ang {ang:.2e} "
          f"dbias {np.abs(g.bias-o.bias).max():.2e} dld {abs(g.ld-o.ld):.2e} ms gpu {rg['ms']:.2f}")
# time split on the GPU engine
e = g.est
t0 = time.perf_counter(); s = e.Solve(8); t1 = time.perf_counter(); pr = e.SaveMarginalizationInfo(); t2 = time.perf_counter()
print(f"solve(8) {1e3*(t1-t0):.2f} ms (device {s.device_ms:.2f}), marginalize + get_prior {1e3*(t2-t1):.2f} ms")
