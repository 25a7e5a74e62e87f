"""ncu target: a few streaming windows (C5) and one C3 marginalization so that K7's kernels show up in the launch list."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("ctrl-vio_b200")
st = importlib.import_module("ctrl-vio_b200.streaming")
seq = st.config_c5_sequence(4)
r = st.StreamingRunner(pkg.load(), seq)
r.run(3)
print(r.records[-1])
e, _, _, _ = st.c3_window_a(pkg.load())
e.Solve(2)
pr = e.SaveMarginalizationInfo()
print("c3 prior", pr.n)
