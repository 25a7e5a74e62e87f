"""ncu target: a few streaming windows (C5) so that the marginalization kernels show up in the launch list."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("ctrl-vio_b200")
st = importlib.import_module("ctrl-vio_b200.streaming")
seq = st.config_c5_sequence(4)
r = st.StreamingRunner(pkg.load(), seq, iters=8)
r.run(3)
print(r.records[-1])
