"""debug: fp64 tensor-core (DMMA m8n8k4) throughput / latency on this GPU (needs tools/build_timing.sh)."""
import ctypes as C, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("ctrl-vio_b200")
LIB = pkg.CtvioLib(os.path.join(os.path.dirname(pkg.LIB_PATH), "libctvio_b200_timing.so"), "ctvio_")
for cps in (1, 2, 4, 8):
    tf, lat = C.c_double(), C.c_double()
    LIB.lib.ctvio_debug_dmma(cps, C.byref(tf), C.byref(lat))
    print(f"DMMA m8n8k4: {cps} CTAs(256 thr)/SM -> {tf.value:.1f} TFLOP/s, dependent latency {lat.value:.1f} cycles")
