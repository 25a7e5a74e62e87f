#!/usr/bin/env bash
# Build libctvio_b200.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -euo pipefail
cd "$(dirname "$0")"
OUT=libctvio_b200.so
SRCS="engine.cu kernels_residual.cu kernels_linear.cu chol_coop.cu misc_kernels.cu marginalize.cu comm.cu"
HDRS="kernels.h spline_eval.cuh device_math.cuh marginalize.h poly_min.h ../../include/ctvio.h"
if [[ -z "${CTVIO_FORCE_BUILD:-}" && -f "$OUT" ]]; then
  newer=0
  for f in $SRCS $HDRS build.sh; do [[ "$f" -nt "$OUT" ]] && newer=1; done
  [[ $newer -eq 0 ]] && exit 0
fi
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
# NCCL: headers from the system package; at run time libnccl.so.2 resolves to the copy torch already loaded
NCCL_LIB_DIR=${NCCL_LIB_DIR:-$(python - <<'PY'
import os, glob
c = glob.glob('/usr/lib/x86_64-linux-gnu/libnccl.so*')
if c:
    print(os.path.dirname(c[0]))
else:
    import importlib.util
    s = importlib.util.find_spec('nvidia.nccl')
    print(os.path.join(list(s.submodule_search_locations)[0], 'lib') if s else '')
PY
)}
FLAGS="${CTVIO_EXTRA_NVCC_FLAGS:-} -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -Xptxas -v"
objs=""
for f in $SRCS; do
  o="${f%.cu}.o"
  if [[ -n "${CTVIO_FORCE_BUILD:-}" || ! -f "$o" || "$f" -nt "$o" || kernels.h -nt "$o" || spline_eval.cuh -nt "$o" || device_math.cuh -nt "$o" || marginalize.h -nt "$o" || poly_min.h -nt "$o" || ../../include/ctvio.h -nt "$o" || build.sh -nt "$o" ]]; then
    $NVCC $FLAGS -c "$f" -o "$o" 2> "${o}.log" || { cat "${o}.log"; exit 1; }
  fi
  objs="$objs $o"
done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" $objs -L"$NCCL_LIB_DIR" -l:libnccl.so.2 -lcudart
echo "built $OUT"
