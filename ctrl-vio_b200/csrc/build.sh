#!/usr/bin/env bash
# Build libctvio_b200.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -euo pipefail
cd "$(dirname "$0")"
OUT=libctvio_b200.so
SRCS="engine.cu kernels_residual.cu kernels_linear.cu chol_coop.cu chol_dag.cu misc_kernels.cu marginalize.cu comm.cu"
HDRS="kernels.h chol_tiles.cuh spline_eval.cuh device_math.cuh marginalize.h poly_min.h ../../include/ctvio.h"
if [[ -z "${CTVIO_FORCE_BUILD:-}" && -f "$OUT" ]]; then
  newer=0
  for f in $SRCS $HDRS build.sh; do [[ "$f" -nt "$OUT" ]] && newer=1; done
  [[ $newer -eq 0 ]] && exit 0
fi
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
# NCCL: only <nccl.h> (types) is needed at build time; the library is dlopen'ed by comm.cu on first use.
FLAGS="${CTVIO_EXTRA_NVCC_FLAGS:-} -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -Xptxas -v"
objs=""
for f in $SRCS; do
  o="${f%.cu}.o"
  if [[ -n "${CTVIO_FORCE_BUILD:-}" || ! -f "$o" || "$f" -nt "$o" || kernels.h -nt "$o" || spline_eval.cuh -nt "$o" || device_math.cuh -nt "$o" || marginalize.h -nt "$o" || poly_min.h -nt "$o" || chol_tiles.cuh -nt "$o" || ../../include/ctvio.h -nt "$o" || build.sh -nt "$o" ]]; then
    $NVCC $FLAGS -c "$f" -o "$o" 2> "${o}.log" || { cat "${o}.log"; exit 1; }
  fi
  objs="$objs $o"
done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" $objs -lcudart -ldl
echo "built $OUT"
