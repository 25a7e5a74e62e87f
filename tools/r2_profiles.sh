#!/usr/bin/env bash
# Collect the round-2 profile artefacts on a GPU box (run under gpurun); outputs go to gpurun_out/, tools/ncu_summary.py
# (run in the build container, where the reports are merged back) turns them into profiles/r2/*.
set -x
cd "$(dirname "$0")/.."
export CTVIO_CHOL_CLUSTER=0   # ncu cannot replay the cooperative cluster launch of K5 (LaunchFailed): profile its L2-path variant
NCU="ncu --clock-control none"
# 1. launch list of the bench command itself (C2 headline only)
timeout -s KILL 400 $NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-c4 --no-c3 --c5-windows 0 > gpurun_out/r2_launches_bench.log 2>&1
# 2. launch lists of a few LM iterations at C2 / C4, and of one streaming window incl. its marginalization
for w in c2 c4; do
  timeout -s KILL 300 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_launches_$w.csv \
      python tools/ncu_target.py --workload $w --iters 3 --solves 2 > /dev/null 2>&1
done
timeout -s KILL 300 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r2_launches_c5_marginalize.csv \
    python tools/c5_ncu_target.py > /dev/null 2>&1
# 3. full captures of the top kernels (launches of the second, warm solve)
for w in c2 c4; do
  timeout -s KILL 400 $NCU --set full --import-source on -k regex:"chol_dag|visual_kernel|schur_tile" -s 10 -c 9 \
      -o gpurun_out/r2_full_$w python tools/ncu_target.py --workload $w --iters 3 --solves 2 > /dev/null 2>&1
done
timeout -s KILL 400 $NCU --set full --import-source on -k regex:"jacobi_blocked" -s 4 -c 4 \
    -o gpurun_out/r2_full_c5 python tools/c5_ncu_target.py > /dev/null 2>&1
ls -la gpurun_out/ | tail
