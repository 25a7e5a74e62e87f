#!/usr/bin/env bash
# Collect the round-2 profile artefacts on a GPU box (run under gpurun); outputs go to gpurun_out/, summaries are then
# copied into profiles/r2/ by hand.
set -x
cd "$(dirname "$0")/.."
# 1. launch list of the bench command itself (C2 headline only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-c4 --no-c3 --c5-windows 0 > gpurun_out/r2_launches_bench.log 2>&1
# 2. launch lists of a few LM iterations at C2 / C4
for w in c2 c4; do
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_$w.csv \
      python tools/ncu_target.py --workload $w --iters 3 --solves 2 > /dev/null 2>&1
done
# 3. full captures of the top kernels (last launches of the run: warm)
for w in c2 c4; do
  ncu --set full --clock-control none --import-source on -k regex:"chol_dag|visual_kernel|schur_tile" -s 20 -c 6 \
      -o gpurun_out/r2_full_$w python tools/ncu_target.py --workload $w --iters 3 --solves 2 > /dev/null 2>&1
done
ls -la gpurun_out/ | tail
