#!/bin/bash
# ncu launch lists (per-kernel durations, serialised) of one search step at nq = 1000 and nq = 1
for nq in 1000 1; do
  python scripts/r2_step.py $nq 2 > gpurun_out/r2c_plain_$nq.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches_$nq.csv \
    python scripts/r2_step.py $nq 1 > gpurun_out/r2c_ncu_$nq.log 2>&1
done
tail -2 gpurun_out/r2c_plain_*.log
