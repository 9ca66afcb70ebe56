#!/bin/bash
# small-batch (DB-stream-bound) regime: step time at nq = 1 / 32 / 128 for several wave growth factors
for g in 30 60 90 150; do for nq in 1 128; do
  echo "growth $g nq $nq: $(SSS_WAVE_GROWTH=$g python scripts/r2_step.py $nq 2 2>&1 | tail -1 | cut -c1-120)"
done; done
