#!/bin/bash
# sweep of the bootstrapped wave schedule (first wave rows x growth factor x10) on the headline workload
for first in ${FIRSTS:-131072}; do
  for g in ${GROWTHS:-20 25 30 35 40}; do
    echo -n "first=$first growth=$g: "
    SSS_WAVE_FIRST=$first SSS_WAVE_GROWTH=$g python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline $EXTRA 2>/dev/null | python scripts/bench_line.py
  done
done
