#!/bin/bash
# per-shard cost of the strong-scaling runs, measured on one GPU: rows = 10M / N for N = 1, 2, 4, 8
for rows in 10000000 5000000 2500000 1250000; do
  echo -n "rows=$rows: "
  python bench.py --rows $rows --steps 20 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python scripts/bench_line.py
done
