#!/bin/bash
# every shard of a W-way sharded run, one after the other on ONE GPU (finds the shard that sets the max-over-ranks time)
rows=${1:-100000000}; W=${2:-8}
for r in $(seq 0 $((W-1))); do
  echo -n "shard $r/$W: "
  python bench.py --rows $rows --shard $r/$W --steps 3 --warmup 2 --no-extra --no-cpu-baseline 2>/dev/null | python scripts/bench_line.py | cut -c1-220
done
