#!/bin/bash
# DB-stream-bound regime (nq = 1, 32, 256): whole step and scan kernel against the HBM roofline
for nq in ${NQS:-1 32 128 256}; do
  echo -n "nq=$nq: "
  python bench.py --nq $nq --steps 20 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | python scripts/bench_line.py
done
