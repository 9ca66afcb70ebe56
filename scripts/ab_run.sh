#!/bin/bash
# ABAB comparison on ONE box: the in-tree library against ab/<name>.so (bench.py without extras)
name=$1; shift
lib=sessionsimilaritysearch_b200/libsss_b200.so
cp $lib /tmp/lib_main.so
for rep in 1 2 3; do
  cp /tmp/lib_main.so $lib; echo -n "main   : "; python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline "$@" 2>/dev/null | python scripts/bench_line.py | cut -c1-150
  cp ab/$name.so $lib;      echo -n "$name: "; python bench.py --steps 20 --warmup 3 --no-extra --no-cpu-baseline "$@" 2>/dev/null | python scripts/bench_line.py | cut -c1-150
done
cp /tmp/lib_main.so $lib
