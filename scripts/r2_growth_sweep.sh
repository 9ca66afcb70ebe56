#!/bin/bash
# wave-growth sweep with the lazy refine (round 2): bench.py step time per (first wave, growth x10)
mkdir -p gpurun_out
for cfg in "131072 20" "131072 30" "131072 40" "131072 60" "131072 80" "262144 30" "262144 40" "524288 40"; do
  set -- $cfg
  SSS_WAVE_FIRST=$1 SSS_WAVE_GROWTH=$2 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra \
    > gpurun_out/r2a_growth_$1_$2.json 2> gpurun_out/r2a_growth_$1_$2.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2a_growth_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(j["ms_per_step"], 4), j["waves_per_step"], j["overflow_reruns"], round(j["roofline"]["kernel_ms_per_step"], 4))
    except Exception as e:
        print(f, "ERR", e)
PY
