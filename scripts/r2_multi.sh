#!/bin/bash
# multi-GPU validation: sharded parity tests + bench at N ranks (N = number of visible GPUs)
N=$(python -c "import torch; print(torch.cuda.device_count())")
python -m pytest tests/test_dist_gpu.py tests/test_search_gpu.py::test_two_handles_on_two_devices_in_one_process -q -m gpu 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py \
  --gpus $N --steps 20 --warmup 3 --configs ${CONFIGS:-headline} > gpurun_out/r2e_bench_n$N.json 2> gpurun_out/r2e_bench_n$N.err
echo rc=$?
tail -3 gpurun_out/r2e_bench_n$N.err
python - <<PY
import json
j=json.loads(open("gpurun_out/r2e_bench_n$N.json").read().strip().splitlines()[-1])
for k in ["value","ms_per_step","profiled_ms_per_step","graph_replay","waves_per_step","parity","e2e","per_rank_ms_per_step"]: print(k, j.get(k))
print({k:j["roofline"][k] for k in ("frac","frac_burst","step_frac","kernel_ms_per_step","kernel_share_of_step")})
print(json.dumps(j.get("configs"), indent=1)[:3000])
PY
