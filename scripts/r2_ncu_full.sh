#!/bin/bash
# ncu --set full captures for profiles/: the encoder GEMM launches and the fp8 (Hamming) scan
python __graft_entry__.py smoke > gpurun_out/r2m_smoke.log 2>&1; tail -2 gpurun_out/r2m_smoke.log
ncu --set full --import-source on --clock-control none -k regex:gemm_bf16x3 --launch-skip 16 --launch-count 8 \
  -o gpurun_out/r2m_gemm python scripts/encoder_prof_run.py 2 > gpurun_out/r2m_gemm.log 2>&1
cat > /tmp/bin_run.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import sessionsimilaritysearch_b200 as sss
g = torch.Generator(device='cuda').manual_seed(1)
ix = sss.IndexBinaryFlat(256)
for _ in range(5):
    ix.add(torch.randint(0, 256, (4_000_000, 32), generator=g, device='cuda', dtype=torch.uint8))
q = torch.randint(0, 256, (1000, 32), generator=g, device='cuda', dtype=torch.uint8)
for _ in range(3):
    ix.search(q, 100)
torch.cuda.synchronize()
print(ix.stats())
PY
python /tmp/bin_run.py > gpurun_out/r2m_bin_plain.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:scan_bf16_2cta --launch-skip 12 --launch-count 1 \
  -o gpurun_out/r2m_fp8scan python /tmp/bin_run.py > gpurun_out/r2m_bin.log 2>&1
tail -1 gpurun_out/r2m_bin_plain.log
