#!/bin/bash
# encoder: parity tests, timing, ncu launch list
timeout 600 python -m pytest tests/test_encoder_gpu.py tests/test_metrics_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 120 python scripts/encoder_prof_run.py 6 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2h_enc_launches.csv python scripts/encoder_prof_run.py 3 > gpurun_out/r2h_enc_ncu.log 2>&1
