#!/bin/bash
# first GPU trip: goldens from the reference, parity report, quick timing
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
ls /root/reference 2>&1 | head -3
python tools/make_golden.py --out gpurun_out/golden 2>&1 | tail -20
python tools/gpu_check.py --golden gpurun_out/golden 2>&1 | tail -60
