#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== parity default"; python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for cfg in ""; do
  echo "== timing [$cfg]"; env $cfg QUICK=1 python tools/gpu_time_kernels.py 2>&1 | grep -E "phase|step fused|finite"
done
} > gpurun_out/ab_imp.log 2>&1
tail -60 gpurun_out/ab_imp.log
