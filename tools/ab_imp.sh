#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for cfg in "" "B200_IMP_MINB=3"; do
  echo "== timing [$cfg]"; env $cfg QUICK=1 python tools/gpu_time_kernels.py 2>&1 | grep -E "implicit stage|step fused|finite"
done
} > gpurun_out/ab_imp.log 2>&1
tail -60 gpurun_out/ab_imp.log
