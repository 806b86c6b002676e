#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== parity default"; python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for i in 1 2 3; do echo "== repeat $i"; python -m pytest tests -m gpu -q -k "one_step or bitwise or hundred or drift or variants" 2>&1 | tail -2; done
for cfg in "" "B200_PDL=0"; do
  echo "== timing [$cfg]"; env $cfg QUICK=1 python tools/gpu_time_kernels.py 2>&1 | grep -E "implicit stage|step fused|finite|phase|dss"
done
} > gpurun_out/ab_imp.log 2>&1
tail -60 gpurun_out/ab_imp.log
