#!/bin/bash
# A/B of the persistent bulk-copy-fed kernels (default) against the one-CTA-per-element kernels (B200_TMA=0), he30/ze63 Float32.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
  echo "== B200_TMA=0"; B200_TMA=0 QUICK=1 timeout 120 python tools/gpu_time_kernels.py 30
  echo "== default (persistent + cp.async.bulk)"; QUICK=1 timeout 120 python tools/gpu_time_kernels.py 30
} 2>&1 | tee gpurun_out/ab_tma.txt
