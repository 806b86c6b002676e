#!/bin/bash
# A/B of the fused implicit-stage variants (GPU box): parity tests per variant, then timings.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== parity default"; python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== parity GENERIC_NV"; B200_GENERIC_NV=1 python -m pytest tests -m gpu -x -q -k "step or fused or tracer or smoke" 2>&1 | tail -3
echo "== parity SOLVER=1"; B200_IMP_SOLVER=1 python -m pytest tests -m gpu -x -q -k "step or fused or tracer" 2>&1 | tail -3
echo "== parity SOLVER=0"; B200_IMP_SOLVER=0 python -m pytest tests -m gpu -x -q -k "step or fused or tracer" 2>&1 | tail -3
echo "== parity MINB=3"; B200_IMP_MINB=3 python -m pytest tests -m gpu -x -q -k "step or fused or tracer" 2>&1 | tail -3
for cfg in "" "B200_IMP_MINB=3" "B200_GENERIC_NV=1" "B200_IMP_SOLVER=1" "B200_IMP_KERNEL=2"; do
  echo "== timing [$cfg]"; env $cfg QUICK=1 python tools/gpu_time_kernels.py 2>&1 | grep -E "implicit stage|step fused|finite"
done
} > gpurun_out/ab_imp.log 2>&1
tail -60 gpurun_out/ab_imp.log
