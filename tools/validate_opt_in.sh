#!/bin/bash
# First GPU call of the next round: run the GPU tests of everything that has only been checked in the CPU CTA emulator so far, then time
# each opt-in variant next to the validated default (he30/ze63 Float32).  Usage:
#   gpurun --timeout 300 -- 'bash tools/validate_opt_in.sh'
# Output: gpurun_out/opt_in_pytest.log, gpurun_out/opt_in_timing.txt, gpurun_out/vdiff_timing.jsonl
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B200_RUN_UNVALIDATED=1 timeout 200 python -m pytest tests/test_gpu_vertical_diffusion.py -m gpu -q -k "borrowing or pcr or second_generation or fused_implicit" \
  2>&1 | tee gpurun_out/opt_in_pytest.log | tail -5
{
  echo "== implicit diffusion, hook sequence, validated default (k_vdiff_jac + k_ldiv_diff)"; timeout 60 python tools/gpu_time_vdiff.py 2>&1 | grep implicit
  echo "== B200_LDIV_DIFF=2 (k_vdiff_jac2 + k_ldiv_diff2, PCR)"; B200_LDIV_DIFF=2 timeout 60 python tools/gpu_time_vdiff.py 2>&1 | grep implicit
  echo "== B200_LDIV_DIFF=2 B200_HOOK_KERNELS=2"; B200_LDIV_DIFF=2 B200_HOOK_KERNELS=2 timeout 60 python tools/gpu_time_vdiff.py 2>&1 | grep implicit
  echo "== B200_VDIFF_FUSED=1 (k_imp_stage_diff in the fused, graph-replayed step)"; B200_VDIFF_FUSED=1 timeout 60 python tools/gpu_time_vdiff.py 2>&1 | grep implicit
  echo "== dry hook-by-hook step, first-generation hook kernels"; timeout 120 python bench.py --unfused --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], 'ms/step', d['gpu_launches'], 'launches')"
  echo "== dry hook-by-hook step, B200_HOOK_KERNELS=2"; B200_HOOK_KERNELS=2 timeout 120 python bench.py --unfused --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], 'ms/step', d['gpu_launches'], 'launches')"
} 2>&1 | tee gpurun_out/opt_in_timing.txt
