#!/bin/bash
# All BASELINE.json configurations on one GPU (extra evidence next to the default bench line); writes gpurun_out/configs_1gpu.jsonl
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/configs_1gpu.jsonl
python bench.py --steps 20 --warmup 3 2>/dev/null | tail -1 >> gpurun_out/configs_1gpu.jsonl
for cfg in he16 moist tracer hs strong vdiff vdiff_implicit; do
  python bench.py --config $cfg --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/configs_1gpu.jsonl
done
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 >> gpurun_out/configs_1gpu.jsonl
python - <<'PY'
import json
for l in open("gpurun_out/configs_1gpu.jsonl"):
    try: d = json.loads(l)
    except Exception: print("??", l[:200]); continue
    print(d.get("impl", "b200"), d["config"]["workload"][:60], "| ms/step %.3f" % d["ms_per_step"], "| SYPD %.1f" % d["value"],
          "| e2e %.1f" % d["e2e"]["value"], "| step frac %.3f" % d.get("roofline_step", {}).get("frac", float("nan")),
          "| kernel frac %.3f" % d.get("roofline", {}).get("frac", float("nan")), "| launches", d.get("gpu_launches"))
PY
