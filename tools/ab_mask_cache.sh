#!/bin/bash
# A/B of the tail kernel's slicer mask cache (HBD_MASK_CACHE=0: masks rebuilt in every call): headline step and wideband step.
mkdir -p gpurun_out
for mc in 0 1; do
  HBD_MASK_CACHE=$mc python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-parity --no-strong > gpurun_out/mc_$mc.json 2> gpurun_out/mc_$mc.err
  python - "$mc" <<'P'
import json, sys
mc = sys.argv[1]
d = json.loads(open("gpurun_out/mc_%s.json" % mc).read().strip().splitlines()[-1])
w = d.get("wideband") or {}
print("mask_cache=%s value %.0f ms/step %.4f k1 %.4f frac %.3f | wideband ms/step %s k1 %s" % (
    mc, d["value"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], w.get("ms_per_step"), w.get("k1_nco_avg_ms")))
P
done
