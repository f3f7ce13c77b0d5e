"""Reads bench.py JSON lines on stdin and prints a one-line summary each (helper for gpurun sessions)."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline") or {}
    e = d.get("e2e") or {}
    print(sys.argv[1] if len(sys.argv) > 1 else "", "impl", d.get("impl", "ours"), "n_gpus", d.get("n_gpus"), "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4),
          "| k1_ms", r.get("avg_launch_ms") and round(r["avg_launch_ms"], 4), "x", r.get("launches_timed"), "frac", r.get("frac") and round(r["frac"], 3),
          "rest_ms", r.get("rest_of_step_ms") and round(r["rest_of_step_ms"], 4), "gaps", [v and round(v, 4) for v in (r.get("pipeline_gaps") or {}).values()], "| e2e", e.get("value") and round(e["value"]),
          "| launches", d.get("gpu_launches"), "host_ms", d.get("host_issue_ms_per_step") and round(d["host_issue_ms_per_step"], 4), "final_collect_ms", d.get("final_collect_ms") and round(d["final_collect_ms"], 3), "| clocks", d.get("clocks"), "| results", d.get("results"), "| cpu", (d.get("cpu_baseline") or {}).get("value"))
