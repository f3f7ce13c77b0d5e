"""Regenerate tests/golden/telemetry_golden.txt: request / answer pairs produced by the reference's own
parse_sentence / parse_sentence_time / parse_gps_pos / timestamp_from_HMS / CalcGpsDistance
(oracle/_ref/telemetry_ref, built by oracle/Makefile from /root/reference/code/common) with the clock frozen."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import telemetry_cases as tc  # noqa: E402

if __name__ == "__main__":
    req = tc.all_requests(seed=1, n_fuzz=1500)
    ans = tc.run_reference(req)
    with open(os.path.join(HERE, "telemetry_golden.txt"), "w", encoding="latin-1", newline="\n") as f:
        for r, a in zip(req, ans):
            f.write(r + "\n" + a + "\n")
    print(len(req), "cases")
