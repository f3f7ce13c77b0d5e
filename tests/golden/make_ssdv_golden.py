"""Golden SSDV transcripts from the reference's own SSDV_wraper_t (oracle/_ref; the packet test below it is the
published-algorithm restatement oracle/ssdv_published.h -- fsphil/ssdv itself is absent).  Dev container only:

    python tests/golden/make_ssdv_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import ssdv_cases  # noqa: E402

out = {}
for seed in (1, 2, 3):
    stream, chunks = ssdv_cases.make_stream(seed, 60)
    ev, images = ssdv_cases.wrapper_events("ref", chunks)
    out["stream%d" % seed] = np.frombuffer(stream, dtype=np.uint8)
    out["chunks%d" % seed] = np.array([len(c) for c in chunks], dtype=np.int32)
    out["events%d" % seed] = np.array([(e[0], e[2], e[3], e[4], e[5], e[6], e[7]) for e in ev], dtype=np.int64).reshape(-1, 7)
    out["callsigns%d" % seed] = np.array([e[1] for e in ev])
    print("seed", seed, len(stream), "bytes", len(chunks), "chunks", len(ev), "events")
np.savez_compressed(os.path.join(HERE, "ssdv_transcripts.npz"), **out)
