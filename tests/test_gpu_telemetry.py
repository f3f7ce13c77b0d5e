"""End to end on the GPU: IQ -> characters -> sentences -> telemetry records and station statistics, with the tracker
attached to the batch decoder (hbd_attach_tracker) -- the chain DECODER_THREAD -> sentence_callback_ ->
SentenceCallback of the reference (websocketServer/main.cpp:238-245,292-366).  The records must equal what the
reference's own parse_sentence / CalcGpsDistance make of the oracle's sentences."""
import os

import numpy as np
import pytest

import telemetry_cases as tc
from habdec_b200 import api, synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

FS, BAUD, CHUNK = 2.048e6, 300.0, 65536
NOW = tc.CLOCKS[2]
STATION = (52.0, 21.0, 100.0)


def test_tracker_attached_to_decoder(oracle_kind):
    n_ch, n_sent = 6, 3
    chans = [synth.channel_iq(c, n_sent, FS, BAUD, snr_db=-15.0) for c in range(n_ch)]
    n = max(len(iq) for iq, _ in chans)
    n = (n + CHUNK - 1) // CHUNK * CHUNK
    iq = np.zeros((n_ch, n), dtype=np.complex64)
    for c, (x, _) in enumerate(chans):
        iq[c, :len(x)] = x

    dec = api.BatchDecoder(n_ch, baud=BAUD, rtty_bits=8, rtty_stops=2.0, dec_factor=256)
    tracker = api.Tracker(station=STATION, now_unix=NOW)
    dec.attach_tracker(tracker, ch_offset=100)
    seen = []
    tracker.set_callback(lambda ch, rec, s: seen.append((ch, s)))
    for o in range(0, n, CHUNK):
        dec.pushSamplesBatch(np.ascontiguousarray(iq[:, o:o + CHUNK]), FS)
        dec.process()

    total = 0
    for c in range(n_ch):
        ref = (po.RefDecoder if oracle_kind == "ref" else po.PortDecoder)(po.make_config(baud=BAUD, dec_factor=256)).run(iq[c], FS, CHUNK)
        sentences = ref.sentences()
        assert sentences == dec.poll_sentences(c) and len(sentences) == n_sent
        # the reference side of the same chain: its sentences through its own SentenceCallback arithmetic
        req = ["NOW\t%d" % NOW, "STATION\t%g\t%g\t%g" % STATION]
        for s in sentences:
            body, crc = s.decode().rstrip("\n").rsplit("*", 1)
            call, data = body.split(",", 1)
            req.append("CB\t%s\t%s\t%s" % (call, data, crc))
        if os.path.exists(tc.REF_BIN):
            want = tc.run_reference(req)
        else:       # GPU box without the compiled reference: the stand-alone tracker, itself pinned by the CPU tier
            o = tc.Ours(); want = [o.answer(r) for r in req]
        st = tracker.stats(100 + c); d = st.D_
        got_last = "\t".join([str(st.num_ok_)] + [tc.hx(x) for x in (d.dist_line_, d.dist_circle_, d.dist_radians_, d.elevation_, d.bearing_, st.dist_max_, st.elev_min_)]
                             + [tracker.stats_payload(100 + c).decode()])
        assert tc.same(got_last, want[-1]), (got_last, want[-1])
        recs = tracker.poll(100 + c)
        assert [r["frame"] for r in recs] == list(range(n_sent))
        assert all(r["payload_callsign"] == b"CH%04d" % c for r in recs)
        assert tracker.stats(c).num_ok_ == 0            # nothing filed under the un-offset id
        total += len(recs)
    assert [ch for ch, _ in seen].count(100) == n_sent and len(seen) == total
    dec.attach_tracker(None)
    dec.close()
