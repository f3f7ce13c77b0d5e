"""Telemetry layer through the C ABI vs the reference's own parse_sentence / CalcGpsDistance:
committed golden transcript (tests/golden/telemetry_golden.txt, made by tests/golden/make_telemetry_golden.py from
oracle/_ref/telemetry_ref) and, where the compiled reference is present, fresh fuzz transcripts.  Bit exact: floats are
compared as C hex floats.  Host-only code: runs without a GPU."""
import os
import time

import pytest

import telemetry_cases as tc
from habdec_b200 import api

HERE = os.path.dirname(os.path.abspath(__file__))


def _golden():
    with open(os.path.join(HERE, "golden", "telemetry_golden.txt"), encoding="latin-1", newline="") as f:
        lines = f.read().split("\n")[:-1]
    return lines[0::2], lines[1::2]


def _compare(req, ans):
    ours = tc.Ours()
    bad = []
    kinds = {}
    for r, a in zip(req, ans):
        o = ours.answer(r)
        kinds[(r.split("\t")[0], a.split("\t")[0] if a in ("NONE", "THROW", "OK") else "value")] = kinds.get((r.split("\t")[0], "x"), 0) + 1
        if not tc.same(o, a):
            bad.append((r, a, o))
    assert not bad, "%d of %d differ, first: %r" % (len(bad), len(req), bad[:3])
    return kinds


def test_golden_transcript():
    req, ans = _golden()
    assert len(req) > 1500
    kinds = _compare(req, ans)
    # the transcript exercises every outcome of every entry point
    for k in [("TIME", "value"), ("TIME", "NONE"), ("POS", "value"), ("POS", "THROW"), ("SENT", "value"), ("SENT", "NONE"), ("SENT", "THROW"),
              ("CB", "value"), ("CB", "NONE"), ("CB", "THROW"), ("DIST", "value"), ("STAMP", "value")]:
        assert k in kinds, k


def test_golden_is_current():
    """The committed requests are what the generator produces today (the cases module and the fixture stay in step)."""
    req, _ = _golden()
    assert req == tc.all_requests(seed=1, n_fuzz=1500)


@pytest.mark.skipif(not os.path.exists(tc.REF_BIN), reason="compiled reference (oracle/_ref/telemetry_ref) not present")
@pytest.mark.parametrize("seed", [11, 12, 13])
def test_fuzz_vs_compiled_reference(seed):
    req = ["NOW\t%d" % tc.CLOCKS[seed % 5], "STATION\t50.5\t19.25\t300"] + tc.fuzz_requests(seed, 3000)
    _compare(req, tc.run_reference(req))


def test_midnight_window():
    """timestamp_from_HMS (sentence_parse.cpp:72-98): hour 23 seen at 00h is yesterday, hour 0 seen at 23h is tomorrow."""
    just_before, just_after = tc.CLOCKS[0], tc.CLOCKS[1]          # 2026-10-17 23:59:59 / 2026-10-18 00:01:40 UTC
    assert api.timestamp_from_hms(0, 0, 1, just_before) == b"2026-10-18T00:00:01Z"
    assert api.timestamp_from_hms(23, 59, 59, just_before) == b"2026-10-17T23:59:59Z"
    assert api.timestamp_from_hms(23, 59, 59, just_after) == b"2026-10-17T23:59:59Z"
    assert api.timestamp_from_hms(0, 0, 1, just_after) == b"2026-10-18T00:00:01Z"
    assert api.timestamp_from_hms(22, 0, 0, just_after) == b"2026-10-18T22:00:00Z"
    assert api.timestamp_from_hms(12, 3, 5.5, just_after) == b"2026-10-18T12:03:5.5Z"     # the reference's setw(2) on a float
    # system clock variant: same date as the frozen variant evaluated around the call
    t0 = int(time.time()); got = api.timestamp_from_hms(12, 0, 0); t1 = int(time.time())
    assert got in (api.timestamp_from_hms(12, 0, 0, t0), api.timestamp_from_hms(12, 0, 0, t1))


def test_tracker_bookkeeping():
    """sentences_map_ keyed by frame id (a repeated id replaces), num_ok_ = distinct ids, dist_max_/elev_min_ running
    extremes, no distance without a station latitude (websocketServer/main.cpp:339-366)."""
    t = api.Tracker(station=(52.0, 21.0, 100.0), now_unix=tc.CLOCKS[2])
    got = []
    t.set_callback(lambda ch, rec, s: got.append((ch, rec["frame"], s)))
    assert t.push(5, b"CALL", b"1,12:00:00,52.1,21.5,1000", b"AAAA") == 1
    assert t.push(5, b"CALL", b"2,12:00:10,52.3,21.9,9000", b"BBBB") == 1
    d_far = t.stats(5).D_.dist_line_
    assert t.push_sentence(5, b"CALL,2,12:00:20,52.2,21.6,5000*CCCC\n") == 1
    assert t.push(5, b"CALL", b"x,12:00:00,52.1,21.5,1000", b"DDDD") == -1
    assert t.push(5, b"CALL", b"3,12:00:00,0,0,1000", b"EEEE") == 0
    st = t.stats(5)
    assert st.num_ok_ == 2 and st.dist_max_ == d_far and st.D_.dist_line_ < d_far and 0 < st.elev_min_ < 90
    assert t.get_sentence(5, 2) == b"CALL,2,12:00:20,52.2,21.6,5000*CCCC" and t.get_sentence(5, 9) == b""
    assert [g[:2] for g in got] == [(5, 1), (5, 2), (5, 2)] and got[2][2] == b"CALL,2,12:00:20,52.2,21.6,5000*CCCC"
    recs = t.poll(5)
    assert [r["frame"] for r in recs] == [1, 2, 2] and t.poll(5) == [] and t.poll(6) == []
    assert recs[0]["tracking"] == b"CALL,2026-10-18T12:00:00Z,52.1,21.5,1000"
    assert t.stats(6).num_ok_ == 0 and t.stats(6).elev_min_ == 90.0
    assert t.stats_payload(6) == b"cmd::info:stats=ok:0,dist_line:0,dist_circ:0,max_dist:0,min_elev:90,lat:52,lon:21,alt:100"
    assert t.stats_payload(5, with_age=True).endswith(b",age:0")
    # channels are independent; a station at latitude 0 switches the distance off like `if(station_lat_)`
    t2 = api.Tracker(station=(0.0, 21.0, 100.0), now_unix=tc.CLOCKS[2])
    assert t2.push(1, b"CALL", b"1,12:00:00,52.1,21.5,1000", b"AAAA") == 1
    assert t2.stats(1).num_ok_ == 1 and t2.stats(1).dist_max_ == 0.0
