"""CPU tier of the SSDV packet-sync row: the restated wrapper (oracle/ssdv_oracle.h) and the
product's host automaton (csrc/host_tail.cpp SsdvChannel, through the hbd_ssdv_host_replay hook) against the reference's
own SSDV_wraper_t compiled into oracle/_ref, plus known answers of the published packet test."""
import os
import zlib

import numpy as np
import pytest

import ssdv_cases
from habdec_b200 import api
from oracle import pyoracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ssdv_transcripts.npz")


# ---- GF(256) arithmetic written independently of oracle/ssdv_published.h (carry-less multiply, no tables) ----
def gf_mul(a, b):
    r = 0
    while b:
        if b & 1:
            r ^= a
        a <<= 1
        if a & 0x100:
            a ^= 0x187
        b >>= 1
    return r


def gf_pow(a, e):
    r = 1
    for _ in range(e % 255):
        r = gf_mul(r, a)
    return r


def test_crc32_and_reed_solomon_known_answers():
    rng = np.random.default_rng(5)
    payload = rng.integers(0, 256, 237, dtype=np.uint8).tobytes()
    p = po.ssdv_make_packet("HABDEC", 7, 3, payload)
    assert p[:2] == b"\x55\x66" and p[6] == 7 and p[7:9] == b"\x00\x03"
    # base-40 callsign, last character most significant (published encode_callsign)
    code = 0
    for c in reversed("HABDEC"):
        code = code * 40 + (ord(c) - ord("A") + 14)
    assert int.from_bytes(p[2:6], "big") == code
    assert zlib.crc32(p[1:220]) == int.from_bytes(p[220:224], "big")
    # the codeword (bytes 1..255, first byte = highest power) vanishes at alpha^(11 * (112 + i)), i = 0..31, alpha = 2
    for i in range(32):
        x = gf_pow(2, 11 * (112 + i))
        acc = 0
        for byte in p[1:]:
            acc = gf_mul(acc, x) ^ byte
        assert acc == 0
    q = po.ssdv_make_packet("HABDEC", 7, 3, payload, fec=False)
    assert q[1] == 0x67 and zlib.crc32(q[1:252]) == int.from_bytes(q[252:256], "big")
    assert po.ssdv_is_packet(p)[:2] == (0, 0) and po.ssdv_is_packet(q)[:2] == (0, 0)


def test_packet_test_corrects_up_to_16_symbols_and_rejects_more():
    rng = np.random.default_rng(6)
    for trial in range(60):
        p = ssdv_cases.random_packet(rng, fec=True)
        n_err = trial % 20 + 1
        bad = ssdv_cases.corrupt(p, n_err, rng, lo=2)
        v, e, c = po.ssdv_is_packet(bad)
        if n_err <= 16:
            assert (v, e, c) == (0, n_err, p)
        else:
            assert v == -1 and c == bad
    # a damaged type byte is overwritten before the FEC pass and does not count as an error
    p = ssdv_cases.random_packet(rng, fec=True)
    bad = bytearray(ssdv_cases.corrupt(p, 16, rng, lo=2)); bad[1] ^= 0x5A
    assert po.ssdv_is_packet(bytes(bad)) == (0, 16, p)
    # a clean no-FEC packet with one flipped payload byte is not rescued by the FEC path
    q = ssdv_cases.random_packet(rng, fec=False)
    assert po.ssdv_is_packet(ssdv_cases.corrupt(q, 1, rng, lo=20))[0] == -1
    # sanity checks of the header: zero width, MCU index beyond the image
    assert po.ssdv_is_packet(ssdv_cases.random_packet(rng, width16=0))[0] == -1
    assert po.ssdv_is_packet(ssdv_cases.random_packet(rng, width16=2, height16=2, mcu_id=4))[0] == -1
    assert po.ssdv_is_packet(ssdv_cases.random_packet(rng, width16=2, height16=2, mcu_id=3))[0] == 0
    assert po.ssdv_is_packet(ssdv_cases.random_packet(rng, width16=2, height16=2, mcu_id=7, flags=1))[0] == 0
    assert po.ssdv_is_packet(ssdv_cases.random_packet(rng, mcu_id=0, mcu_offset=205, fec=True))[0] == -1
    assert po.ssdv_is_packet(ssdv_cases.random_packet(rng, mcu_id=0xFFFF, mcu_offset=255, fec=True))[0] == 0


@pytest.mark.parametrize("seed", range(10, 22))
def test_restated_wrapper_equals_reference_wrapper(seed):
    if not po.available("ref"):
        pytest.skip("oracle/_ref not built")
    _, chunks = ssdv_cases.make_stream(seed)
    ev_ref, img_ref = ssdv_cases.wrapper_events("ref", chunks)
    ev_orc, img_orc = ssdv_cases.wrapper_events("orc", chunks)
    assert ev_ref == ev_orc and img_ref == img_orc
    assert len(ev_ref) >= 5


def _events_from_replay(chunks, stream):
    acc = ssdv_cases.accepted_windows(stream)
    got = api.ssdv_host_replay(chunks, acc)
    # per event: the image set after filing, rebuilt from the packets the replay returned
    return got, acc


@pytest.mark.parametrize("seed", [30, 31, 32, 33, 34, 35])
def test_host_automaton_equals_reference_wrapper(seed, oracle_kind):
    stream, chunks = ssdv_cases.make_stream(seed)
    ev, _ = ssdv_cases.wrapper_events(oracle_kind, chunks)
    got, _ = _events_from_replay(chunks, stream)
    assert [(g[0], g[1], g[2], g[3], g[4], g[5], g[7]) for g in got] == [(e[0], e[1], e[2], e[3], e[4], e[5], e[6]) for e in ev]


def test_golden_transcripts():
    g = np.load(GOLD)
    for seed in (1, 2, 3):
        stream = g["stream%d" % seed].tobytes()
        sizes = g["chunks%d" % seed]
        chunks, o = [], 0
        for n in sizes:
            chunks.append(stream[o:o + int(n)])
            o += int(n)
        want = [(int(r[0]), str(c), int(r[1]), int(r[2]), int(r[3]), int(r[4]), int(r[5]), int(r[6])) for r, c in zip(g["events%d" % seed], g["callsigns%d" % seed])]
        ev, _ = ssdv_cases.wrapper_events("orc", chunks)
        assert ev == want
        got, _ = _events_from_replay(chunks, stream)
        assert [(x[0], x[1], x[2], x[3], x[4], x[5], x[7]) for x in got] == [w[:7] for w in want]
