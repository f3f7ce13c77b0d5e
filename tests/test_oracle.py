"""CPU tier: the oracle itself.

 * the CPU restatement (oracle/habdec_oracle.cpp) against the golden fixtures that were produced by the
   UNMODIFIED reference (tests/golden/make_golden.py) -- bit-exact floats, identical characters;
 * where the compiled reference is present (oracle/_ref, dev container and shipped .so): restatement vs
   reference on fresh seeded inputs, every stage.
"""
import glob
import hashlib
import os

import numpy as np
import pytest

from habdec_b200 import synth
from oracle import pyoracle as po

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "g[0-9]*.npz")))
SCALE = 32.0


def load_gold(path):
    g = np.load(path)
    q = g["iq_q"]
    iq = (q[:, 0].astype(np.float32) / np.float32(SCALE) + 1j * (q[:, 1].astype(np.float32) / np.float32(SCALE))).astype(np.complex64)
    cfg = dict(baud=float(g["baud"]), rtty_bits=int(g["bits"]), rtty_stops=float(g["stops"]), dec_factor=int(g["factor"]),
               dc_remove=bool(g["dc_remove"]))
    return g, iq, cfg


SETTERS = ["baud", "rtty_bits", "rtty_stops", "dc_remove", "lowpass_bw", "lowpass_trans"]


def gold_calls(g, n_total):
    """[(chunk size, [(setter, value), ...] applied before the call)] of a fixture (fixed chunk, or `chunks` + `events`)."""
    if "chunks" not in g.files:
        c = int(g["chunk"])
        return [(min(c, n_total - o), []) for o in range(0, n_total, c)]
    ev = g["events"]
    return [(int(n), [(SETTERS[int(w)], float(v)) for call, w, v in ev if int(call) == i]) for i, n in enumerate(g["chunks"])]


def drive_gold(d, g, iq):
    o = 0
    for n, events in gold_calls(g, len(iq)):
        for name, v in events:
            d.set_param(name, v)
        d.push_process(iq[o:o + n], float(g["fs"]))
        o += n
    return d


def bits_equal(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_port_matches_golden_reference_vectors(path):
    g, iq, cfg = load_gold(path)
    d = drive_gold(po.PortDecoder(po.make_config(**cfg)), g, iq)
    assert bits_equal(d.stage(po.STAGE_LPTAPS), g["lptaps"])
    assert bits_equal(d.stage(po.STAGE_DECIMATED), g["decimated"])
    assert bits_equal(d.stage(po.STAGE_FILTERED), g["filtered"])
    assert bits_equal(d.stage(po.STAGE_DEMOD), g["demod"])
    assert bits_equal(d.stage(po.STAGE_PENDING), g["pending"])
    assert hashlib.sha256(d.stage(po.STAGE_FFT).tobytes()).hexdigest() == str(g["fft_sha"])
    assert bits_equal(d.stage(po.STAGE_POWER), g["power"])
    assert d.chars() == g["chars"].tobytes()
    assert d.rtty() == g["rtty"].tobytes()
    assert d.last_sentence() == g["last_sentence"].tobytes()
    assert b"\n".join(d.sentences()) == g["sentences"].tobytes()
    a = d.afc()
    assert [a.frequency_correction, a.shift_hz, a.noise_floor, a.noise_variance, a.peak_left, a.peak_right] == list(g["afc"])


def test_golden_set_is_meaningful():
    assert len(GOLD) >= 4 and any("chunks" in np.load(p).files for p in GOLD)
    g, _, _ = load_gold([p for p in GOLD if "g2_" in p][0])
    assert g["sentences"].tobytes().startswith(b"CH0011,0,12:00:00") and len(g["chars"]) > 40


needs_ref = pytest.mark.skipif(not po.available("ref"), reason="oracle/_ref/libhabdec_ref.so not built (needs /root/reference)")

FRESH = [
    # fs, baud, bits, stops, factor, snr, nsent, chunk, f_off, dc
    (2.048e6, 300.0, 8, 2.0, 256, -15.0, 1, 65536, 0.0, False),
    (2.048e6, 300.0, 8, 2.0, 256, -27.0, 1, 65536, 0.0, False),
    (2.048e6, 300.0, 8, 2.0, 256, -15.0, 1, 262144, 90.0, True),
    (1.024e6, 100.0, 7, 1.0, 128, -14.0, 1, 65536, 0.0, False),
    (0.512e6, 300.0, 8, 1.5, 64, -12.0, 1, 50000, 0.0, False),
    (0.128e6, 300.0, 8, 2.0, 16, -8.0, 2, 65536, 0.0, False),
    (64e3, 300.0, 8, 2.0, 8, -6.0, 2, 65536, -40.0, False),
    (32e3, 600.0, 8, 2.0, 4, -3.0, 2, 16384, 0.0, False),
]


@needs_ref
@pytest.mark.parametrize("fs,baud,bits,stops,factor,snr,nsent,chunk,foff,dc", FRESH)
def test_port_matches_compiled_reference(fs, baud, bits, stops, factor, snr, nsent, chunk, foff, dc):
    iq, _ = synth.channel_iq(21, nsent, fs, baud, bits, int(np.ceil(stops)), snr_db=snr, f_off=foff)
    cfg = dict(baud=baud, rtty_bits=bits, rtty_stops=stops, dec_factor=factor, dc_remove=dc)
    r = po.RefDecoder(po.make_config(**cfg)).run(iq, fs, chunk)
    p = po.PortDecoder(po.make_config(**cfg)).run(iq, fs, chunk)
    for st in (po.STAGE_LPTAPS, po.STAGE_DECIMATED, po.STAGE_FILTERED, po.STAGE_DEMOD, po.STAGE_PENDING, po.STAGE_FFT, po.STAGE_POWER):
        assert bits_equal(r.stage(st), p.stage(st)), "stage %d" % st
    assert r.chars() == p.chars() and len(r.chars()) > 10
    assert r.sentences() == p.sentences()
    assert r.rtty() == p.rtty() and r.last_sentence() == p.last_sentence()
    ra, pa = r.afc(), p.afc()
    for f, _ in po.AfcInfo._fields_:
        assert getattr(ra, f) == getattr(pa, f), f


@needs_ref
def test_reference_afc_reset_matches_port():
    fs, baud = 2.048e6, 300.0
    iq, _ = synth.channel_iq(4, 2, fs, baud, snr_db=-12.0, f_off=150.0)
    r = po.RefDecoder(po.make_config(baud=baud)); p = po.PortDecoder(po.make_config(baud=baud))
    half = len(iq) // 2 // 65536 * 65536
    for d in (r, p):
        d.run(iq[:half], fs)
        d.reset_frequency_correction(d.afc().frequency_correction)
        d.run(iq[half:], fs)
    ra, pa = r.afc(), p.afc()
    for f, _ in po.AfcInfo._fields_:
        assert getattr(ra, f) == getattr(pa, f), f


def test_premix_oracle_is_phase_continuous_and_invertible():
    """The NCO oracle (pyoracle.premix): chunked == whole, and mixing back restores the input."""
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(50000) + 1j * rng.standard_normal(50000)).astype(np.complex64)
    fs, f = 2.048e6, 37123.456
    whole, ph_end = po.premix(x, fs, f)
    ph, parts = 0.0, []
    for o in range(0, len(x), 7777):
        y, ph = po.premix(x[o:o + 7777], fs, f, ph)
        parts.append(y)
    chunked = np.concatenate(parts)
    assert abs(ph - ph_end) < 1e-9
    assert np.max(np.abs(chunked - whole)) < 1e-5
    back, _ = po.premix(whole, fs, -f)
    assert np.linalg.norm(back - x) / np.linalg.norm(x) < 1e-6
    # a pure tone at +f lands on DC
    tone = np.exp(2j * np.pi * f / fs * np.arange(4096)).astype(np.complex64)
    dc, _ = po.premix(tone, fs, f)
    assert np.max(np.abs(dc - 1.0)) < 1e-5


def test_wire_format_restatement_matches_reference_serialisation():
    """PWR_/DEM_ payloads: restatement == golden bytes produced by the reference's own SerializeSpectrum /
    CompressedVector code (tests/golden/make_wire_golden.py), and == the live reference build when present."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_wire_golden as g
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "wire_frames.npz"))
    port = g.frames("orc")
    assert set(port) == set(gold.files)
    for k in gold.files:
        assert port[k].tobytes() == gold[k].tobytes(), k
        assert len(gold[k]) > 20
    if po.available("ref"):
        live = g.frames("ref")
        for k in gold.files:
            assert live[k].tobytes() == gold[k].tobytes(), k


def test_port_equals_reference_with_setters_between_calls():
    """baud / rtty_bits / rtty_stops / dc_remove changed mid-stream (Decoder.h:654-706): restatement == reference."""
    if not po.available("ref"):
        pytest.skip("oracle/_ref not built")
    fs = 2.048e6
    a, _ = synth.channel_iq(31, 2, fs, 300.0, snr_db=-15.0)
    b, _ = synth.channel_iq(32, 2, fs, 100.0, nbits=7, nstops=1, snr_db=-15.0)
    a = a[:len(a) // 65536 * 65536]
    b = b[:len(b) // 65536 * 65536]
    iq = np.concatenate([a, b, a])
    marks = {len(a) // 65536: [("baud", 100.0), ("rtty_bits", 7), ("rtty_stops", 1.0), ("dc_remove", 1)],
             (len(a) + len(b)) // 65536: [("baud", 300.0), ("rtty_bits", 8), ("rtty_stops", 2.0), ("dc_remove", 0)]}
    outs = []
    for K in (po.RefDecoder, po.PortDecoder):
        d = K(po.make_config(baud=300.0))
        for i, o in enumerate(range(0, len(iq), 65536)):
            for k, v in marks.get(i, []):
                d.set_param(k, v)
            d.push_process(iq[o:o + 65536], fs)
        outs.append((d.chars(), d.sentences(), d.stage(po.STAGE_DEMOD), d.stage(po.STAGE_DECIMATED)))
    assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1]
    assert np.array_equal(outs[0][2], outs[1][2]) and np.array_equal(outs[0][3], outs[1][3])
    assert len(outs[0][1]) == 6


@pytest.mark.skipif(not po.available("ref"), reason="oracle/_ref not built")
def test_restatement_lowpass_work_buffer_equals_reference():
    """The low-pass keeps its work buffer across tap-count changes (FirFilter.h:139-160): shrink, grow back, grow after a
    larger call, grow past the buffer.  Restatement and compiled reference agree bit for bit on the filtered stream."""
    marks = {5: 0.05, 9: 0.025, 13: 0.1, 14: 0.0125, 22: 0.02, 24: 0.01}
    sizes = {18: 131072, 19: 131072}
    fs, baud = 2.048e6, 300.0
    iq, _ = synth.channel_iq(43, 2, fs, baud, snr_db=-15.0)
    out = {}
    for kind, cls in (("ref", po.RefDecoder), ("orc", po.PortDecoder)):
        d = cls(po.make_config(baud=baud))
        o = i = 0
        taps = set()
        while o + sizes.get(i, 65536) <= len(iq):
            if i in marks:
                d.set_param("lowpass_trans", marks[i])
            n = sizes.get(i, 65536)
            d.push_process(iq[o:o + n], fs)
            taps.add(len(d.stage(po.STAGE_LPTAPS)))
            o += n; i += 1
        out[kind] = (d.stage(po.STAGE_FILTERED), d.chars(), taps)
    assert {161, 81, 41, 257, 201, 321} <= out["ref"][2] == out["orc"][2]
    assert np.array_equal(out["ref"][0].view(np.uint32), out["orc"][0].view(np.uint32))
    assert out["ref"][1] == out["orc"][1]
