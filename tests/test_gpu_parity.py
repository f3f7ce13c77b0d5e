"""GPU parity: the CUDA path behind the C ABI vs the CPU oracle on identical synthetic IQ.

Bars (BASELINE.json north_star): characters / sentences / CRC verdicts bit-exact; per-stage float
arrays within relative L2 <= 1e-5 of the reference's own float path.
"""
import numpy as np
import pytest

from habdec_b200 import api, synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

REL_L2 = 1e-5


def rel_l2(a, b):
    a = np.asarray(a).astype(np.complex128 if np.iscomplexobj(a) else np.float64)
    b = np.asarray(b).astype(a.dtype)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def make_oracle(kind, **cfg):
    return (po.RefDecoder if kind == "ref" else po.PortDecoder)(po.make_config(**cfg))


def run_gpu_single(iq, fs, chunk, **cfg):
    """One channel through the batch decoder with stage recording; returns dict of concatenated stages."""
    dec = api.BatchDecoder(1, record=True, **cfg)
    stages = {"dec": [], "filt": [], "demod": []}
    for o in range(0, len(iq), chunk):
        dec.pushSamples(0, iq[o:o + chunk], fs)
        dec.process()
        stages["dec"].append(dec.debug_stage(0, api.STAGE_DECIMATED).copy())
        stages["filt"].append(dec.debug_stage(0, api.STAGE_FILTERED).copy())
        stages["demod"].append(dec.debug_stage(0, api.STAGE_DEMOD).copy())
    out = {k: np.concatenate(v) for k, v in stages.items()}
    out["bits"] = dec.debug_stage(0, api.STAGE_BITS)
    out["pending"] = dec.debug_stage(0, api.STAGE_PENDING)
    out["taps"] = dec.debug_stage(0, api.STAGE_LPTAPS)
    return dec, out


CASES = [
    # fs, baud, bits, stops, factor, snr_db(full band), n_sentences, chunk
    (2.048e6, 300.0, 8, 2.0, 256, None, 2, 65536),
    (2.048e6, 300.0, 8, 2.0, 256, -15.0, 2, 65536),
    (2.048e6, 300.0, 8, 2.0, 256, -24.0, 2, 65536),
    (2.048e6, 300.0, 8, 2.0, 256, -15.0, 2, 262144),
    (2.048e6, 300.0, 8, 2.0, 256, -15.0, 1, 40000),      # chunk not a multiple of the factor
    (2.5e6, 50.0, 7, 2.0, 256, -18.0, 1, 65536),
    (2.048e6, 100.0, 8, 1.0, 256, -15.0, 1, 65536),
    (1.024e6, 300.0, 8, 2.0, 128, -14.0, 1, 65536),
    (0.512e6, 300.0, 8, 2.0, 64, -12.0, 1, 65536),
    (0.256e6, 300.0, 8, 2.0, 32, -10.0, 1, 65536),
    (0.128e6, 300.0, 8, 2.0, 16, -8.0, 1, 65536),
    (64e3, 300.0, 8, 2.0, 8, -6.0, 1, 65536),
    (32e3, 300.0, 8, 2.0, 4, -4.0, 1, 65536),
    (16e3, 300.0, 8, 2.0, 2, -3.0, 1, 65536),
]


@pytest.mark.parametrize("fs,baud,bits,stops,factor,snr,nsent,chunk", CASES)
def test_single_channel_stages_and_chars(oracle_kind, fs, baud, bits, stops, factor, snr, nsent, chunk):
    iq, text = synth.channel_iq(3, nsent, fs, baud, bits, int(stops), snr_db=snr)
    cfg = dict(baud=baud, rtty_bits=bits, rtty_stops=stops, dec_factor=factor)
    ref = make_oracle(oracle_kind, **cfg).run(iq, fs, chunk)
    dec, got = run_gpu_single(iq, fs, chunk, **cfg)

    # low-pass taps are designed on the host with the reference's exact arithmetic: bit-exact
    assert np.array_equal(got["taps"].view(np.uint32), ref.stage(po.STAGE_LPTAPS).view(np.uint32))
    for name, st in (("dec", po.STAGE_DECIMATED), ("filt", po.STAGE_FILTERED), ("demod", po.STAGE_DEMOD)):
        want = ref.stage(st)
        assert got[name].shape == want.shape, name
        err = rel_l2(got[name], want)
        assert err <= REL_L2, "%s rel L2 %.3g" % (name, err)
    assert dec.poll_chars(0) == ref.chars()
    assert dec.poll_sentences(0) == ref.sentences()
    assert dec.getLastSentence(0) == ref.last_sentence()
    assert dec.getRTTY(0) == ref.rtty()
    if snr is None or snr > -20:
        assert len(ref.sentences()) == nsent          # the workload really decodes
    # slicer residue: same number of pending samples, same values to float tolerance
    want_p = ref.stage(po.STAGE_PENDING)
    assert got["pending"].shape == want_p.shape
    if len(want_p):
        assert rel_l2(got["pending"], want_p) <= 1e-4


def test_bits_match_port_oracle():
    """Every bit the slicer emits equals the restatement's (the reference does not expose its bits)."""
    fs, baud = 2.048e6, 300.0
    iq, _ = synth.channel_iq(5, 2, fs, baud, snr_db=-22.0)
    cfg = dict(baud=baud, rtty_bits=8, rtty_stops=2.0, dec_factor=256)
    port = po.PortDecoder(po.make_config(**cfg)).run(iq, fs)
    dec, got = run_gpu_single(iq, fs, 65536, **cfg)
    assert np.array_equal(got["bits"], port.stage(po.STAGE_BITS))
    assert dec.poll_raw_chars(0) == bytes(port.stage(po.STAGE_RAWCHARS).astype(np.uint8))


def test_batch_mixed_baud_snr_sweep(oracle_kind):
    """cfg 3 in miniature: channels with 50/100/300 baud and an SNR sweep, decoded as one batch."""
    fs = 2.048e6
    n_ch = 12
    bauds = [50.0, 100.0, 300.0]
    n = 65536 * 40
    iqs, cfgs = [], []
    for c in range(n_ch):
        baud = bauds[c % 3]
        snr = -28.0 + 20.0 * c / (n_ch - 1)       # in-band 0 .. 20 dB
        iq, _ = synth.channel_iq(c, 1, fs, baud, snr_db=snr, n_samples=n)
        iqs.append(iq)
        cfgs.append(dict(baud=baud, rtty_bits=8, rtty_stops=2.0, dec_factor=256))
    iq = np.stack(iqs)
    dec = api.BatchDecoder(n_ch, dec_factor=256)
    for c in range(n_ch):
        dec.baud(cfgs[c]["baud"], c)
    for o in range(0, n, 65536):
        dec.pushSamplesBatch(np.ascontiguousarray(iq[:, o:o + 65536]), fs)
        dec.process()
    for c in range(n_ch):
        ref = make_oracle(oracle_kind, **cfgs[c]).run(iq[c], fs)
        assert dec.poll_chars(c) == ref.chars(), "channel %d" % c
        assert dec.poll_sentences(c) == ref.sentences(), "channel %d" % c


def test_async_steps_equal_sync_steps():
    """process_async x N + collect gives the same characters as N synchronous process() calls."""
    fs, baud = 2.048e6, 300.0
    iq, _ = synth.channel_iq(1, 1, fs, baud, snr_db=-15.0)
    n = len(iq) // 65536 * 65536
    a = api.BatchDecoder(1, baud=baud)
    b = api.BatchDecoder(1, baud=baud)
    for o in range(0, n, 65536):
        a.pushSamples(0, iq[o:o + 65536], fs); a.process()
        b.pushSamples(0, iq[o:o + 65536], fs); b.process_async()
    b.collect()
    assert a.poll_chars(0) == b.poll_chars(0)
    assert a.poll_sentences(0) == b.poll_sentences(0)
    assert len(a.getLastSentence(0)) > 10


@pytest.mark.parametrize("fs,baud,bits,n_sent,nfft", [
    (2.048e6, 300.0, 8, 3, 4096),        # the reference's fft_bins_cnt_
    (2.5e6, 50.0, 7, 1, 16384),          # BASELINE configs[1]: AirSpy rate, 50 baud 7N2, 16k-bin spectrum
    (2.048e6, 300.0, 8, 6, 16384),
])
def test_fft_and_afc(oracle_kind, fs, baud, bits, n_sent, nfft):
    """Spectrum vs a float64 DFT of the same decimated frame; AFC scalars vs the oracle (for 16384 bins the oracle is the
    reference Decoder with its fft_bins_cnt_ member set to 16384 at run time, FFT/AFC classes unchanged)."""
    iq, _ = synth.channel_iq(2, n_sent, fs, baud, bits, 2, snr_db=-15.0, f_off=60.0)
    cfg = dict(baud=baud, rtty_bits=bits, rtty_stops=2.0, dec_factor=256)
    ref = make_oracle(oracle_kind, fft_bins=nfft, **cfg).run(iq, fs)
    dec = api.BatchDecoder(1, record=True, fft_bins=nfft, **cfg)
    assert dec.getBinsCount() == nfft
    frames, cur = [], []
    for o in range(0, len(iq), 65536):
        dec.pushSamples(0, iq[o:o + 65536], fs)
        dec.process()
        d = dec.debug_stage(0, api.STAGE_DECIMATED)
        cur.extend(d.tolist())
        if len(cur) >= nfft:
            frames.append(np.asarray(cur[:nfft], dtype=np.complex64)); cur = []
    assert len(frames) >= 2
    spec = dec.getFFT(0)
    assert spec.shape == (nfft,)
    want = np.fft.fftshift(np.fft.fft(frames[-1].astype(np.complex128)))
    assert rel_l2(spec, want) <= REL_L2
    assert rel_l2(spec, ref.stage(po.STAGE_FFT)) <= REL_L2
    pw, pw_ref = dec.getPowerSpectrum(0), ref.stage(po.STAGE_POWER)
    assert np.max(np.abs(pw - pw_ref)) < 2e-2      # dB; bins near a spectral null amplify float noise
    a = ref.afc()
    assert dec.getPeaks(0) == (a.peak_left, a.peak_right)
    nf, nv = dec.getNoiseFloor(0)
    assert abs(nf - a.noise_floor) < 1e-3 and abs(nv - a.noise_variance) < 1e-3
    assert dec.getShift(0) == pytest.approx(a.shift_hz, abs=1e-9)
    assert dec.getFrequencyCorrection(0) == pytest.approx(a.frequency_correction, abs=1e-9)
    info, power = dec.getSpectrumInfo(0)
    assert info.peak_left_ == abs(a.peak_left) and info.peak_right_ == abs(a.peak_right)
    assert info.sampling_rate_ == fs / 256 and len(power) == nfft
    assert dec.poll_chars(0) == ref.chars()


# ---- golden fixtures produced by the unmodified reference (tests/golden/make_golden.py) ---------------------
import glob
import os

_GOLD = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "g[0-9]*.npz")))


@pytest.mark.parametrize("path", _GOLD, ids=[os.path.basename(p)[:-4] for p in _GOLD])
def test_gpu_matches_golden_reference_vectors(path):
    g = np.load(path)
    q = g["iq_q"]
    iq = (q[:, 0].astype(np.float32) / np.float32(32.0) + 1j * (q[:, 1].astype(np.float32) / np.float32(32.0))).astype(np.complex64)
    cfg = dict(baud=float(g["baud"]), rtty_bits=int(g["bits"]), rtty_stops=float(g["stops"]), dec_factor=int(g["factor"]),
               dc_remove=bool(g["dc_remove"]))
    if "chunks" in g.files:      # irregular call pattern + run-time setters (tests/golden/make_golden.py: g4)
        from test_oracle import gold_calls
        dec = api.BatchDecoder(1, record=True, **cfg)
        stages = {"dec": [], "filt": [], "demod": []}
        o = 0
        for n, events in gold_calls(g, len(iq)):
            for name, v in events:
                getattr(dec, name)(v, 0)
            dec.pushSamples(0, iq[o:o + n], float(g["fs"]))
            dec.process()
            o += n
            stages["dec"].append(dec.debug_stage(0, api.STAGE_DECIMATED).copy())
            stages["filt"].append(dec.debug_stage(0, api.STAGE_FILTERED).copy())
            stages["demod"].append(dec.debug_stage(0, api.STAGE_DEMOD).copy())
        got = {k: np.concatenate(v) for k, v in stages.items()}
        got["pending"] = dec.debug_stage(0, api.STAGE_PENDING)
        got["taps"] = dec.debug_stage(0, api.STAGE_LPTAPS)
    else:
        dec, got = run_gpu_single(iq, float(g["fs"]), int(g["chunk"]), **cfg)
    assert np.array_equal(got["taps"].view(np.uint32), g["lptaps"].view(np.uint32))
    for name in ("decimated", "filtered", "demod"):
        a = got[{"decimated": "dec", "filtered": "filt", "demod": "demod"}[name]]
        assert a.shape == g[name].shape, name
        assert rel_l2(a, g[name]) <= REL_L2, name
    assert dec.poll_chars(0) == g["chars"].tobytes()
    assert dec.getRTTY(0) == g["rtty"].tobytes()
    assert dec.getLastSentence(0) == g["last_sentence"].tobytes()
    assert b"\n".join(dec.poll_sentences(0)) == g["sentences"].tobytes()
    assert got["pending"].shape == g["pending"].shape
    if len(g["power"]):
        afc = g["afc"]
        assert dec.getPeaks(0) == (int(afc[4]), int(afc[5]))
        assert dec.getFrequencyCorrection(0) == pytest.approx(afc[0], abs=1e-9)
        assert dec.getShift(0) == pytest.approx(afc[1], abs=1e-9)


def test_cfg3_256_channels_mixed_baud_snr_sweep(oracle_kind):
    """BASELINE configs[2] at full width: 256 channels on one GPU, 50/100/300 baud round robin, in-band SNR swept
    0..20 dB (full band -28..-8 dB); every channel's characters -- including the wrong ones at low SNR -- equal the
    reference's.  The oracle side runs one reference Decoder per channel."""
    fs = 2.048e6
    n_ch = 256
    bauds = [50.0, 100.0, 300.0]
    n = 65536 * 24
    iq = np.empty((n_ch, n), dtype=np.complex64)
    for c in range(n_ch):
        snr = -28.0 + 20.0 * c / (n_ch - 1)
        iq[c] = synth.channel_iq(c, 1, fs, bauds[c % 3], snr_db=snr, n_samples=n, lead_in=12)[0]
    dec = api.BatchDecoder(n_ch, dec_factor=256)
    for c in range(n_ch):
        dec.baud(bauds[c % 3], c)
    for o in range(0, n, 65536):
        dec.pushSamplesBatch(np.ascontiguousarray(iq[:, o:o + 65536]), fs)
        dec.process_async()
    dec.collect()
    n_chars = 0
    for c in range(n_ch):
        ref = make_oracle(oracle_kind, baud=bauds[c % 3], rtty_bits=8, rtty_stops=2.0, dec_factor=256).run(iq[c], fs)
        got = dec.poll_chars(c)
        assert got == ref.chars(), "channel %d" % c
        assert dec.poll_sentences(c) == ref.sentences(), "channel %d" % c
        n_chars += len(got)
    assert n_chars > 2000
