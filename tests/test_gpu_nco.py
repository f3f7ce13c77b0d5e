"""GPU parity of the NCO pre-mixer, the wideband channeliser front-end (BASELINE configs[4]) and the closed AFC
loop (BASELINE configs[1]).  Oracle = CPU pre-mix (oracle/pyoracle.premix: float64 phase, cf32 product) followed by
the reference Decoder, with the retune decisions of code/websocketServer/main.cpp:247-265 applied on both sides.
"""
import numpy as np
import pytest

from habdec_b200 import api, synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

REL_L2 = 1e-5
CHUNK = 65536


def rel_l2(a, b):
    a = np.asarray(a).astype(np.complex128 if np.iscomplexobj(a) else np.float64)
    b = np.asarray(b).astype(a.dtype)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def make_oracle(kind, **cfg):
    return (po.RefDecoder if kind == "ref" else po.PortDecoder)(po.make_config(**cfg))


def short_sentence(c, k=0):
    # > 20 characters: the reference only scans for sentences beyond that (Decoder.h:591)
    body = "WIDE%02d,%d,12:00:%02d,%d" % (c, k, c, 700 * c + k)
    return "$$" + body + "*" + synth.crc16_ccitt(body.encode()) + "\n"


def test_offset_channel_per_channel_push(oracle_kind):
    """One channel 37 kHz off centre: NCO on the GPU vs premix + reference; stage floats and characters."""
    fs, baud, f_c = 2.048e6, 300.0, 37000.0
    iq, _ = synth.channel_iq(4, 1, fs, baud, f_off=f_c, snr_db=-14.0)
    cfg = dict(baud=baud, rtty_bits=8, rtty_stops=2.0, dec_factor=256)
    ref = make_oracle(oracle_kind, **cfg)
    dec = api.BatchDecoder(1, record=True, **cfg)
    dec.set_nco(f_c, 0)
    assert dec.get_nco(0) == f_c
    ph = 0.0
    got = {"dec": [], "demod": []}
    for o in range(0, len(iq), CHUNK):
        mixed, ph = po.premix(iq[o:o + CHUNK], fs, f_c, ph)
        ref.push_process(mixed, fs)
        dec.pushSamples(0, iq[o:o + CHUNK], fs)
        dec.process()
        got["dec"].append(dec.debug_stage(0, api.STAGE_DECIMATED).copy())
        got["demod"].append(dec.debug_stage(0, api.STAGE_DEMOD).copy())
    assert rel_l2(np.concatenate(got["dec"]), ref.stage(po.STAGE_DECIMATED)) <= REL_L2
    assert rel_l2(np.concatenate(got["demod"]), ref.stage(po.STAGE_DEMOD)) <= REL_L2
    assert dec.poll_chars(0) == ref.chars()
    assert dec.poll_sentences(0) == ref.sentences() and len(ref.sentences()) == 1


def _wideband(fs, offsets, bauds, n_sent, snr_db, seed=7):
    texts = ["".join(short_sentence(c, k) for k in range(n_sent)) for c in range(len(offsets))]
    bit_streams = [synth.uart_bits(t.encode(), 8, 2, 30, 40) for t in texts]
    n = max(int(len(b) * fs / bd) for b, bd in zip(bit_streams, bauds))
    wide = np.zeros(n, dtype=np.complex64)
    for c, (b, bd, f) in enumerate(zip(bit_streams, bauds, offsets)):
        wide += synth.fsk_iq(b, fs, bd, 425.0, f, None, n_samples=n, phase0=0.3 * c)
    rng = np.random.default_rng(seed)
    sigma = 10.0 ** (-snr_db / 20.0) / np.sqrt(2.0)
    wide.real += (sigma * rng.standard_normal(n)).astype(np.float32)
    wide.imag += (sigma * rng.standard_normal(n)).astype(np.float32)
    return wide, texts


@pytest.mark.parametrize("fs,offsets,bauds", [
    (2.5e6, [-900e3, -310e3, -40e3, 55e3, 420e3, 1.1e6], [600.0, 300.0, 600.0, 300.0, 600.0, 600.0]),
    (20e6, [-7.3e6, -1.25e6, 0.6e6, 8.8e6], [600.0, 600.0, 600.0, 600.0]),   # cfg 5: 20 MS/s capture, 600 baud bursts, dec=8
])
def test_wideband_channeliser(oracle_kind, fs, offsets, bauds):
    """One capture, many frequency-offset channels: hbd_push_wideband + per-channel NCO vs premix + reference."""
    wide, texts = _wideband(fs, offsets, bauds, 1, snr_db=-8.0)
    n = len(wide) // CHUNK * CHUNK
    n_ch = len(offsets)
    dec = api.BatchDecoder(n_ch, dec_factor=256)
    for c in range(n_ch):
        dec.baud(bauds[c], c)
        dec.set_nco(offsets[c], c)
    for o in range(0, n, CHUNK):
        dec.pushWideband(wide[o:o + CHUNK], fs)
        dec.process()
    decoded = 0
    for c in range(n_ch):
        ref = make_oracle(oracle_kind, baud=bauds[c], rtty_bits=8, rtty_stops=2.0, dec_factor=256)
        ph = 0.0
        for o in range(0, n, CHUNK):
            mixed, ph = po.premix(wide[o:o + CHUNK], fs, offsets[c], ph)
            ref.push_process(mixed, fs)
        assert dec.poll_chars(c) == ref.chars(), "channel %d" % c
        assert dec.poll_sentences(c) == ref.sentences(), "channel %d" % c
        decoded += len(ref.sentences())
    assert decoded == n_ch      # every channel's sentence came through the channeliser


def test_device_push_with_nco_matches_host_push():
    """hbd_push_samples_device with an active NCO goes through the staging matrix and equals the host path."""
    import torch
    fs, baud = 2.048e6, 300.0
    n_ch = 3
    offs = [0.0, 12500.0, -48000.0]
    iqs = [synth.channel_iq(c, 1, fs, baud, f_off=offs[c], snr_db=-12.0)[0] for c in range(n_ch)]
    n = min(len(x) for x in iqs) // CHUNK * CHUNK
    iq = np.stack([x[:n] for x in iqs])
    a = api.BatchDecoder(n_ch, baud=baud)
    b = api.BatchDecoder(n_ch, baud=baud)
    for c in range(n_ch):
        a.set_nco(offs[c], c); b.set_nco(offs[c], c)
    dev = torch.from_numpy(iq.view(np.float32).reshape(n_ch, n, 2)).cuda()
    torch.cuda.synchronize()
    for o in range(0, n, CHUNK):
        a.pushSamplesBatch(np.ascontiguousarray(iq[:, o:o + CHUNK]), fs); a.process()
        b.pushSamplesDevice(dev.data_ptr() + o * 8, CHUNK, n, fs); b.process()
    for c in range(n_ch):
        ca = a.poll_chars(c)
        assert ca == b.poll_chars(c) and len(ca) > 30
        assert len(a.poll_sentences(c)) == 1


def _retune_loop(push_process, get_corr, retune, fs, n_total, period_s):
    """DECODER_THREAD's AFC policy (main.cpp:247-265) on stream time: once `period_s` has passed since the last
    retune, every call checks the correction and retunes when it exceeds 100 Hz."""
    last = 0.0
    log = []
    for o in range(0, n_total, CHUNK):
        push_process(o)
        t = (o + CHUNK) / fs
        if t - last > period_s:
            corr = get_corr()
            if 100 < abs(corr):
                retune(corr)
                log.append((o, corr))
                last = t
    return log


def test_afc_closed_loop_drifting_carrier(oracle_kind):
    """cfg 2: 2.5 MS/s, 50 baud 7N2, carrier 1.5 kHz off and drifting; the AFC measures, the NCO follows.
    Without the loop nothing decodes (tones must straddle DC)."""
    fs, baud = 2.5e6, 50.0
    sentence = short_sentence(2, 5)
    text = "U" * 15 + "\n" + sentence          # a preamble with both tones on air lets the two-peak AFC lock
    bits = synth.uart_bits(text.encode(), 7, 2, lead_in=40, lead_out=40)
    n = int(len(bits) * fs / baud) // CHUNK * CHUNK
    t = np.arange(n, dtype=np.float64) / fs
    f_off = 1500.0 + 15.0 * t                      # Hz, drifting upwards
    iq = synth.fsk_iq(bits, fs, baud, 425.0, f_off, snr_db=-16.0, seed=5, n_samples=n)
    cfg = dict(baud=baud, rtty_bits=7, rtty_stops=2.0, dec_factor=256)
    period = 1.5                                   # seconds of stream time (the reference waits 5 s of wall clock)

    ref = make_oracle(oracle_kind, **cfg)
    st = {"f": 0.0, "ph": 0.0}

    def ref_push(o):
        mixed, st["ph"] = po.premix(iq[o:o + CHUNK], fs, st["f"], st["ph"])
        ref.push_process(mixed, fs)

    def ref_retune(corr):
        st["f"] += corr
        ref.reset_frequency_correction(corr)

    log_ref = _retune_loop(ref_push, lambda: ref.afc().frequency_correction, ref_retune, fs, n, period)

    dec = api.BatchDecoder(1, **cfg)

    def gpu_push(o):
        dec.pushSamples(0, iq[o:o + CHUNK], fs)
        dec.process()

    def gpu_retune(corr):
        dec.set_nco(dec.get_nco(0) + corr, 0)
        dec.resetFrequencyCorrection(corr, 0)

    log_gpu = _retune_loop(gpu_push, lambda: dec.getFrequencyCorrection(0), gpu_retune, fs, n, period)

    assert len(log_ref) >= 2 and abs(log_ref[0][1] - 1520.0) < 100.0        # the AFC found the offset, then followed the drift
    assert [o for o, _ in log_gpu] == [o for o, _ in log_ref]               # same retune instants
    assert np.allclose([c for _, c in log_gpu], [c for _, c in log_ref], atol=1e-6)
    assert dec.poll_chars(0) == ref.chars()
    assert dec.poll_sentences(0) == ref.sentences()
    assert ref.sentences() == [sentence.strip()[2:].encode()]               # the loop made the channel decodable


def test_afc_retune_all_channels_on_gpu():
    """hbd_afc_retune == per-channel get / set_nco / reset sequence."""
    fs, baud = 2.048e6, 300.0
    offs = [0.0, 700.0, -900.0, 30.0]
    n_ch = len(offs)
    iqs = [synth.channel_iq(c, 3, fs, baud, f_off=offs[c], snr_db=-12.0)[0] for c in range(n_ch)]
    n = min(len(x) for x in iqs) // CHUNK * CHUNK
    iq = np.stack([x[:n] for x in iqs])
    a = api.BatchDecoder(n_ch, baud=baud)
    b = api.BatchDecoder(n_ch, baud=baud)
    half = (n // CHUNK // 2) * CHUNK
    for o in range(0, half, CHUNK):
        blk = np.ascontiguousarray(iq[:, o:o + CHUNK])
        a.pushSamplesBatch(blk, fs); a.process()
        b.pushSamplesBatch(blk, fs); b.process()
    corr = [a.getFrequencyCorrection(c) for c in range(n_ch)]
    for c in range(n_ch):
        if 100 < abs(corr[c]):
            a.set_nco(a.get_nco(c) + corr[c], c)
            a.resetFrequencyCorrection(corr[c], c)
    applied = b.afc_retune(100.0)
    assert np.allclose(applied, [x if 100 < abs(x) else 0.0 for x in corr], atol=1e-9)
    assert abs(applied[1] - 700.0) < 100.0 and abs(applied[2] + 900.0) < 100.0 and applied[0] == 0.0 and applied[3] == 0.0
    for o in range(half, n, CHUNK):
        blk = np.ascontiguousarray(iq[:, o:o + CHUNK])
        a.pushSamplesBatch(blk, fs); a.process()
        b.pushSamplesBatch(blk, fs); b.process()
    for c in range(n_ch):
        assert a.poll_chars(c) == b.poll_chars(c)
        assert a.getPeaks(c) == b.getPeaks(c)
        assert a.get_nco(c) == b.get_nco(c)


@pytest.mark.parametrize("chunk", [65536, 40000, 40001, 262144])
def test_fused_nco_equals_premix_kernel(oracle_kind, chunk):
    """The NCO fused into K1 (device / wideband pushes: no K0 launch, no staging matrix) against the K0 + staging path
    (HBD_NCO_FUSED=0) and against premix + reference: chunks that are not multiples of the factor (mixed remainder in
    the carry), 262 144-sample chunks (more than 32 blocks of 4096: the E table is refilled) and a retune between calls."""
    import os
    fs = 2.5e6
    offsets, bauds = [-310e3, 55e3, 0.0, 420e3], [300.0, 300.0, 300.0, 600.0]
    wide, _ = _wideband(fs, offsets, bauds, 1, snr_db=-8.0)
    n = len(wide) // chunk * chunk
    n_ch = len(offsets)
    decs = []
    for fused in ("1", "0"):
        os.environ["HBD_NCO_FUSED"] = fused
        try:
            d = api.BatchDecoder(n_ch, dec_factor=256, record=True)
        finally:
            del os.environ["HBD_NCO_FUSED"]
        for c in range(n_ch):
            d.baud(bauds[c], c)
            d.set_nco(offsets[c], c)
        decs.append(d)
    refs = [make_oracle(oracle_kind, baud=bauds[c], rtty_bits=8, rtty_stops=2.0, dec_factor=256) for c in range(n_ch)]
    phases = [0.0] * n_ch
    freqs = list(offsets)
    stages = [[[] for _ in range(n_ch)] for _ in decs]
    launches0 = [d.kernel_launches() for d in decs]
    for i, o in enumerate(range(0, n, chunk)):
        if i == 2:      # retune channel 1 by +37 Hz between two calls (what hbd_afc_retune does)
            freqs[1] += 37.0
            for d in decs:
                d.set_nco(freqs[1], 1)
        blk = wide[o:o + chunk]
        for k, d in enumerate(decs):
            d.pushWideband(blk, fs)
            d.process()
            for c in range(n_ch):
                stages[k][c].append(d.debug_stage(c, api.STAGE_DECIMATED).copy())
        for c in range(n_ch):
            mixed, phases[c] = po.premix(blk, fs, freqs[c], phases[c])
            refs[c].push_process(mixed, fs)
    # the fused decoder never launched K0: one kernel less per call
    calls = n // chunk
    assert (decs[1].kernel_launches() - launches0[1]) - (decs[0].kernel_launches() - launches0[0]) >= calls
    for c in range(n_ch):
        fused = np.concatenate(stages[0][c]); k0 = np.concatenate(stages[1][c]); want = refs[c].stage(po.STAGE_DECIMATED)
        assert fused.shape == k0.shape == want.shape
        den = max(np.linalg.norm(want.astype(np.complex128)), 1e-30)
        assert np.linalg.norm(fused.astype(np.complex128) - k0) / den <= 1e-6, "channel %d fused vs K0" % c
        assert np.linalg.norm(fused.astype(np.complex128) - want) / den <= 1e-5, "channel %d fused vs reference" % c
        assert decs[0].poll_chars(c) == refs[c].chars(), "channel %d" % c
        assert decs[0].poll_sentences(c) == refs[c].sentences(), "channel %d" % c


def _torch_fsk_into(cap, bits, fs, baud, f_c, phase0, torch):
    """cap += exp(j phi), continuous-phase 2-FSK of `bits` at offset f_c (float64 phase on the GPU, slices of 8 M samples)."""
    n = cap.shape[0]
    dev = cap.device
    bits_t = torch.from_numpy(np.asarray(bits, dtype=np.int8)).to(dev)
    ph = float(phase0)
    step = 1 << 23
    for o in range(0, n, step):
        m = min(step, n - o)
        idx = torch.clamp((torch.arange(o, o + m, dtype=torch.float64, device=dev) * (baud / fs)).to(torch.int64), max=len(bits) - 1)
        f = torch.where(bits_t[idx] > 0, 0.5 * 425.0, -0.5 * 425.0).to(torch.float64) + f_c
        phi = ph + torch.cumsum(2.0 * np.pi * f / fs, dim=0)
        ph = float(phi[-1].item()) % (2.0 * np.pi)
        cap[o:o + m, 0] += torch.cos(phi).to(torch.float32)
        cap[o:o + m, 1] += torch.sin(phi).to(torch.float32)


def test_cfg5_full_width_1024_nco_channels_with_ssdv_bursts(oracle_kind):
    """BASELINE configs[4] at full width: ONE 20 MS/s capture (6.7 s) channelised into 1024 frequency-offset channels on a
    15 kHz raster through the NCO fused into K1, dec=8 (78 125 S/s per channel), SSDV packet sync on.  64 of the channels
    carry a signal: 56 RTTY sentences at 300 baud, 8 bursts at 600 baud with an SSDV packet between junk.  Every one of the
    64 is compared with pre-mix + reference Decoder on the same capture: characters, sentences and SSDV events exact."""
    import concurrent.futures
    import torch
    import ssdv_cases
    fs, n_ch, raster = 20e6, 1024, 15e3
    n = 2048 * CHUNK                                       # 6.7 s
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev); gen.manual_seed(2024)
    cap = 0.7 * torch.randn((n, 2), dtype=torch.float32, device=dev, generator=gen)
    active = [8 + 16 * k for k in range(64)]
    bauds, payloads = {}, {}
    rng = np.random.default_rng(11)
    for k, c in enumerate(active):
        if k % 8 == 3:                                     # an SSDV burst: junk, one packet, a sentence, junk
            bauds[c] = 600.0
            payloads[c] = ssdv_cases.junk(rng, 20, 0.1) + ssdv_cases.random_packet(rng, ssdv_cases.CALLSIGNS[k % 4], 1 + k // 8, 0, fec=True) + \
                short_sentence(k).encode() + ssdv_cases.junk(rng, 12, 0.0)
        else:
            bauds[c] = 300.0
            payloads[c] = "".join(short_sentence(k, j) for j in range(4)).encode()
        bits = synth.uart_bits(payloads[c], 8, 2, lead_in=30, lead_out=40)
        assert len(bits) * fs / bauds[c] < n
        _torch_fsk_into(cap, bits, fs, bauds[c], (c - n_ch / 2) * raster, 0.3 * k, torch)
    torch.cuda.synchronize()

    dec = api.BatchDecoder(n_ch, baud=300.0, rtty_bits=8, rtty_stops=2.0, dec_factor=256)
    dec.set_stream(torch.cuda.current_stream().cuda_stream)
    for c in range(n_ch):
        dec.set_nco((c - n_ch / 2) * raster, c)
    for c in active:
        dec.baud(bauds[c], c)
    dec.set_ssdv(True)
    events = {c: [] for c in active}
    for i in range(n // CHUNK):
        dec.pushWidebandDevice(cap.data_ptr() + i * CHUNK * 8, CHUNK, fs)
        dec.process_async()
        if (i + 1) % 8 == 0:
            dec.collect_ready(4)
    dec.collect()
    for c in active:
        events[c] = [(cs, iid, pid, w, h, size) for (cs, iid, pid, w, h, err, size, pkt) in dec.poll_ssdv_packets(c)]
    got = {c: (dec.poll_chars(c), dec.poll_sentences(c)) for c in active}
    noise_chars = sum(len(dec.poll_chars(c)) for c in range(0, n_ch, 16))     # channels without a signal decode noise: something, not nothing
    host = cap.cpu().numpy().view(np.complex64).reshape(-1)
    del cap
    torch.cuda.empty_cache()

    def oracle_for(c):
        ref = make_oracle(oracle_kind, baud=bauds[c], rtty_bits=8, rtty_stops=2.0, dec_factor=256)
        ph, f_c = 0.0, (c - n_ch / 2) * raster
        for o in range(0, n, 16 * CHUNK):
            mixed, ph = po.premix(host[o:o + 16 * CHUNK], fs, f_c, ph)
            for q in range(0, len(mixed), CHUNK):
                ref.push_process(mixed[q:q + CHUNK], fs)
        ev = [(e[1], e[2], e[3], e[4], e[5], e[6]) for e in ref.ssdv_events()]
        return c, ref.chars(), ref.sentences(), ev

    bad = []
    n_sent = n_pkt = 0
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as pool:
        for c, chars, sents, ev in pool.map(oracle_for, active):
            if got[c] != (chars, sents) or events[c] != ev:
                bad.append(c)
            n_sent += len(sents); n_pkt += len(ev)
    assert not bad, "channels %r differ from pre-mix + %s" % (bad, oracle_kind)
    assert n_sent >= 56 * 3 and n_pkt >= 6            # the signals really decode (sentences on the RTTY channels, packets on the bursts)
    assert noise_chars > 0
