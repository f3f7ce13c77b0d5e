"""GPU parity, edge cases and full size: the situations the reference's Decoder handles in its own peculiar way
 and BASELINE configs[3] at its full width (4096 channels on one GPU).

Everything goes through the C ABI (habdec_b200.api) and is compared with the compiled reference Decoder
(oracle/_ref) or, where that is absent, with the restatement.
"""
import numpy as np
import pytest

from habdec_b200 import api, synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

REL_L2 = 1e-5   # north star: per-stage floats within relative L2 1e-5 of the reference's float path


def rel_l2(a, b):
    a = np.asarray(a).astype(np.complex128 if np.iscomplexobj(a) else np.float64)
    b = np.asarray(b).astype(a.dtype)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def make_oracle(kind, **cfg):
    return (po.RefDecoder if kind == "ref" else po.PortDecoder)(po.make_config(**cfg))


def raw_chars(port) -> bytes:
    return bytes(port.stage(po.STAGE_RAWCHARS).astype(np.uint8))


def run_pair(kind, iq, fs, chunks, **cfg):
    """The same chunk sequence through the GPU decoder (stage recording on) and the oracle."""
    dec = api.BatchDecoder(1, record=True, **cfg)
    ref = make_oracle(kind, **cfg)
    got = {"dec": [], "filt": [], "demod": []}
    o = 0
    for n in chunks:
        blk = iq[o:o + n]
        o += n
        dec.pushSamples(0, blk, fs)
        dec.process()
        ref.push_process(blk, fs)
        got["dec"].append(dec.debug_stage(0, api.STAGE_DECIMATED).copy())
        got["filt"].append(dec.debug_stage(0, api.STAGE_FILTERED).copy())
        got["demod"].append(dec.debug_stage(0, api.STAGE_DEMOD).copy())
    return dec, ref, {k: np.concatenate(v) for k, v in got.items()}


def check_stages(got, ref):
    for name, which in (("dec", po.STAGE_DECIMATED), ("filt", po.STAGE_FILTERED), ("demod", po.STAGE_DEMOD)):
        want = ref.stage(which)
        assert got[name].shape == want.shape, name
        if want.size:
            assert rel_l2(got[name], want) <= REL_L2, name


def test_dc_removal_on(oracle_kind):
    """Decoder.h:450-459: the DC blocker re-seeds from the first sample of every call (chunk-size dependent)."""
    fs, baud = 2.048e6, 300.0
    iq, _ = synth.channel_iq(3, 2, fs, baud, snr_db=-15.0)
    iq = (iq + np.complex64(0.35 - 0.2j)).astype(np.complex64)     # a DC offset for the blocker to remove
    n = len(iq) // 65536 * 65536
    cfg = dict(baud=baud, rtty_bits=8, rtty_stops=2.0, dec_factor=256, dc_remove=True)
    dec, ref, got = run_pair(oracle_kind, iq[:n], fs, [65536] * (n // 65536), **cfg)
    check_stages(got, ref)
    assert dec.poll_chars(0) == ref.chars()
    assert dec.poll_sentences(0) == ref.sentences()
    assert len(ref.chars()) > 30


def test_varying_chunk_sizes_with_empty_pushes(oracle_kind):
    """Chunks that grow, shrink, are not multiples of the factor, and empty pushes: the remainder queue
    (Decoder.h:429-435), the 256-batch gate (:492-495, :532) and the history re-zeroing when the reference's work
    buffers grow (Decimator.h:74-79, FirFilter.h:141-147) all depend on the exact sequence."""
    fs, baud = 2.048e6, 300.0
    iq, _ = synth.channel_iq(7, 3, fs, baud, snr_db=-14.0)
    sizes, left, k = [], len(iq), 0
    pattern = [65536, 0, 70001, 9500, 100000, 65536, 0, 0, 131072, 12345, 262144, 40000, 65537, 9999]
    while left > 0:
        n = min(pattern[k % len(pattern)], left)
        if 0 < left - n < 9500:      # keep every non-empty chunk above the reference's well-defined minimum
            n = left
        sizes.append(n)
        left -= n
        k += 1
    cfg = dict(baud=baud, rtty_bits=8, rtty_stops=2.0, dec_factor=256)
    dec, ref, got = run_pair(oracle_kind, iq, fs, sizes, **cfg)
    check_stages(got, ref)
    assert dec.poll_chars(0) == ref.chars()
    assert dec.poll_sentences(0) == ref.sentences()
    assert dec.getLastSentence(0) == ref.last_sentence()
    assert len(ref.sentences()) >= 2
    # FFT frames complete on irregular calls here (the host decides per call whether K4 has to run): spectrum and AFC
    a = ref.afc()
    assert rel_l2(dec.getFFT(0), ref.stage(po.STAGE_FFT)) <= REL_L2
    assert dec.getPeaks(0) == (a.peak_left, a.peak_right)
    nf, nv = dec.getNoiseFloor(0)
    assert abs(nf - a.noise_floor) < 1e-3 and abs(nv - a.noise_variance) < 1e-3
    assert dec.getShift(0) == pytest.approx(a.shift_hz, abs=1e-9)
    assert dec.getFrequencyCorrection(0) == pytest.approx(a.frequency_correction, abs=1e-9)


def test_slicer_vent_after_long_carrier(oracle_kind):
    """SymbolExtractor.h:116-120: more than 3e4 pending discriminator samples (a carrier without transitions) are
    dropped in one go before the next append.  fs_dec = 8 kHz, 5 s of noise-free mark carrier, then two sentences."""
    fs, baud = 16e3, 300.0
    iq, _ = synth.channel_iq(11, 2, fs, baud, snr_db=None, lead_in=int(5.0 * baud))
    n = len(iq) // 8192 * 8192
    cfg = dict(baud=baud, rtty_bits=8, rtty_stops=2.0, dec_factor=2)
    dec = api.BatchDecoder(1, record=True, **cfg)
    ref = make_oracle(oracle_kind, **cfg)
    pend_max = 0
    vented = False
    for o in range(0, n, 8192):
        dec.pushSamples(0, iq[o:o + 8192], fs)
        dec.process()
        ref.push_process(iq[o:o + 8192], fs)
        p = dec.debug_stage(0, api.STAGE_PENDING).size
        if pend_max > 30000 and p < 8192:
            vented = True
        pend_max = max(pend_max, p)
    assert vented, "the carrier never filled the slicer queue past 3e4 (max %d): the vent was not exercised" % pend_max
    assert dec.poll_chars(0) == ref.chars()
    assert dec.poll_sentences(0) == ref.sentences()
    got_p, want_p = dec.debug_stage(0, api.STAGE_PENDING), ref.stage(po.STAGE_PENDING)
    assert got_p.shape == want_p.shape
    assert rel_l2(got_p, want_p) <= REL_L2
    assert len(ref.sentences()) >= 1


@pytest.mark.parametrize("bits,stops_tx,stops_rx", [(7, 2, 1.5), (8, 1, 1.0), (7, 1, 1.0), (8, 2, 2.5)])
def test_uart_framings(oracle_kind, bits, stops_tx, stops_rx):
    """RTTY.h:77-137: the stop-bit loop compares a size_t with the float stop count and advances by `i += nstops`
    (size_t += float), so fractional stop counts have their own arithmetic; 7-bit frames; one stop bit."""
    fs, baud = 2.048e6, 300.0
    iq, _ = synth.channel_iq(21, 2, fs, baud, nbits=bits, nstops=stops_tx, snr_db=-16.0)
    n = len(iq) // 65536 * 65536
    cfg = dict(baud=baud, rtty_bits=bits, rtty_stops=stops_rx, dec_factor=256)
    dec = api.BatchDecoder(1, **cfg)
    ref = make_oracle(oracle_kind, **cfg)
    port = make_oracle("orc", **cfg)     # the reference does not expose its unfiltered characters; the restatement does
    for o in range(0, n, 65536):
        dec.pushSamples(0, iq[o:o + 65536], fs)
        dec.process()
        ref.push_process(iq[o:o + 65536], fs)
        port.push_process(iq[o:o + 65536], fs)
    assert dec.poll_raw_chars(0) == raw_chars(port)
    assert dec.poll_chars(0) == ref.chars()
    assert dec.poll_sentences(0) == ref.sentences()
    assert len(ref.chars()) > (20 if stops_rx <= stops_tx else 0)   # more stop bits expected than sent: most frames fail


def test_noise_only_and_silence(oracle_kind):
    """No signal at all: pure noise decodes to the same garbage as the reference; all-zero input (atan2f(0, 0),
    log10f(0) = -inf in the spectrum -> the AFC's NaN/Inf guard, AFC.h:96-100,250-283) produces nothing."""
    fs = 2.048e6
    rng = np.random.default_rng(5)
    n = 65536 * 24
    noise = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    zeros = np.zeros(n, dtype=np.complex64)
    for iq in (noise, zeros):
        cfg = dict(baud=300.0, rtty_bits=8, rtty_stops=2.0, dec_factor=256)
        dec = api.BatchDecoder(1, **cfg)
        ref = make_oracle(oracle_kind, **cfg)
        port = make_oracle("orc", **cfg)
        for o in range(0, n, 65536):
            dec.pushSamples(0, iq[o:o + 65536], fs)
            dec.process()
            ref.push_process(iq[o:o + 65536], fs)
            port.push_process(iq[o:o + 65536], fs)
        assert dec.poll_raw_chars(0) == raw_chars(port)
        assert dec.poll_chars(0) == ref.chars()
        assert dec.poll_sentences(0) == ref.sentences()
        a = ref.afc()
        assert dec.getFrequencyCorrection(0) == a.frequency_correction


def test_cfg4_full_width_4096_channels(oracle_kind):
    """BASELINE configs[3] at full size on one GPU: 4096 channels x 2.048 MS/s, 300 baud 8N2, 65 536-sample chunks,
    zero-copy device pushes, pipelined calls, results taken through the gather path (hbd_gather_results -> sink) while
    calls are in flight.  EVERY channel is compared with the oracle (one reference Decoder per channel on the host
    threads): characters and sentences exact; plus the size-independent properties (each channel decodes its own
    CRC-valid ring sentence, nothing leaks between channels)."""
    import os
    import torch
    fs, baud, chunk, n_ch = 2.048e6, 300.0, 65536, 4096
    L = synth.ring_length(fs, baud)
    slices = L // chunk
    steps = 2 * slices + 7
    dev = torch.device("cuda", 0)
    ring = synth.ring_iq_torch(0, n_ch, dev, fs, baud, snr_db=-15.0)      # [n_ch, L, 2] f32 resident in HBM (54 GB)
    torch.cuda.synchronize()
    dec = api.BatchDecoder(n_ch, baud=baud, rtty_bits=8, rtty_stops=2.0, dec_factor=256)
    dec.set_stream(torch.cuda.current_stream().cuda_stream)
    dec.dist_init(0, 1, None)
    sink = api.ResultSink(n_ch)
    for i in range(steps):
        dec.pushSamplesDevice(ring.data_ptr() + (i % slices) * chunk * 8, chunk, L, fs)
        dec.process_async()
        if (i + 1) % 16 == 0:
            dec.collect_ready(4)
            assert dec.gather_results(sink) == n_ch
    dec.collect()
    while True:                      # records hold 256 characters: a second round only if some channel had more pending
        assert dec.gather_results(sink) == n_ch
        if all(not dec.poll_chars(c) for c in (0, n_ch // 2, n_ch - 1)):
            break
    got_chars = [sink.poll_chars(c) for c in range(n_ch)]
    got_sents = [sink.poll_sentences(c) for c in range(n_ch)]
    for c in range(n_ch):
        want = synth.ring_sentence(c).strip().lstrip("$").encode()       # "Cxxxx,yyy,zzz*CRC"
        assert len(got_sents[c]) == 2 and all(s == want for s in got_sents[c]), "channel %d: %r" % (c, got_sents[c])
        body, crc = want.split(b"*")
        assert synth.crc16_ccitt(body).encode() == crc
    assert sink.totals()["sentences_min"] == 2
    # the oracle on ALL channels, in slabs that fit host memory
    cfg = po.make_config(baud=baud, rtty_bits=8, rtty_stops=2.0, dec_factor=256, record=False)
    threads = os.cpu_count() or 4
    bad = []
    for c0 in range(0, n_ch, 256):
        iq = ring[c0:c0 + 256].cpu().numpy().view(np.complex64).reshape(256, L)
        ref_chars, ref_sents = po.run_ring(oracle_kind, cfg, iq, threads, fs, chunk, 0, steps)
        for k in range(256):
            if got_chars[c0 + k] != ref_chars[k] or got_sents[c0 + k] != ref_sents[k]:
                bad.append(c0 + k)
    assert not bad, "%d of %d channels differ from the %s oracle, first: %r" % (len(bad), n_ch, oracle_kind, bad[:8])
    # AFC scalars travelled with the records
    st = sink.stats(7)
    assert st is not None and abs(st[2] - dec.getNoiseFloor(7)[0]) < 1e-3
    del ring
    torch.cuda.empty_cache()


def test_log_flood_4096_channels_without_collecting(oracle_kind):
    """4096 channels sending back to back at 600 baud decode 3.5 k characters per 32 768-sample call; 640 calls are issued
    with no drain by the caller (2.2 M characters, more than the log holds).  The character log must never be overwritten:
    the library drains by itself when the completed calls' heads say so (hbd_decoder::log_pressure), and every character
    has to come out, in order.  All channels: the right number of sentences; 256 channels: exact characters vs the oracle."""
    import os
    import torch
    fs, baud, chunk, n_ch, steps = 2.048e6, 600.0, 32768, 4096, 640
    dev = torch.device("cuda", 0)
    L = synth.ring_length(fs, baud)
    slices = L // chunk
    assert L % chunk == 0
    ring = synth.ring_iq_torch(0, n_ch, dev, fs, baud, snr_db=-15.0)       # 27 GB
    torch.cuda.synchronize()
    dec = api.BatchDecoder(n_ch, baud=baud, rtty_bits=8, rtty_stops=2.0, dec_factor=256)
    dec.set_stream(torch.cuda.current_stream().cuda_stream)
    for i in range(steps):
        dec.pushSamplesDevice(ring.data_ptr() + (i % slices) * chunk * 8, chunk, L, fs)
        dec.process_async()                      # never collect_ready: only what the library decides to drain
    dec.collect()
    chars = [dec.poll_chars(c) for c in range(n_ch)]
    total = sum(len(x) for x in chars)
    assert total > 2_000_000, "%d characters only: the log was not under pressure" % total
    passes = steps // slices
    for c in range(n_ch):
        sents = dec.poll_sentences(c)
        assert passes - 1 <= len(sents) <= passes, "channel %d: %d sentences" % (c, len(sents))
        assert all(x == synth.ring_sentence(c).strip().lstrip("$").encode() for x in sents)
    cfg = po.make_config(baud=baud, rtty_bits=8, rtty_stops=2.0, dec_factor=256, record=False)
    for c0 in (0, n_ch - 128):
        iq = ring[c0:c0 + 128].cpu().numpy().view(np.complex64).reshape(128, L)
        ref_chars, _ = po.run_ring(oracle_kind, cfg, iq, os.cpu_count() or 4, fs, chunk, 0, steps, pitch=1 << 14)
        for k in range(128):
            assert chars[c0 + k] == ref_chars[k], "channel %d" % (c0 + k)
    del ring
    torch.cuda.empty_cache()


def lp_growth_schedule():
    """(call index -> [(setter, value)]) and chunk sizes: the tap count shrinks, grows back, grows after a LARGER call has
    enlarged the reference's work buffer, and grows past it (which zeroes the history)."""
    marks = {5: [("lowpass_trans", 0.05)],        # 161 -> 81: the oldest 80 history samples stay (FirFilter.h:139-152)
             9: [("lowpass_trans", 0.025)],       # 81 -> 161 inside the old buffer: history = 80 kept + 80 stale inputs of call 8
             13: [("lowpass_trans", 0.1)],        # -> 41
             14: [("lowpass_trans", 0.0125)],     # -> 257 taps (clipped to the 256-sample block, |1): buffer has to grow -> zeroed
             22: [("lowpass_trans", 0.02)]}       # 257 -> 201 after the 512-sample call below: shrink again
    sizes = {18: 131072, 19: 131072}              # two double-size calls (nf = 512): the buffer grows to 512 + 257
    marks[24] = [("lowpass_trans", 0.01)]         # 201 -> 257 (still clipped): fits the enlarged buffer -> stale data, no zeroing
    return marks, sizes


def test_lowpass_grow_reads_the_stale_work_buffer(oracle_kind):
    """FirFilter.h:139-160: the filter's work buffer is [T-1 history | inputs of the call] and is only cleared when it has to
    GROW.  A longer filter set at run time therefore starts from the old history followed by stale inputs of the previous
    call (or, after a shrink, by the not-yet-overwritten rest of the older history).  The queue in HBM mirrors the head of that
    buffer, so every sequence of lowpass_trans changes gives the reference's filtered samples."""
    fs, baud = 2.048e6, 300.0
    iq, _ = synth.channel_iq(43, 4, fs, baud, snr_db=-15.0)
    marks, sizes = lp_growth_schedule()
    cfg = dict(baud=baud, rtty_bits=8, rtty_stops=2.0, dec_factor=256)
    dec = api.BatchDecoder(1, record=True, **cfg)
    ref = make_oracle(oracle_kind, **cfg)
    filt, o, i, taps_seen = [], 0, 0, []
    while o + sizes.get(i, 65536) <= len(iq):
        for name, v in marks.get(i, []):
            getattr(dec, name)(v, 0); ref.set_param(name, v)
        nblk = sizes.get(i, 65536)
        dec.pushSamples(0, iq[o:o + nblk], fs)
        dec.process()
        ref.push_process(iq[o:o + nblk], fs)
        filt.append(dec.debug_stage(0, api.STAGE_FILTERED).copy())
        taps_seen.append(len(dec.debug_stage(0, api.STAGE_LPTAPS)))
        assert np.array_equal(dec.debug_stage(0, api.STAGE_LPTAPS), ref.stage(po.STAGE_LPTAPS)), i
        o += nblk; i += 1
    assert i > 26 and {161, 81, 41, 257, 201} <= set(taps_seen)
    got, want = np.concatenate(filt), ref.stage(po.STAGE_FILTERED)
    assert got.shape == want.shape
    # every call on its own: a wrong history shows in the first T-1 outputs of the call after a change
    k = 0
    for j, f in enumerate(filt):
        if len(f):
            assert rel_l2(f, want[k:k + len(f)]) <= REL_L2, "call %d (taps %d)" % (j, taps_seen[j])
        k += len(f)
    assert dec.poll_chars(0) == ref.chars()
    assert dec.poll_sentences(0) == ref.sentences()


def test_setters_between_calls(oracle_kind):
    """Decoder.h:654-706: baud / rtty_bits / rtty_stops / dc_remove changed between two process() calls act on the
    samples, bits and characters already pending.  Channel 1 of a two-channel batch is re-configured twice while
    channel 0 keeps its settings."""
    fs = 2.048e6
    a, _ = synth.channel_iq(31, 2, fs, 300.0, snr_db=-15.0)
    b, _ = synth.channel_iq(32, 2, fs, 100.0, nbits=7, nstops=1, snr_db=-15.0)
    a = a[:len(a) // 65536 * 65536]
    b = b[:len(b) // 65536 * 65536]
    iq1 = np.concatenate([a, b, a])
    iq0 = np.concatenate([a, a, a])[:len(iq1)]
    if len(iq0) < len(iq1):
        iq0 = np.concatenate([iq0, np.zeros(len(iq1) - len(iq0), dtype=np.complex64)])
    n_calls = len(iq1) // 65536
    marks = {len(a) // 65536: [("baud", 100.0), ("rtty_bits", 7), ("rtty_stops", 1.0), ("dc_remove", 1)],
             (len(a) + len(b)) // 65536: [("baud", 300.0), ("rtty_bits", 8), ("rtty_stops", 2.0), ("dc_remove", 0)]}
    dec = api.BatchDecoder(2, baud=300.0, rtty_bits=8, rtty_stops=2.0, dec_factor=256)
    refs = [make_oracle(oracle_kind, baud=300.0), make_oracle(oracle_kind, baud=300.0)]
    iq = np.stack([iq0, iq1])
    for i in range(n_calls):
        for name, v in marks.get(i, []):
            getattr(dec, name)(v, 1)
            refs[1].set_param(name, v)
        blk = np.ascontiguousarray(iq[:, i * 65536:(i + 1) * 65536])
        dec.pushSamplesBatch(blk, fs)
        dec.process()
        for c in range(2):
            refs[c].push_process(blk[c], fs)
    for c in range(2):
        assert dec.poll_chars(c) == refs[c].chars(), "channel %d" % c
        assert dec.poll_sentences(c) == refs[c].sentences(), "channel %d" % c
    assert len(refs[1].sentences()) >= 5


@pytest.mark.parametrize("first,snr", [((0, 0.0), -15.0), ((7, 1.0), -15.0), ((8, 1.5), -22.0), ((5, 2.0), -24.0)])
def test_uart_backlog_is_rescanned_after_a_framing_change(oracle_kind, first, snr):
    """RTTY::operator() keeps every bit since the last decoded character and rescans ALL of them under the framing in
    force at the next call (RTTY.h:90-134).  Start with the framing unset (bits accumulate, nothing decodes) or wrong
    (garbage decodes, rejected positions stay), then switch to 8N2 mid-stream: the characters the reference digs out of
    the backlog have to come out here as well.  Channel 0 keeps 8N2 throughout; channels 2/3 switch at other calls."""
    fs, baud = 2.048e6, 300.0
    sigs = [synth.channel_iq(60 + c, 2, fs, baud, snr_db=snr)[0] for c in range(4)]
    n = min(len(x) for x in sigs) // 65536 * 65536
    iq = np.stack([x[:n] for x in sigs])
    n_calls = n // 65536
    switch = {1: 7, 2: 11, 3: n_calls // 2}
    dec = api.BatchDecoder(4, baud=baud, rtty_bits=8, rtty_stops=2.0, dec_factor=256)
    refs = [make_oracle(oracle_kind, baud=baud) for _ in range(4)]
    for c in switch:
        dec.rtty_bits(first[0], c); dec.rtty_stops(first[1], c)
        refs[c].set_param("rtty_bits", first[0]); refs[c].set_param("rtty_stops", first[1])
    for i in range(n_calls):
        for c, at in switch.items():
            if i == at:
                dec.rtty_bits(8, c); dec.rtty_stops(2.0, c)
                refs[c].set_param("rtty_bits", 8); refs[c].set_param("rtty_stops", 2.0)
        blk = np.ascontiguousarray(iq[:, i * 65536:(i + 1) * 65536])
        dec.pushSamplesBatch(blk, fs)
        dec.process()
        for c in range(4):
            refs[c].push_process(blk[c], fs)
    for c in range(4):
        assert dec.poll_chars(c) == refs[c].chars(), "channel %d" % c
        assert dec.poll_sentences(c) == refs[c].sentences(), "channel %d" % c
    if first == (0, 0.0):   # everything sent before the switch sat in the backlog: no character is lost
        assert len(refs[3].sentences()) == len(refs[0].sentences()) >= 2


@pytest.mark.parametrize("fs,max_rate,want", [(2.048e6, 8000.0, 256), (2.048e6, 40000.0, 64), (1.024e6, 8000.0, 128),
                                              (256e3, 9000.0, 32), (64e3, 8000.0, 8)])
def test_setup_decimation_stages_bw(oracle_kind, fs, max_rate, want):
    """Decoder::setupDecimationStagesBW (Decoder.h:336-412): 0 before a sampling rate is latched; then the smallest
    power of two (2..128, else 256) that brings the rate under the limit, same stage plan as the factor call.
    (Limits that need more than one plan: test_setup_decimation_stages_bw_cascade.)"""
    baud = 300.0
    iq, _ = synth.channel_iq(4, 1, fs, baud, snr_db=-10.0 if fs > 1e5 else -3.0)
    chunk = 65536
    n = len(iq) // chunk * chunk
    dec = api.BatchDecoder(1, baud=baud, rtty_bits=8, rtty_stops=2.0, dec_factor=1)
    assert dec.setupDecimationStagesBW(max_rate) == 0                  # Decoder.h:340-341: no sampling rate yet
    dec.pushSamples(0, iq[:chunk], fs)
    got = dec.setupDecimationStagesBW(max_rate)
    assert got == want
    if not want:
        return
    assert dec.getDecimationFactor() == want and dec.getDecimatedSamplingRate() == fs / want
    dec.process()
    for o in range(chunk, n, chunk):
        dec.pushSamples(0, iq[o:o + chunk], fs)
        dec.process()
    ref = make_oracle(oracle_kind, baud=baud, rtty_bits=8, rtty_stops=2.0, dec_factor=want).run(iq[:n], fs, chunk)
    assert dec.poll_chars(0) == ref.chars()
    assert dec.poll_sentences(0) == ref.sentences()     # (not every rate/limit pair decodes in the reference either)
    assert dec.poll_raw_chars(0) is not None


@pytest.mark.parametrize("max_rate,want,stages,baud,chunk", [(1000.0, 2048, "64,4,8", 50.0, 131072), (3000.0, 1024, "64,4,4", 100.0, 65536),
                                                             (600.0, 4096, "64,4,8,2", 25.0, 262144)])
def test_setup_decimation_stages_bw_cascade(oracle_kind, max_rate, want, stages, baud, chunk):
    """Decoder.h:350-399: the while loop keeps appending plans of up to 256 until the rate is under the limit, so a limit far
    below the input rate cascades three or four decimators (K1, the middle-stage kernel(s), the tail kernel).  Same call
    sequence on both sides: one one-sample process() at factor 1 (it latches the rate), then the cascade; decimated /
    filtered / demodulated streams within 1e-5, characters, sentences and AFC peaks exact."""
    fs = 2.048e6
    iq, _ = synth.channel_iq(5, 1, fs, baud, snr_db=-20.0)
    n = (len(iq) - 1) // chunk * chunk
    cfg = dict(baud=baud, rtty_bits=8, rtty_stops=2.0, dec_factor=1)
    dec = api.BatchDecoder(1, record=True, **cfg)
    ref = make_oracle(oracle_kind, **cfg)
    got = {"dec": [], "filt": [], "demod": []}
    blocks = [iq[:1]] + [iq[1 + o:1 + o + chunk] for o in range(0, n, chunk)]
    for i, blk in enumerate(blocks):
        dec.pushSamples(0, blk, fs)
        dec.process()
        ref.push_process(blk, fs)
        got["dec"].append(dec.debug_stage(0, api.STAGE_DECIMATED).copy())
        got["filt"].append(dec.debug_stage(0, api.STAGE_FILTERED).copy())
        got["demod"].append(dec.debug_stage(0, api.STAGE_DEMOD).copy())
        if i == 0:
            assert dec.setupDecimationStagesBW(max_rate) == want
            ref.set_param("decimation_bw", max_rate)
            assert dec.getDecimationFactor() == want and dec.getDecimatedSamplingRate() == fs / want
    check_stages({k: np.concatenate(v) for k, v in got.items()}, ref)
    assert dec.poll_chars(0) == ref.chars()
    assert dec.poll_sentences(0) == ref.sentences()
    assert len(ref.sentences()) >= 1, "the %s plan decodes nothing in the reference either: pick another signal" % stages
    a = ref.afc()
    assert dec.getPeaks(0) == (a.peak_left, a.peak_right)


def test_lowpass_setters_between_calls(oracle_kind):
    """Decoder::lowpass_bw / lowpass_trans re-run the design at once (Decoder.h:238-257): a new bandwidth with an unchanged
    tap count changes nothing (FirFilter.h:193-194); a wider transition gives a SHORTER filter whose history is the oldest
    part of the previous one (the work buffer is kept, FirFilter.h:139-152).  Both reproduced; filtered stream compared."""
    fs, baud = 2.048e6, 300.0
    iq, _ = synth.channel_iq(41, 3, fs, baud, snr_db=-15.0)
    n = len(iq) // 65536 * 65536
    cfg = dict(baud=baud, rtty_bits=8, rtty_stops=2.0, dec_factor=256)
    dec = api.BatchDecoder(1, record=True, **cfg)
    ref = make_oracle(oracle_kind, **cfg)
    filt = []
    for i, o in enumerate(range(0, n, 65536)):
        if i == 6:
            dec.lowpass_bw(900.0, 0); ref.set_param("lowpass_bw", 900.0)
        if i == 14:
            dec.lowpass_trans(0.05, 0); ref.set_param("lowpass_trans", 0.05)
        dec.pushSamples(0, iq[o:o + 65536], fs)
        dec.process()
        ref.push_process(iq[o:o + 65536], fs)
        filt.append(dec.debug_stage(0, api.STAGE_FILTERED).copy())
    got, want = np.concatenate(filt), ref.stage(po.STAGE_FILTERED)
    assert got.shape == want.shape and rel_l2(got, want) <= REL_L2
    # the call right after the shrink is where a wrong history would show: compare it on its own
    k = 14 * 256
    assert rel_l2(got[k:k + 256], want[k:k + 256]) <= REL_L2
    assert np.array_equal(dec.debug_stage(0, api.STAGE_LPTAPS), ref.stage(po.STAGE_LPTAPS)) and len(ref.stage(po.STAGE_LPTAPS)) == 81
    assert dec.poll_chars(0) == ref.chars()
    assert dec.poll_sentences(0) == ref.sentences()


@pytest.mark.parametrize("baud0", [300.0, 50.0])
def test_slicer_mask_cache_large_window_with_baud_changes(oracle_kind, monkeypatch, baud0):
    """SymbolExtractor.h:162-224 at fs_dec = 78 125 Hz (window radius 65 at 300 baud, 390 at 50 baud): the tail kernel keeps
    the flip-search masks of the pending samples across calls (tail.cu).  Irregular chunks -- calls that only append, calls
    that erase, empty pushes -- and baud changes between calls (another radius: the cache has to be dropped) must give the
    reference's characters and leave the reference's pending samples behind; HBD_MASK_CACHE=0 (masks rebuilt from scratch
    in every call) is run beside it as a second witness."""
    fs, factor = 312500.0, 4
    a, _ = synth.channel_iq(41, 2, fs, baud0, snr_db=-3.0)
    b, _ = synth.channel_iq(42, 2, fs, 100.0, nbits=7, nstops=1, snr_db=-3.0)
    rng = np.random.default_rng(5)
    noise = (0.7 * (rng.standard_normal(150000) + 1j * rng.standard_normal(150000))).astype(np.complex64)
    iq = np.concatenate([a, noise, b, a[:len(a) // 2]])
    pattern = [1024, 2048, 0, 1031, 4096, 16384, 1024, 1024, 9500, 0, 32768, 1500, 65536, 1024]
    sizes, left, k = [], len(iq), 0
    while left > 0:
        n = min(pattern[k % len(pattern)], left)
        if 0 < left - n < 1024:
            n = left
        sizes.append(n); left -= n; k += 1
    switch = {len(a) + len(noise): [("baud", 100.0), ("rtty_bits", 7), ("rtty_stops", 1.0)],
              len(a) + len(noise) + len(b): [("baud", baud0), ("rtty_bits", 8), ("rtty_stops", 2.0)]}
    cfg = dict(baud=baud0, rtty_bits=8, rtty_stops=2.0, dec_factor=factor)
    monkeypatch.setenv("HBD_MASK_CACHE", "0")
    plain = api.BatchDecoder(1, **cfg)
    monkeypatch.delenv("HBD_MASK_CACHE")
    dec = api.BatchDecoder(1, **cfg)
    ref = make_oracle(oracle_kind, **cfg)
    o, done = 0, set()
    for n in sizes:
        for at, sets in switch.items():
            if o >= at and at not in done:
                done.add(at)
                for name, v in sets:
                    for d in (dec, plain):
                        getattr(d, name)(v, 0)
                    ref.set_param(name, v)
        blk = iq[o:o + n]
        o += n
        for d in (dec, plain):
            d.pushSamples(0, blk, fs)
            d.process()
        ref.push_process(blk, fs)
        got_p, want_p = dec.debug_stage(0, api.STAGE_PENDING), ref.stage(po.STAGE_PENDING)
        assert got_p.shape == want_p.shape, "pending samples after %d input samples" % o
        assert plain.debug_stage(0, api.STAGE_PENDING).shape == want_p.shape
    assert len(done) == 2
    assert dec.poll_chars(0) == ref.chars()
    assert plain.poll_chars(0) == ref.chars()
    assert dec.poll_sentences(0) == ref.sentences()
    assert len(ref.sentences()) >= 3

