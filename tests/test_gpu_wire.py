"""GPU parity of the websocket wire formats (PWR_ / DEM_ payloads): bytes produced by the CUDA kernels vs the oracle
restatement of habdec_ws_protocol.cpp / NetTransport.h / CompressedVector.cpp (itself pinned byte for byte to the
reference's own serialisation code, tests/test_oracle.py), fed with the GPU's own spectrum / discriminator values.
Bar: bit-exact (byte work)."""
import numpy as np
import pytest

from habdec_b200 import api, synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
CHUNK = 65536


def _decode_some(n_ch, fs, bauds, n_calls, chunk=CHUNK, accumulate=False, on_call=None, fft_bins=4096):
    iqs = [synth.channel_iq(c, 4, fs, bauds[c], snr_db=-14.0, f_off=40.0 * c, n_samples=n_calls * chunk)[0] for c in range(n_ch)]
    iq = np.stack(iqs)
    dec = api.BatchDecoder(n_ch, dec_factor=256, fft_bins=fft_bins)
    for c in range(n_ch):
        dec.baud(bauds[c], c)
    if accumulate:
        dec.set_demod_accumulate(True)
    for k in range(n_calls):
        dec.pushSamplesBatch(np.ascontiguousarray(iq[:, k * chunk:(k + 1) * chunk]), fs)
        dec.process()
        if on_call:
            on_call(dec, k)
    return dec


def _meta(dec, ch):
    pl, pr = dec.getPeaks(ch)
    nf, nv = dec.getNoiseFloor(ch)
    return po.SpectrumMeta(nf, nv, dec.getDecimatedSamplingRate(), dec.getShift(ch), pl, pr)


@pytest.mark.parametrize("nfft", [4096, 16384])
def test_spectrum_frames_bit_exact(nfft):
    fs = 2.048e6
    bauds = [300.0, 300.0, 100.0]
    dec = _decode_some(3, fs, bauds, 40 if nfft == 4096 else 140, fft_bins=nfft)
    assert dec.spectrum_frame(0, 0.5, 512, 1) != b""
    for zoom, res, ts in [(0.0, 4096, 4), (0.0, 1024, 1), (0.5, 512, 1), (0.5, 512, 2), (0.9, 300, 2), (0.25, 8192, 1), (0.3, 1, 1), (2.0, 77, 4), (0.4, 0, 1)]:
        batch = dec.spectrum_frames(zoom, res, ts)
        for ch in range(3):
            want = po.spectrum_frame("orc", dec.getPowerSpectrum(ch), _meta(dec, ch), zoom, res, ts)
            assert dec.spectrum_frame(ch, zoom, res, ts) == want, (zoom, res, ts, ch)
            assert batch[ch] == want, (zoom, res, ts, ch)
    # the peaks made it into the header for a decodable signal
    hdr = np.frombuffer(dec.spectrum_frame(0, 0.0, nfft, 4)[:52], dtype=np.int32)
    zb, ze = int(np.float32(0.005) * np.float32(nfft)), int(np.float32(0.995) * np.float32(nfft))
    assert hdr[0] == 52 and hdr[12] == ze - zb and hdr[11] == 4      # zoom is clamped to [0.01, 0.99]


def test_no_frame_before_the_first_spectrum():
    dec = api.BatchDecoder(2, dec_factor=256)
    assert dec.spectrum_frame(0, 0.5, 512, 1) == b""
    assert dec.spectrum_frames(0.5, 512, 1) == [b"", b""]
    dec.set_demod_accumulate(True)
    assert dec.demod_frame(1, 100, 1) == b""


@pytest.mark.parametrize("chunk", [65536, 40000])
def test_demod_frames_bit_exact(chunk):
    """Accumulation follows websocketServer/main.cpp:267-282: the last demodulated block is appended after EVERY call
    (a call that demodulates nothing re-appends the previous block: chunk 40000 exercises that) and trimmed to 50 symbols."""
    fs = 2.048e6
    bauds = [300.0, 50.0, 100.0]
    acc = [np.zeros(0, dtype=np.float32) for _ in bauds]
    checks = []

    def on_call(dec, k):
        fs_dec = dec.getDecimatedSamplingRate()
        for ch in range(3):
            acc[ch] = np.concatenate([acc[ch], dec.getDemodulated(ch)])
            max_sz = int(fs_dec / bauds[ch] * 50)
            if len(acc[ch]) > max_sz:
                acc[ch] = acc[ch][len(acc[ch]) - max_sz:]
        if k in (2, 9, 33, 59):
            for res, ts in [(600, 1), (4000, 2), (123, 4), (100000, 4)]:
                batch = dec.demod_frames(res, ts)
                for ch in range(3):
                    want = po.demod_frame("orc", acc[ch], res, ts)
                    assert dec.demod_frame(ch, res, ts) == want, (k, res, ts, ch)
                    assert batch[ch] == want
                    checks.append(len(want))

    _decode_some(3, fs, bauds, 60, chunk=chunk, accumulate=True, on_call=on_call)
    assert max(checks) > 20000 and len(acc[1]) == int(8000.0 / 50.0 * 50)
