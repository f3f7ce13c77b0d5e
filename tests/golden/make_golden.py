"""Regenerates the golden fixtures from the UNMODIFIED reference (oracle/_ref/libhabdec_ref.so, which is
compiled from /root/reference/code by `make -C oracle ref`).  Run in the dev container only:

    python tests/golden/make_golden.py

Inputs are stored with the fixture, quantised to int8 I/Q (value/SCALE is exactly representable in float32),
so the fixtures do not depend on any platform's libm or RNG."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from habdec_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

SCALE = 32.0

CASES = {
    # name: (fs, baud, bits, stops, factor, snr_db, n_samples or None (whole message), n_sentences, chunk, dc_remove)
    "g1_dec256_2048k_300bd": (2.048e6, 300.0, 8, 2.0, 256, -14.0, 6 * 65536, 1, 65536, False),
    "g2_dec2_16k_300bd": (16e3, 300.0, 8, 2.0, 2, -3.0, None, 1, 4096, False),
    "g3_dec32_256k_600bd_7n1_dc": (256e3, 600.0, 7, 1.0, 32, -9.0, None, 1, 65536, True),
}


# a fourth fixture drives the reference with an irregular call pattern and run-time setters:
#   chunks: sizes of the successive pushes (0 = empty push; short ones hit the in-place stage-2 history, Decoder.h:441-446;
#           growing ones re-zero the histories, Decimator.h:74-79);  events: (call index, setter, value) applied before that call
G4_NAME = "g4_dec256_2048k_600bd_chunks_setters"
G4_CHUNKS = [65536, 0, 70001, 9500, 100000, 65536, 0, 131072, 12345, 40000, 65537, 9999, 200000, 65536]
G4_EVENTS = [(4, "lowpass_bw", 900.0), (7, "dc_remove", 1.0), (9, "rtty_stops", 1.0), (10, "lowpass_trans", 0.05), (12, "baud", 300.0)]
SETTERS = ["baud", "rtty_bits", "rtty_stops", "dc_remove", "lowpass_bw", "lowpass_trans"]


def quantise(iq):
    q = np.empty((len(iq), 2), dtype=np.int8)
    q[:, 0] = np.clip(np.round(iq.real * SCALE), -127, 127)
    q[:, 1] = np.clip(np.round(iq.imag * SCALE), -127, 127)
    return q


def dequantise(q):
    return (q[:, 0].astype(np.float32) / np.float32(SCALE) + 1j * (q[:, 1].astype(np.float32) / np.float32(SCALE))).astype(np.complex64)


def main():
    for name, (fs, baud, bits, stops, factor, snr, n, nsent, chunk, dc) in CASES.items():
        iq, text = synth.channel_iq(11, nsent, fs, baud, bits, int(stops), snr_db=snr, n_samples=n)
        q = quantise(iq)
        iq = dequantise(q)
        d = po.RefDecoder(po.make_config(baud=baud, rtty_bits=bits, rtty_stops=stops, dec_factor=factor, dc_remove=dc)).run(iq, fs, chunk)
        a = d.afc()
        out = dict(iq_q=q, fs=fs, baud=baud, bits=bits, stops=stops, factor=factor, chunk=chunk, dc_remove=dc,
                   decimated=d.stage(po.STAGE_DECIMATED), filtered=d.stage(po.STAGE_FILTERED), demod=d.stage(po.STAGE_DEMOD),
                   lptaps=d.stage(po.STAGE_LPTAPS), pending=d.stage(po.STAGE_PENDING),
                   fft_sha=hashlib.sha256(d.stage(po.STAGE_FFT).tobytes()).hexdigest(),
                   power=d.stage(po.STAGE_POWER).astype(np.float32),
                   chars=np.frombuffer(d.chars(), dtype=np.uint8), rtty=np.frombuffer(d.rtty(), dtype=np.uint8),
                   last_sentence=np.frombuffer(d.last_sentence(), dtype=np.uint8),
                   sentences=np.frombuffer(b"\n".join(d.sentences()), dtype=np.uint8),
                   afc=np.array([a.frequency_correction, a.shift_hz, a.noise_floor, a.noise_variance, a.peak_left, a.peak_right]),
                   text=np.frombuffer(text.encode(), dtype=np.uint8))
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, os.path.getsize(path) // 1024, "KiB", "chars", d.chars()[:40], "sentences", d.sentences())
    # ---- g4: irregular chunks + setters
    fs, baud = 2.048e6, 600.0
    iq, text = synth.channel_iq(11, 1, fs, baud, 8, 2, snr_db=-12.0, n_samples=sum(G4_CHUNKS))
    q = quantise(iq)
    iq = dequantise(q)
    d = po.RefDecoder(po.make_config(baud=baud, rtty_bits=8, rtty_stops=2.0, dec_factor=256))
    o = 0
    for i, n in enumerate(G4_CHUNKS):
        for call, name, value in G4_EVENTS:
            if call == i:
                d.set_param(name, value)
        d.push_process(iq[o:o + n], fs)
        o += n
    a = d.afc()
    out = dict(iq_q=q, fs=fs, baud=baud, bits=8, stops=2.0, factor=256, chunk=0, dc_remove=False,
               chunks=np.array(G4_CHUNKS, dtype=np.int64),
               events=np.array([(c, SETTERS.index(nm), v) for c, nm, v in G4_EVENTS], dtype=np.float64),
               decimated=d.stage(po.STAGE_DECIMATED), filtered=d.stage(po.STAGE_FILTERED), demod=d.stage(po.STAGE_DEMOD),
               lptaps=d.stage(po.STAGE_LPTAPS), pending=d.stage(po.STAGE_PENDING),
               fft_sha=hashlib.sha256(d.stage(po.STAGE_FFT).tobytes()).hexdigest(),
               power=d.stage(po.STAGE_POWER).astype(np.float32),
               chars=np.frombuffer(d.chars(), dtype=np.uint8), rtty=np.frombuffer(d.rtty(), dtype=np.uint8),
               last_sentence=np.frombuffer(d.last_sentence(), dtype=np.uint8),
               sentences=np.frombuffer(b"\n".join(d.sentences()), dtype=np.uint8),
               afc=np.array([a.frequency_correction, a.shift_hz, a.noise_floor, a.noise_variance, a.peak_left, a.peak_right]),
               text=np.frombuffer(text.encode(), dtype=np.uint8))
    path = os.path.join(HERE, G4_NAME + ".npz")
    np.savez_compressed(path, **out)
    print(G4_NAME, os.path.getsize(path) // 1024, "KiB", "chars", d.chars()[:40], len(d.stage(po.STAGE_LPTAPS)), "taps")


if __name__ == "__main__":
    main()
