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


if __name__ == "__main__":
    main()
