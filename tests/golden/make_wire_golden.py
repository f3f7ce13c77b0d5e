"""Regenerate tests/golden/wire_frames.npz: PWR_/DEM_ payloads produced by the reference's own serialisation code
(oracle/_ref/libhabdec_ref.so: ref_spectrum_frame / ref_demod_frame) for a fixed set of inputs."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

CASES = [  # zoom, resolution, type_size
    (0.0, 4096, 4), (0.0, 1024, 1), (0.5, 512, 1), (0.5, 512, 2), (0.9, 300, 2), (0.25, 8192, 1), (0.3, 1, 1), (2.0, 77, 4),
]
DEMOD_CASES = [(1350, 600, 1), (1350, 4000, 2), (500, 123, 4), (256, 256, 1)]


def inputs():
    rng = np.random.default_rng(2024)
    power = (-90.0 + 8.0 * rng.standard_normal(4096)).astype(np.float32)
    power[1900:1910] += 40.0
    power[2150:2160] += 38.0
    meta = po.SpectrumMeta(-88.123456789, 7.987654321, 8000.0, 488.28125, 1905, -2154)
    demod = (0.17 * np.sign(np.sin(np.arange(1350) * 0.05)) + 0.05 * rng.standard_normal(1350)).astype(np.float32)
    flat = np.full(300, -77.5, dtype=np.float32)       # max == min: the 0/0 quantisation corner
    return power, meta, demod, flat


def frames(kind):
    power, meta, demod, flat = inputs()
    out = {}
    for i, (zoom, res, ts) in enumerate(CASES):
        out["pwr_%d" % i] = np.frombuffer(po.spectrum_frame(kind, power, meta, zoom, res, ts), dtype=np.uint8)
    for i, (n, res, ts) in enumerate(DEMOD_CASES):
        out["dem_%d" % i] = np.frombuffer(po.demod_frame(kind, demod[:n], res, ts), dtype=np.uint8)
    out["pwr_flat"] = np.frombuffer(po.spectrum_frame(kind, flat, meta, 0.1, 100, 1), dtype=np.uint8)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "wire_frames.npz"), **frames("ref"))
    print("written", {k: len(v) for k, v in frames("ref").items()})
