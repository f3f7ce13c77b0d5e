"""Deterministic synthetic RTTY/2-FSK IQ generator shared by tests and bench.

Signal definition follows SURVEY.md section 8(d): UKHAS-style sentences
"$$<body>*<CRC16>\\n", UART framing (start 0, n data bits LSB first, stop 1s),
continuous-phase 2-FSK with bit '1' on the HIGHER tone (the reference slices the
discriminator at 0 with `avg > 0` == 1, code/Decoder/SymbolExtractor.h:149 and
code/Decoder/FSK2_Demod.h:38), unit amplitude, AWGN with sigma given for the
full sampling band.
"""
from __future__ import annotations

import numpy as np


def crc16_ccitt(data: bytes) -> str:
    """CRC16-CCITT-FALSE rendered as 4 upper-case hex digits (reference: code/Decoder/CRC.cpp:21-47)."""
    crc = 0xFFFF
    for b in data:
        crc ^= b << 8
        for _ in range(8):
            crc = ((crc << 1) ^ 0x1021) & 0xFFFF if crc & 0x8000 else (crc << 1) & 0xFFFF
    return "%04X" % crc


def make_sentence(channel: int, frame: int, altitude: int | None = None) -> str:
    alt = 1000 + 10 * frame if altitude is None else altitude
    body = "CH%04d,%d,12:00:00,52.1234,21.5678,%d" % (channel, frame, alt)
    return "$$" + body + "*" + crc16_ccitt(body.encode()) + "\n"


def uart_bits(payload: bytes, nbits: int = 8, nstops: int = 2, lead_in: int = 40, lead_out: int = 60) -> np.ndarray:
    bits = [1] * lead_in
    for c in payload:
        bits.append(0)
        bits.extend((c >> k) & 1 for k in range(nbits))
        bits.extend([1] * nstops)
    bits.extend([1] * lead_out)
    return np.asarray(bits, dtype=np.int8)


def fsk_iq(bits: np.ndarray, fs: float, baud: float, shift: float = 425.0, f_off=0.0,
           snr_db: float | None = None, seed: int = 1234, n_samples: int | None = None,
           phase0: float = 0.0) -> np.ndarray:
    """Continuous-phase 2-FSK. Returns complex64[n]. `f_off` may be a scalar or an array[n] (Hz)."""
    n_total = int(len(bits) * fs / baud) if n_samples is None else int(n_samples)
    n = np.arange(n_total, dtype=np.float64)
    bi = np.minimum((n * (baud / fs)).astype(np.int64), len(bits) - 1)
    f = np.where(bits[bi] > 0, 0.5 * shift, -0.5 * shift) + f_off
    ph = phase0 + np.cumsum(2.0 * np.pi * f / fs)
    iq = np.empty(n_total, dtype=np.complex64)
    iq.real = np.cos(ph)
    iq.imag = np.sin(ph)
    if snr_db is not None:
        rng = np.random.default_rng(seed)
        sigma = 10.0 ** (-snr_db / 20.0) / np.sqrt(2.0)
        iq.real += (sigma * rng.standard_normal(n_total)).astype(np.float32)
        iq.imag += (sigma * rng.standard_normal(n_total)).astype(np.float32)
    return iq


def channel_iq(channel: int, n_sentences: int, fs: float, baud: float, nbits: int = 8, nstops: int = 2,
               shift: float = 425.0, f_off=0.0, snr_db: float | None = None,
               n_samples: int | None = None, lead_in: int = 40, lead_out: int = 60):
    """One channel of the standard workload: returns (iq complex64[n], expected_text)."""
    text = "".join(make_sentence(channel, k) for k in range(n_sentences))
    bits = uart_bits(text.encode(), nbits, nstops, lead_in, lead_out)
    iq = fsk_iq(bits, fs, baud, shift, f_off, snr_db, seed=1234 + channel, n_samples=n_samples)
    return iq, text
