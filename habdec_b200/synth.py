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


# ---------------------------------------------------------------------------------------------------
# Periodic "ring" workload for bench.py (BASELINE.json configs[3]: 300 baud 8N2, many channels).
# One ring = 240 bit periods = 21 UART characters (one CRC-valid sentence) + 9 idle bits; at fs = 2.048 MS/s and
# 300 baud that is exactly 1 638 400 samples (3 bits == 20 480 samples), i.e. 25 chunks of 65 536.
# The sentence has 21 characters because the reference only scans its text stream for sentences once it holds more
# than 20 (Decoder.h:591): every ring pass then yields one extracted sentence per channel.
# A sub-Hz frequency trim makes the phase wrap exactly, so replaying the ring is an endless,
# continuous-phase RTTY stream.
# ---------------------------------------------------------------------------------------------------
RING_BITS = 240


def ring_sentence(channel: int) -> str:
    body = "C%04d,%03d,%03d" % (channel % 10000, (channel * 7 + 123) % 1000, (channel * 13 + 7) % 1000)
    return "$$" + body + "*" + crc16_ccitt(body.encode()) + "\n"      # 21 characters


def ring_bits(channel: int, nbits: int = 8, nstops: int = 2) -> np.ndarray:
    bits = uart_bits(ring_sentence(channel).encode(), nbits, nstops, lead_in=0, lead_out=0)
    assert len(bits) <= RING_BITS
    return np.concatenate([bits, np.ones(RING_BITS - len(bits), dtype=np.int8)])


def ring_length(fs: float, baud: float) -> int:
    n = RING_BITS * fs / baud
    assert abs(n - round(n)) < 1e-9, "ring must be a whole number of samples"
    return int(round(n))


def ring_iq_numpy(channel: int, fs: float = 2.048e6, baud: float = 300.0, shift: float = 425.0,
                  snr_db: float | None = -15.0) -> np.ndarray:
    """One channel of the ring workload, complex64[ring_length]."""
    L = ring_length(fs, baud)
    bits = ring_bits(channel)
    n = np.arange(L, dtype=np.float64)
    bi = np.minimum((n * (baud / fs)).astype(np.int64), RING_BITS - 1)
    f = np.where(bits[bi] > 0, 0.5 * shift, -0.5 * shift)
    total = 2.0 * np.pi * f.sum() / fs
    trim = (np.round(total / (2 * np.pi)) * 2 * np.pi - total) / L          # radians per sample, |trim| tiny
    ph = np.cumsum(2.0 * np.pi * f / fs + trim)
    iq = np.empty(L, dtype=np.complex64)
    iq.real = np.cos(ph)
    iq.imag = np.sin(ph)
    if snr_db is not None:
        rng = np.random.default_rng(99991 + channel)
        sigma = 10.0 ** (-snr_db / 20.0) / np.sqrt(2.0)
        iq.real += (sigma * rng.standard_normal(L)).astype(np.float32)
        iq.imag += (sigma * rng.standard_normal(L)).astype(np.float32)
    return iq


def ring_iq_torch(ch0: int, n_ch: int, device, fs: float = 2.048e6, baud: float = 300.0, shift: float = 425.0,
                  snr_db: float | None = -15.0, slice_channels: int = 32):
    """Channels ch0..ch0+n_ch-1 of the ring workload generated on `device`: float32 tensor [n_ch, L, 2]
    (interleaved cf32, row pitch L samples).  Same construction as ring_iq_numpy (different noise stream).  The noise
    generator is re-seeded for every slice of 32 channels from the slice's first GLOBAL channel number, so channel c gets
    the same samples whether it is generated as part of a 512- or a 4096-channel block (strong-scaling invariance)."""
    import torch
    L = ring_length(fs, baud)
    out = torch.empty((n_ch, L, 2), dtype=torch.float32, device=device)
    n = torch.arange(L, dtype=torch.float64, device=device)
    bi = torch.clamp((n * (baud / fs)).to(torch.int64), max=RING_BITS - 1)
    gen = torch.Generator(device=device)
    sigma = None if snr_db is None else 10.0 ** (-snr_db / 20.0) / np.sqrt(2.0)
    assert ch0 % slice_channels == 0, "channel blocks start on a multiple of %d (the noise is seeded per slice)" % slice_channels
    for s in range(0, n_ch, slice_channels):
        e = min(n_ch, s + slice_channels)
        gen.manual_seed(424242 + ch0 + s)      # a channel's samples depend on its GLOBAL number only, not on the sharding
        bits = torch.from_numpy(np.stack([ring_bits(ch0 + c) for c in range(s, e)])).to(device)
        f = torch.where(bits[:, bi] > 0, 0.5 * shift, -0.5 * shift).to(torch.float64)
        total = 2.0 * np.pi * f.sum(dim=1, keepdim=True) / fs
        trim = (torch.round(total / (2 * np.pi)) * 2 * np.pi - total) / L
        ph = torch.cumsum(2.0 * np.pi * f / fs + trim, dim=1)
        out[s:e, :, 0] = torch.cos(ph).to(torch.float32)
        out[s:e, :, 1] = torch.sin(ph).to(torch.float32)
        del f, ph
        if sigma is not None:
            out[s:e] += sigma * torch.randn((e - s, L, 2), dtype=torch.float32, device=device, generator=gen)
    return out
