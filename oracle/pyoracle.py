"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the two CPU oracles (oracle/oracle_abi.h).

  RefDecoder  -> oracle/_ref/libhabdec_ref.so  (the unmodified reference Decoder<float>)
  PortDecoder -> oracle/libhabdec_oracle.so    (our CPU restatement)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

STAGE_DECIMATED, STAGE_FILTERED, STAGE_DEMOD, STAGE_FFT, STAGE_POWER, STAGE_LPTAPS, STAGE_PENDING, STAGE_BITS, STAGE_RAWCHARS = range(9)


class Config(C.Structure):
    _fields_ = [("baud", C.c_double), ("rtty_bits", C.c_int), ("rtty_stops", C.c_float),
                ("lowpass_bw", C.c_float), ("lowpass_trans", C.c_float), ("dec_factor", C.c_int),
                ("dc_remove", C.c_int), ("record", C.c_int), ("fft_bins", C.c_int)]


class SsdvEvent(C.Structure):
    _fields_ = [("call", C.c_uint32), ("image_id", C.c_uint16), ("packet_id", C.c_uint16), ("width", C.c_uint16),
                ("height", C.c_uint16), ("set_size", C.c_uint16), ("reserved", C.c_uint16), ("set_crc32", C.c_uint32),
                ("callsign", C.c_char * 8)]

    def astuple(self):
        return (self.call, self.callsign.decode(), self.image_id, self.packet_id, self.width, self.height, self.set_size, self.set_crc32)


class AfcInfo(C.Structure):
    _fields_ = [("frequency_correction", C.c_double), ("shift_hz", C.c_double), ("noise_floor", C.c_double),
                ("noise_variance", C.c_double), ("peak_left", C.c_int), ("peak_right", C.c_int)]


def make_config(baud=300.0, rtty_bits=8, rtty_stops=2.0, lowpass_bw=1500.0, lowpass_trans=0.025,
                dec_factor=256, dc_remove=False, record=True, fft_bins=0) -> Config:
    return Config(float(baud), int(rtty_bits), float(rtty_stops), float(lowpass_bw), float(lowpass_trans),
                  int(dec_factor), int(bool(dc_remove)), int(bool(record)), int(fft_bins))


def _bind(lib, p):
    f = lambda name: getattr(lib, p + "_" + name)
    f("create").restype = C.c_void_p
    f("create").argtypes = [C.POINTER(Config)]
    f("destroy").argtypes = [C.c_void_p]
    f("push_process").argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double]
    for name in ("chars", "rtty", "last_sentence", "sentences"):
        f(name).restype = C.c_size_t
        f(name).argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    f("stage").restype = C.c_size_t
    f("stage").argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    f("afc").argtypes = [C.c_void_p, C.POINTER(AfcInfo)]
    f("reset_frequency_correction").argtypes = [C.c_void_p, C.c_double]
    f("set_param").argtypes = [C.c_void_p, C.c_int, C.c_double]
    f("ssdv_events").restype = C.c_size_t
    f("ssdv_events").argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    f("ssdv_push").argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    f("ssdv_image").restype = C.c_size_t
    f("ssdv_image").argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_size_t]
    f("bench").restype = C.c_double
    f("bench").argtypes = [C.POINTER(Config), C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                           C.c_double, C.c_int, C.POINTER(C.c_uint64)]


    f("run_ring").argtypes = [C.POINTER(Config), C.c_int, C.c_void_p] + [C.c_size_t] * 6 + [C.c_double, C.c_void_p, C.c_size_t, C.c_void_p,
                              C.c_void_p, C.c_size_t, C.c_void_p]


_libs = {}


def _load(kind):
    if kind in _libs:
        return _libs[kind]
    path = os.path.join(_HERE, "_ref", "libhabdec_ref.so") if kind == "ref" else os.path.join(_HERE, "libhabdec_oracle.so")
    if not os.path.exists(path):
        raise FileNotFoundError(path + " (run `make -C oracle` / __graft_entry__.build())")
    lib = C.CDLL(path)
    _bind(lib, kind)
    _libs[kind] = lib
    return lib


def available(kind: str) -> bool:
    try:
        _load(kind)
        return True
    except (OSError, FileNotFoundError):
        return False


class _Decoder:
    kind = ""

    def __init__(self, cfg: Config | None = None, **kw):
        self._lib = _load(self.kind)
        self.cfg = cfg if cfg is not None else make_config(**kw)
        self._f = lambda name: getattr(self._lib, self.kind + "_" + name)
        self._h = self._f("create")(C.byref(self.cfg))

    def close(self):
        if self._h:
            self._f("destroy")(self._h)
            self._h = None

    __del__ = close

    def push_process(self, iq: np.ndarray, fs: float):
        iq = np.ascontiguousarray(iq, dtype=np.complex64)
        self._f("push_process")(self._h, iq.ctypes.data, iq.size, float(fs))

    def run(self, iq: np.ndarray, fs: float, chunk: int = 65536):
        for o in range(0, len(iq), chunk):
            self.push_process(iq[o:o + chunk], fs)
        return self

    def _str(self, name) -> bytes:
        n = self._f(name)(self._h, None, 0)
        buf = C.create_string_buffer(max(n, 1))
        self._f(name)(self._h, buf, n)
        return buf.raw[:n]

    def chars(self) -> bytes:
        return self._str("chars")

    def rtty(self) -> bytes:
        return self._str("rtty")

    def last_sentence(self) -> bytes:
        return self._str("last_sentence")

    def sentences(self) -> list[bytes]:
        s = self._str("sentences")
        return [x for x in s.split(b"\n") if x]

    def stage(self, which: int) -> np.ndarray:
        n = self._f("stage")(self._h, which, None, 0)
        out = np.empty(n, dtype=np.float32)
        if n:
            self._f("stage")(self._h, which, out.ctypes.data, n)
        if which in (STAGE_DECIMATED, STAGE_FILTERED, STAGE_FFT):
            return out.view(np.complex64)
        return out

    def afc(self) -> AfcInfo:
        a = AfcInfo()
        self._f("afc")(self._h, C.byref(a))
        return a

    def set_param(self, which: str, value: float):
        """Run-time setter between calls (Decoder.h:654-706): 'baud', 'rtty_bits', 'rtty_stops', 'dc_remove', 'lowpass_bw', 'lowpass_trans';
        'decimation_bw' = setupDecimationStagesBW(value) (Decoder.h:336-412; after the first push)."""
        self._f("set_param")(self._h, {"baud": 0, "rtty_bits": 1, "rtty_stops": 2, "dc_remove": 3, "lowpass_bw": 4, "lowpass_trans": 5, "decimation_bw": 6}[which], float(value))

    def reset_frequency_correction(self, corr: float):
        self._f("reset_frequency_correction")(self._h, float(corr))

    # ---- SSDV packet sync / bookkeeping (SSDV_wraper_t::push, ssdv_wrapper.cpp:37-148) ----
    def ssdv_push(self, chars: bytes):
        """One SSDV_wraper_t::push(chars) on the decoder's wrapper, like Decoder.h:573."""
        buf = (C.c_ubyte * max(len(chars), 1)).from_buffer_copy(bytes(chars) or b"\0")
        self._f("ssdv_push")(self._h, buf, len(chars))

    def ssdv_events(self) -> list[tuple]:
        n = self._f("ssdv_events")(self._h, None, 0)
        arr = (SsdvEvent * max(n, 1))()
        self._f("ssdv_events")(self._h, arr, n)
        return [arr[i].astuple() for i in range(n)]

    def ssdv_image(self, callsign: str, image_id: int) -> bytes:
        n = self._f("ssdv_image")(self._h, callsign.encode(), int(image_id), None, 0)
        buf = (C.c_ubyte * max(n, 1))()
        self._f("ssdv_image")(self._h, callsign.encode(), int(image_id), buf, n)
        return bytes(buf[:n])


class RefDecoder(_Decoder):
    kind = "ref"


class PortDecoder(_Decoder):
    kind = "orc"


class SpectrumMeta(C.Structure):
    _fields_ = [("noise_floor", C.c_double), ("noise_variance", C.c_double), ("sampling_rate", C.c_double), ("shift", C.c_double),
                ("peak_left", C.c_int), ("peak_right", C.c_int)]


def spectrum_frame(kind: str, power: np.ndarray, meta: SpectrumMeta, zoom: float, resolution: int, type_size: int) -> bytes:
    """PWR_ payload (header + quantised bins) of cmd::power:res=R,zoom=Z -- habdec_ws_protocol.cpp:353-404."""
    lib = _load(kind)
    fn = getattr(lib, kind + "_spectrum_frame")
    fn.restype = C.c_size_t
    fn.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(SpectrumMeta), C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    power = np.ascontiguousarray(power, dtype=np.float32)
    n = fn(power.ctypes.data, power.size, C.byref(meta), zoom, resolution, type_size, None, 0)
    buf = C.create_string_buffer(max(n, 1))
    fn(power.ctypes.data, power.size, C.byref(meta), zoom, resolution, type_size, buf, n)
    return buf.raw[:n]


def demod_frame(kind: str, demod: np.ndarray, resolution: int, type_size: int) -> bytes:
    """DEM_ payload of cmd::demod:res=R -- habdec_ws_protocol.cpp:408-429."""
    lib = _load(kind)
    fn = getattr(lib, kind + "_demod_frame")
    fn.restype = C.c_size_t
    fn.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    demod = np.ascontiguousarray(demod, dtype=np.float32)
    n = fn(demod.ctypes.data, demod.size, resolution, type_size, None, 0)
    buf = C.create_string_buffer(max(n, 1))
    fn(demod.ctypes.data, demod.size, resolution, type_size, buf, n)
    return buf.raw[:n]


def ssdv_is_packet(pkt: bytes):
    """Published-algorithm packet test (oracle/ssdv_published.h): (verdict, errors, corrected packet)."""
    lib = _load("orc")
    lib.hbo_ssdv_is_packet.restype = C.c_int
    lib.hbo_ssdv_is_packet.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    buf = (C.c_ubyte * 256).from_buffer_copy(bytes(pkt))
    err = C.c_int(0)
    v = lib.hbo_ssdv_is_packet(buf, C.byref(err))
    return v, err.value, bytes(buf)


def ssdv_make_packet(callsign: str, image_id: int, packet_id: int, payload: bytes, fec: bool = True, width16: int = 4,
                     height16: int = 3, flags: int = 0, mcu_offset: int = 0, mcu_id: int = 0) -> bytes:
    """A well-formed SSDV packet of the published format (test-vector builder)."""
    lib = _load("orc")
    lib.hbo_ssdv_make_packet.argtypes = [C.c_void_p, C.c_int, C.c_char_p] + [C.c_int] * 7 + [C.c_void_p]
    need = 205 if fec else 237
    pl = (bytes(payload) + bytes(need))[:need]
    out = (C.c_ubyte * 256)()
    src = (C.c_ubyte * need).from_buffer_copy(pl)
    lib.hbo_ssdv_make_packet(out, 0 if fec else 1, callsign.encode(), image_id, packet_id, width16, height16, flags, mcu_offset, mcu_id, src)
    return bytes(out)


def premix(iq: np.ndarray, fs: float, f_hz: float, phase0: float = 0.0):
    """The oracle of the NCO pre-mixer ("CPU pre-mix, then the reference Decoder").

    y[i] = x[i] * exp(-2 pi i * frac(phase0 + i * f_hz / fs)): phase in float64, phasor rounded to cf32, product
    formed like std::complex<float>::operator* (four float products, one float add/sub each, no FMA).
    Returns (y complex64, phase of the sample after the last one)."""
    iq = np.ascontiguousarray(iq, dtype=np.complex64)
    inc = float(f_hz) / float(fs)
    i = np.arange(iq.size, dtype=np.float64)
    ph = phase0 + i * inc
    ph -= np.floor(ph)
    pr = np.cos(2.0 * np.pi * ph).astype(np.float32)
    pi = (-np.sin(2.0 * np.pi * ph)).astype(np.float32)
    xr, xi = iq.real.copy(), iq.imag.copy()
    y = np.empty(iq.size, dtype=np.complex64)
    y.real = (xr * pr).astype(np.float32) - (xi * pi).astype(np.float32)
    y.imag = (xr * pi).astype(np.float32) + (xi * pr).astype(np.float32)
    end = phase0 + iq.size * inc
    return y, end - np.floor(end)


def bench(kind: str, cfg: Config, iq: np.ndarray, n_threads: int, fs: float, chunk: int = 65536,
          reps: int = 1, stride: int = 0):
    """Returns (seconds, total_chars). iq: complex64[n] (stride 0: shared) or [n_threads, n]."""
    lib = _load(kind)
    iq = np.ascontiguousarray(iq, dtype=np.complex64)
    n = iq.shape[-1]
    if iq.ndim == 2:
        stride = n
    chars = C.c_uint64(0)
    secs = getattr(lib, kind + "_bench")(C.byref(cfg), n_threads, iq.ctypes.data, n, stride, chunk, float(fs), reps,
                                         C.byref(chars))
    return secs, chars.value


def run_ring(kind: str, cfg: Config, iq: np.ndarray, n_threads: int, fs: float, chunk: int, first_chunk: int, n_chunks: int,
             pitch: int = 4096):
    """Whole-batch parity checker: iq complex64 [n_channels, ring_n] (periodic streams); every channel is decoded by its
    own decoder on a fresh OS thread, `n_threads` at a time, over chunks first_chunk .. first_chunk+n_chunks-1.
    Returns ([printable chars per channel], [CRC-valid sentences per channel])."""
    lib = _load(kind)
    iq = np.ascontiguousarray(iq, dtype=np.complex64)
    n_ch, ring_n = iq.shape
    chars = np.zeros((n_ch, pitch), dtype=np.uint8)
    sents = np.zeros((n_ch, pitch), dtype=np.uint8)
    cl, sl = np.zeros(n_ch, dtype=np.uint32), np.zeros(n_ch, dtype=np.uint32)
    getattr(lib, kind + "_run_ring")(C.byref(cfg), int(n_threads), iq.ctypes.data, n_ch, ring_n, ring_n, int(chunk), int(first_chunk),
                                     int(n_chunks), float(fs), chars.ctypes.data, pitch, cl.ctypes.data, sents.ctypes.data, pitch, sl.ctypes.data)
    assert int(cl.max(initial=0)) <= pitch and int(sl.max(initial=0)) <= pitch, "run_ring: output pitch too small"
    return ([chars[c, :cl[c]].tobytes() for c in range(n_ch)],
            [[x for x in sents[c, :sl[c]].tobytes().split(b"\n") if x] for c in range(n_ch)])


def port_extract_sentence(stream: bytes):
    """std::regex extraction of the restatement (== the reference's extractSentence); None if no match."""
    lib = _load("orc")
    lib.orc_extract_sentence.restype = C.c_int
    lib.orc_extract_sentence.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
    cs, data, crc = (C.create_string_buffer(len(stream) + 8) for _ in range(3))
    rest = C.c_size_t(0)
    if not lib.orc_extract_sentence(stream, len(stream), cs, data, crc, len(stream) + 8, C.byref(rest)):
        return None
    return cs.value, data.value, crc.value, rest.value


def port_crc16(s: bytes) -> bytes:
    lib = _load("orc")
    lib.orc_crc16.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
    out = C.create_string_buffer(5)
    lib.orc_crc16(s, len(s), out)
    return out.value
