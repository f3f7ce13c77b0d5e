"""The C++17 host side (include/habdec_b200/Decoder.hpp + IQSource.hpp over the C ABI) used exactly like the reference's
DECODER_THREAD (code/websocketServer/main.cpp:203-283): a compiled program reads a cf32 file through the file IQ source,
pushes 65 536-sample IQVectors into habdec_b200::Decoder and prints what the callbacks and getters deliver; the output
has to equal what the reference Decoder (oracle) produces for the same file."""
import os
import subprocess

import numpy as np
import pytest

from habdec_b200 import synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "decoder_thread")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    lib_dir = os.path.join(ROOT, "habdec_b200")
    subprocess.run([cxx, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "decoder_thread.cpp"),
                    "-L", lib_dir, "-lhabdec_b200", "-Wl,-rpath," + lib_dir, "-lpthread", "-o", exe], check=True)
    return exe


def _run(exe, path, fs, baud, bits, stops, factor):
    out = subprocess.run([exe, path, repr(fs), repr(baud), str(bits), str(stops), str(factor)], check=True, capture_output=True).stdout
    head, _, tail = out.partition(b"CHARS ")
    n_chars, _, rest = tail.partition(b"\n")
    chars = rest[:int(n_chars)]
    assert rest[int(n_chars):].strip() == b"END"
    return head.decode("latin1").splitlines(), chars


def test_decoder_thread_program_matches_reference(tmp_path, oracle_kind):
    """tests/cpp/decoder_thread.cpp: a client of the C++ facade that makes the member calls the reference's server makes
    (feed loop, the three callback members, getSpectrumInfo() -> zoom / peak re-indexing / decimation -> PWR_ payload)."""
    import struct
    fs, baud = 2.048e6, 300.0
    iq, _ = synth.channel_iq(9, 3, fs, baud, snr_db=-15.0)
    n = len(iq) // 65536 * 65536
    path = str(tmp_path / "cap.cf32")
    iq[:n].astype(np.complex64).tofile(path)
    lines, chars = _run(_build(tmp_path), path, fs, baud, 8, 2, 256)

    ref = (po.RefDecoder if oracle_kind == "ref" else po.PortDecoder)(po.make_config(baud=baud)).run(iq[:n], fs)
    assert chars == ref.chars()
    sent = [l for l in lines if l.startswith("SENT ")]
    assert [l[5:].rsplit(" last=", 1)[0].encode() for l in sent] == ref.sentences()
    assert all(l.endswith(" last=1") for l in sent)          # getLastSentence() called from inside sentence_callback_
    assert "NSENT %d" % len(ref.sentences()) in lines and len(ref.sentences()) >= 2
    assert ("LAST " + ref.last_sentence().decode()) in lines
    assert "RATE 8000 256 4096" in lines
    a = ref.afc()
    assert "PEAKS %d %d" % (a.peak_left, a.peak_right) in lines
    # SpectrumToStream over getSpectrumInfo(): byte-identical to the library's own PWR_ frame (pinned to the reference's
    # serialiser in test_gpu_wire.py) and carrying the oracle's AFC numbers in its header
    pwr = [l.split() for l in lines if l.startswith("PWR ")]
    assert len(pwr) == 3 and all(p[4] == "same" for p in pwr), pwr
    assert [int(p[2]) for p in pwr] == [4075 - 20, 1024, 300]              # zoom 0.01 floor (bins [20, 4075)) / 0.5 / 0.9 and the shrink
    hdr = struct.unpack("<i4f4i2f2i", bytes.fromhex(pwr[0][5]))
    assert hdr[0] == 52 and hdr[11] == 4 and hdr[12] == 4055
    assert hdr[1] == pytest.approx(a.noise_floor, abs=1e-3) and hdr[3] == 8000.0 and hdr[4] == pytest.approx(a.shift_hz, abs=1e-3)
    assert hdr[5] == abs(a.peak_left) - 20 and hdr[6] == abs(a.peak_right) - 20 and hdr[7] == int(a.peak_left > 0) and hdr[8] == int(a.peak_right > 0)


def test_decoder_thread_ssdv_callback(tmp_path, oracle_kind):
    """ssdv_callback_(callsign, image_id, bytes) installed like websocketServer/main.cpp:587-604: one call per accepted
    packet, on the same process() call as the reference, with the image's packet set as of that moment."""
    import test_gpu_ssdv as tg
    fs, baud, factor, chunk = 256e3, 600.0, 32, 65536
    iq = tg._rtty_iq(tg._payload(2), fs, baud, seed=902, snr_db=-6.0)
    n = (len(iq) + chunk - 1) // chunk * chunk
    rng = np.random.default_rng(4)
    full = np.concatenate([iq, (0.3 * (rng.standard_normal(n - len(iq)) + 1j * rng.standard_normal(n - len(iq)))).astype(np.complex64)])
    path = str(tmp_path / "ssdv.cf32")
    full.astype(np.complex64).tofile(path)
    lines, chars = _run(_build(tmp_path), path, fs, baud, 8, 2, factor)
    want_ev, _, want_chars = tg._reference_transcript(oracle_kind, full, fs, baud, factor, chunk)
    assert chars == want_chars
    got = [l.split() for l in lines if l.startswith("SSDV ")]
    # oracle event: (call, callsign, image_id, packet_id, width, height, set_size, crc32 of the set's packets)
    assert [(g[1], int(g[2]), int(g[3]), int(g[4])) for g in got] == [(e[1], e[2], 256 * e[6], e[7]) for e in want_ev]
    assert len(want_ev) >= 4 and "NSSDV %d" % len(want_ev) in lines
    assert all(g[5] == "hdr=%d" % (4 + 8 + len(g[1])) for g in got)
