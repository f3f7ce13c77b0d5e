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


def test_decoder_thread_program_matches_reference(tmp_path, oracle_kind):
    fs, baud = 2.048e6, 300.0
    iq, _ = synth.channel_iq(9, 3, fs, baud, snr_db=-15.0)
    n = len(iq) // 65536 * 65536
    path = str(tmp_path / "cap.cf32")
    iq[:n].astype(np.complex64).tofile(path)
    exe = str(tmp_path / "decoder_thread")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    lib_dir = os.path.join(ROOT, "habdec_b200")
    subprocess.run([cxx, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "decoder_thread.cpp"),
                    "-L", lib_dir, "-lhabdec_b200", "-Wl,-rpath," + lib_dir, "-lpthread", "-o", exe], check=True)
    out = subprocess.run([exe, path, repr(fs), repr(baud), "8", "2"], check=True, capture_output=True).stdout
    head, _, tail = out.partition(b"CHARS ")
    n_chars, _, rest = tail.partition(b"\n")
    chars = rest[:int(n_chars)]
    assert rest[int(n_chars):].strip() == b"END"
    lines = head.decode().splitlines()

    ref = (po.RefDecoder if oracle_kind == "ref" else po.PortDecoder)(po.make_config(baud=baud)).run(iq[:n], fs)
    assert chars == ref.chars()
    assert [l[5:].encode() for l in lines if l.startswith("SENT ")] == ref.sentences()
    assert "NSENT %d" % len(ref.sentences()) in lines and len(ref.sentences()) >= 2
    assert ("LAST " + ref.last_sentence().decode()) in lines
    assert "RATE 8000 256 4096" in lines
    a = ref.afc()
    assert "PEAKS %d %d" % (a.peak_left, a.peak_right) in lines
