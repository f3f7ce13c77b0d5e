"""CPU tier: host logic of the product that needs no GPU -- ABI surface, low-pass design, sentence layer, sharding."""
import os
import random
import re
import sys

import numpy as np
import pytest

from habdec_b200 import api, dist as hdist, synth
from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = api.load()
    header = open(os.path.join(ROOT, "include", "habdec_b200.h")).read()
    declared = set(re.findall(r"\b(hbd_[a-z0-9_]+)\s*\(", header))
    declared -= {"hbd_sentence_cb", "hbd_chars_cb"}
    assert len(declared) >= 50
    for name in sorted(declared):
        assert hasattr(lib, name), "libhabdec_b200.so does not export %s" % name
        assert name in api.SIGNATURES, "python binding misses %s" % name
    assert set(api.SIGNATURES) <= declared


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.HbdError):
        api.BatchDecoder(1)


def test_crc_known_answers():
    assert api.crc16(b"CALL,1,12:00:00,52.1234,21.5678,1000") == b"2BEE"
    for s in (b"", b"A", b"CH0011,0,12:00:00,52.1234,21.5678,1000", bytes(range(32, 127))):
        assert api.crc16(s) == po.port_crc16(s) == synth.crc16_ccitt(s).encode()


@pytest.mark.parametrize("bw,fs_dec,trans,n", [(1500.0, 8000.0, 0.025, 256), (1500.0, 9765.625, 0.025, 1024), (1500.0, 78125.0, 0.025, 256),
                                               (800.0, 8000.0, 0.05, 256), (1500.0, 8000.0, 0.01, 256), (1500.0, 8000.0, 0.0, 256)])
def test_lowpass_design_bit_exact(bw, fs_dec, trans, n):
    taps = api.design_lowpass(np.float32(bw / fs_dec), trans, n)
    # drive the port's filter through one call so it designs with the same inputs, then read its taps back
    d = po.PortDecoder(po.make_config(baud=300.0, lowpass_bw=bw, lowpass_trans=trans, dec_factor=2))
    d.push_process(np.zeros(2 * n, dtype=np.complex64), 2 * fs_dec)
    want = d.stage(po.STAGE_LPTAPS)
    assert taps.shape == want.shape and np.array_equal(taps.view(np.uint32), want.view(np.uint32))
    if (bw, fs_dec, trans) == (1500.0, 8000.0, 0.025):
        assert len(taps) == 161 and abs(taps[80] - 0.119366) < 1e-6 and abs(taps.sum() - 1.0) < 1e-6


def test_sentence_matcher_equals_std_regex_on_fuzz():
    rng = random.Random(7)
    alphabet = "$$$**,,,  --__AB12ab#\n"
    checked = matched = 0
    for it in range(6000):
        n = rng.randint(0, 60)
        s = "".join(rng.choice(alphabet) for _ in range(n))
        if it % 3 == 0:   # plant something sentence-like
            body = "".join(rng.choice("AB12,-_ ") for _ in range(rng.randint(1, 8)))
            s = s[:n // 2] + "$" * rng.randint(1, 3) + body + "," + "".join(rng.choice("12,.ab$") for _ in range(rng.randint(0, 6))) + \
                rng.choice("*$") + "".join(rng.choice("0123ABCD_#") for _ in range(rng.randint(0, 6))) + s[n // 2:]
        b = s.encode()
        got, want = api.extract_sentence(b), po.port_extract_sentence(b)
        assert got == want, repr(s)
        checked += 1
        matched += want is not None
    assert matched > 300 and checked == 6000


def _text_layer_model(chunks):
    """Decoder.h:572-613,635-636 with the restatement's std::regex extraction (the reference's extractSentence)."""
    stream, last, sentences = b"", b"", []
    for raw in chunks:
        if not raw:
            continue
        stream += bytes(c for c in raw if (32 <= c < 127) or c == 10)
        if len(stream) > 20:
            while True:
                m = po.port_extract_sentence(stream)
                if m is None:
                    break
                cs, data, crc, rest = m
                stream = stream[rest:].replace(b"\n", b" ")
                last = cs + b"," + data + b"*" + crc
                if crc == po.port_crc16(cs + b"," + data):
                    sentences.append(last)
        if len(stream) > 1000:
            k = stream.rfind(b"$")
            stream = stream[k:] if k >= 0 else b""
    return sentences, last, stream


def test_text_layer_incremental_scan_equals_full_scan_on_fuzz():
    """TextChannel::feed skips the sentence scan unless the new characters can complete a CRC group; the result has to be
    what rescanning the whole stream after every push gives (the reference's behaviour), for any chunking."""
    rng = random.Random(11)
    n_sent = 0
    for it in range(300):
        parts = []
        for k in range(rng.randint(1, 6)):
            body = "C%d,%d" % (rng.randint(0, 99), rng.randint(0, 999)) + "".join(rng.choice(",.12ab$ ") for _ in range(rng.randint(0, 5)))
            crc = api.crc16(body.encode()).decode() if rng.random() < 0.7 else "".join(rng.choice("0123ABC_#") for _ in range(rng.randint(0, 5)))
            parts.append("".join(rng.choice("$*,ab1 \n\x00\xff#") for _ in range(rng.randint(0, 12))) + "$" * rng.randint(1, 3) + body + rng.choice("**$") + crc + rng.choice(["\n", "", "x"]))
        text = "".join(parts).encode("latin1")
        chunks, o = [], 0
        while o < len(text):
            k = rng.choice([0, 1, 1, 2, 3, 5, 8, 30])
            chunks.append(text[o:o + k]); o += k
        got = api.text_replay(chunks)
        want = _text_layer_model(chunks)
        assert (got[0], got[1], got[2]) == (want[0], want[1], want[2]), repr(text)
        n_sent += len(want[0])
    assert n_sent > 300


def test_sentence_matcher_real_sentences():
    s = synth.make_sentence(12, 3)
    stream = ("noise" + s + s[:20]).encode()
    cs, data, crc, rest = api.extract_sentence(stream)
    assert cs == b"CH0012" and crc == api.crc16(cs + b"," + data)
    assert stream[rest:rest + 1] == crc[-1:]          # the reference keeps the last CRC character in the stream


def test_block_partition():
    for n, w in ((4096, 8), (4096, 1), (10, 4), (3, 8)):
        got = [c for r in range(w) for c in hdist.shard(n, w, r)]
        assert got == list(range(n))
    assert len(hdist.shard(4096, 8, 3)) == 512


def test_ring_workload_properties():
    L = synth.ring_length(2.048e6, 300.0)
    assert L == 1638400 and L % 65536 == 0
    s = synth.ring_sentence(77)
    assert len(s) == 21 and s.startswith("$$C0077,")      # > 20 characters: the reference scans its stream once per pass (Decoder.h:591)
    assert len(synth.ring_bits(77)) == synth.RING_BITS


# ---- the step in front of the decoder: cf32 file IQ source (include/habdec_b200/IQSource.hpp) ---------------
def _build_cpp(tmp_path, src, name, extra=()):
    import subprocess
    exe = str(tmp_path / name)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle"), src, "-o", exe,
                    "-lpthread", *extra], check=True)
    return exe


def test_iqsource_file_matches_reference_transcript(tmp_path):
    """Same scripted session (options, EOF, loop, short reads, stop quirk) as the reference's IQSource_File<float>:
    golden transcript generated from the reference itself (tests/golden/make_iqsource_golden.py)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_iqsource_golden as g
    exe = _build_cpp(tmp_path, os.path.join(ROOT, "tests", "cpp", "iqsource_ours.cpp"), "iqsource_ours")
    path = str(tmp_path / "cap.cf32")
    g.make_file(path)
    ours = g.transcript(exe, path)
    golden = open(os.path.join(ROOT, "tests", "golden", "iqsource_transcript.txt")).read()
    assert ours == golden
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "iqsource_ref")
    if os.path.exists(ref_bin):
        assert g.transcript(ref_bin, path) == golden


def test_cpp_facade_builds_links_and_fails_loudly_without_a_gpu(tmp_path):
    """include/habdec_b200/Decoder.hpp + IQSource.hpp compile as C++17 against the C ABI library; on a box without a CUDA
    device the Decoder constructor throws (no CPU fallback, nothing is decoded by other means)."""
    import subprocess
    import numpy as np
    lib_dir = os.path.join(ROOT, "habdec_b200")
    exe = _build_cpp(tmp_path, os.path.join(ROOT, "tests", "cpp", "decoder_thread.cpp"), "decoder_thread",
                     extra=("-L", lib_dir, "-lhabdec_b200", "-Wl,-rpath," + lib_dir))
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present: the run itself is covered by tests/test_gpu_cpp_facade.py")
    except ImportError:
        pass
    path = str(tmp_path / "z.cf32")
    np.zeros(65536, dtype=np.complex64).tofile(path)
    r = subprocess.run([exe, path, "2048000", "300", "8", "2", "256"], capture_output=True)
    assert r.returncode != 0
    assert b"no usable CUDA device" in r.stderr and b"CHARS" not in r.stdout


def test_range_pool_runs_every_part_once_on_a_stable_thread(tmp_path):
    """habdec_b200/csrc/range_pool.h (the drain's persistent worker threads): tests/cpp/range_pool_test.cpp under the thread
    sanitizer where gcc has it, plain otherwise."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "tests", "cpp", "range_pool_test.cpp")
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    exe = str(tmp_path / "range_pool_test")
    built = False
    for extra in (["-fsanitize=thread"], []):
        r = subprocess.run([gxx, "-std=c++17", "-O2", "-pthread", *extra, src, "-o", exe], capture_output=True, text=True)
        if r.returncode == 0:
            built = True
            break
    assert built, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr


def _slicer_masks_full(v, R):
    """slicer_build_masks (slicer_dev.cuh) for every position: bit q of A = the window means left and right of q differ in
    sign, windows clipped to the buffer like window_sums; positions below R are not evaluated."""
    n = len(v)
    words = [0] * ((n + 31) // 32)
    for q in range(R, n):
        sl = float(np.sum(v[max(q - R, 0):q], dtype=np.float64))
        sr = float(np.sum(v[q:min(q + R, n)], dtype=np.float64))
        if (sl > 0) - (sl < 0) != (sr > 0) - (sr < 0):
            words[q >> 5] |= 1 << (q & 31)
    return words


def test_slicer_mask_cache_model_equals_full_rebuild():
    """The bookkeeping of the tail kernel's mask cache (tail.cu), modelled word for word on the host: whole cached words are
    reused, the build restarts at the word holding the first unknown position, and after the slicer erased `erase` samples the
    words move down by a funnel shift with `keep - R` positions staying valid.  Over random append / erase sequences every
    position the flip search may consult (q >= R) has the bit a from-scratch build gives."""
    rng = random.Random(11)
    nrng = np.random.default_rng(11)
    for R in (4, 7, 33, 65):
        v = np.zeros(0, dtype=np.float32)
        cached, valid = [], 0                       # HBM words and ChanState::mask_valid
        for step in range(60):
            v = np.concatenate([v, nrng.standard_normal(rng.choice([32, 256, 257, 300])).astype(np.float32)])
            n = len(v)
            full = _slicer_masks_full(v, R)
            # load: whole words below `valid` (= pending length before this append, minus R)
            m_words = valid >> 5
            work = list(cached[:m_words]) + [0] * (len(full) - m_words)
            for w in range(m_words, len(full)):      # rebuild from the first unknown word on
                work[w] = full[w]
            for q in range(R, n):
                assert (work[q >> 5] >> (q & 31)) & 1 == (full[q >> 5] >> (q & 31)) & 1, (R, step, q)
            # the slicer erases up to some flip point (or nothing)
            erase = rng.choice([0, 0, rng.randrange(0, n), max(0, n - R - rng.randrange(0, 40))])
            keep = n - erase
            nv = keep - R
            if nv > 0:
                nvw, nw, ws, bs = (nv + 31) >> 5, (n + 31) >> 5, erase >> 5, erase & 31
                new = list(cached) + [0] * max(0, nvw - len(cached))
                for w in range(m_words if erase == 0 else 0, nvw):
                    lo = work[w + ws]
                    hi = work[w + ws + 1] if w + ws + 1 < nw else 0
                    new[w] = ((lo | (hi << 32)) >> bs) & 0xffffffff
                cached, valid = new, nv
            else:
                valid = 0
            v = v[erase:]
