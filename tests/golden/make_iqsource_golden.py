"""Regenerate tests/golden/iqsource_transcript.txt from the reference's own IQSource_File<float>
(oracle/_ref/iqsource_ref, built by `make -C oracle ref` where /root/reference exists)."""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def make_file(path):
    rng = np.random.default_rng(11)
    x = (rng.standard_normal(2500) + 1j * rng.standard_normal(2500)).astype(np.complex64)
    x.tofile(path)


def transcript(binary, path):
    out = subprocess.run([binary, path], capture_output=True, text=True, check=True).stdout
    return out.replace(path, "<FILE>")


if __name__ == "__main__":
    p = "/tmp/iqsource_golden.cf32"
    make_file(p)
    t = transcript(os.path.join(ROOT, "oracle", "_ref", "iqsource_ref"), p)
    open(os.path.join(HERE, "iqsource_transcript.txt"), "w").write(t)
    sys.stdout.write(t)
