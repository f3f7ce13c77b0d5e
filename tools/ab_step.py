"""A/B of library builds on the same box: steady-state step time, K1 and rest-of-step durations (CUDA events) for the
library named by HBD_LIB (default: the in-tree build).  usage: HBD_LIB=... python tools/ab_step.py [steps] [channels]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from habdec_b200 import api, synth
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n_ch = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
fs, chunk = 2.048e6, 65536
L = synth.ring_length(fs, 300.0)
ring = synth.ring_iq_torch(0, n_ch, torch.device("cuda", 0), fs, 300.0)
dec = api.BatchDecoder(n_ch, dec_factor=256)
stream = torch.cuda.current_stream()
dec.set_stream(stream.cuda_stream)
def run(n, timing):
    dec.set_kernel_timing(timing)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(n):
        dec.pushSamplesDevice(ring.data_ptr() + (i % (L // chunk)) * chunk * 8, chunk, L, fs); dec.process_async()
        if (i + 1) % 4 == 0: dec.collect_ready(3)
    dec.collect(); e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
run(30, False)
ms_plain = run(steps, False)
ms_timed = run(steps, 2)
k1, n1 = dec.kernel_timing(0); rest, n2 = dec.kernel_timing(1)
try:
    dr, n3 = dec.kernel_timing(5)
except Exception:
    dr, n3 = 0.0, 0
print("%-28s ch %d: step %.4f ms (no events) / %.4f ms (events), K1 %.4f ms, rest %.4f ms, host replay %.4f ms/call" % (os.path.basename(api.LIB_PATH), n_ch, ms_plain, ms_timed, k1 / max(n1, 1), rest / max(n2, 1), dr / max(n3, 1)))
