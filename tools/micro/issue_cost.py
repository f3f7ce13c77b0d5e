"""Host cost of one hbd_push_samples_device + hbd_process_async (the issue path) and of the drains, for a given channel
count; GPU work is tiny (few channels) or normal.  usage: python tools/micro/issue_cost.py [channels ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from habdec_b200 import api, synth

fs, chunk = 2.048e6, 65536
for n_ch in [int(x) for x in sys.argv[1:]] or [512, 4096]:
    L = synth.ring_length(fs, 300.0)
    ring = synth.ring_iq_torch(0, n_ch, torch.device("cuda", 0), fs, 300.0)
    dec = api.BatchDecoder(n_ch, dec_factor=256)
    dec.set_stream(torch.cuda.current_stream().cuda_stream)
    dec.set_raw_chars(False)
    for timing in (False, True):
        for i in range(30):
            dec.pushSamplesDevice(ring.data_ptr() + (i % 25) * chunk * 8, chunk, L, fs); dec.process_async()
        dec.collect()
        dec.set_kernel_timing(timing)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); ti = 0.0
        for i in range(200):
            a = time.perf_counter()
            dec.pushSamplesDevice(ring.data_ptr() + (i % 25) * chunk * 8, chunk, L, fs); dec.process_async()
            ti += time.perf_counter() - a
            if (i + 1) % 4 == 0:
                dec.collect_ready(3)
        tc0 = time.perf_counter()
        dec.collect()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        print("channels %5d timing %d: issue %.1f us/step, whole loop %.1f us/step (GPU bound if >> issue), final collect %.2f ms"
              % (n_ch, timing, ti / 200 * 1e6, (t1 - t0) / 200 * 1e6, (t1 - tc0) * 1e3))
        dec.set_kernel_timing(False)
    dec.close(); del ring
