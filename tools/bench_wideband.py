"""Diagnostic (not the bench contract): BASELINE configs[4] -- one 20 MS/s capture channelised into 1024 frequency-offset
RTTY channels (dec=8) through the per-channel NCO (K0) + K1..K4.  Prints channel-samples/s (1024 x capture rate) and the
real-time factor.  The capture is synthetic noise resident in HBM; parity of this path is in tests/test_gpu_nco.py.
usage: python tools/bench_wideband.py [--channels 1024] [--steps 100]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from habdec_b200 import api, dist as hdist

ap = argparse.ArgumentParser()
ap.add_argument("--channels", type=int, default=1024)
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--chunk", type=int, default=65536)
ap.add_argument("--fs", type=float, default=20e6)
ap.add_argument("--no-offsets", action="store_true", help="diagnostic: all NCOs at 0 Hz (plain copy semantics, shared row)")
ap.add_argument("--matrix", action="store_true", help="diagnostic: per-channel device matrix (pushSamplesDevice) with NCOs instead of one shared row")
a = ap.parse_args()
world, rank, local_rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:     # torchrun: the channels are block-partitioned, rank 0 owns the capture and broadcasts it (NCCL)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
n_slices = 64
cap = torch.randn((n_slices * a.chunk, 2), dtype=torch.float32, device=dev) * 0.7 if rank == 0 else torch.zeros((n_slices * a.chunk, 2), dtype=torch.float32, device=dev)
hdist.broadcast_capture(cap, world, src=0)
total_channels = a.channels
ch0, my_offsets = hdist.wideband_plan([(c - total_channels / 2) * 15e3 for c in range(total_channels)], world, rank)
a.channels = len(my_offsets)
dec = api.BatchDecoder(a.channels, device=local_rank, baud=300.0, rtty_bits=8, rtty_stops=2.0, dec_factor=256)
stream = torch.cuda.current_stream()
dec.set_stream(stream.cuda_stream)
if not a.no_offsets:
    for c in range(a.channels):
        dec.set_nco(my_offsets[c], c)                        # 15 kHz raster over the capture
mat = torch.randn((a.channels, 2 * a.chunk, 2), dtype=torch.float32, device=dev) if a.matrix else None


def step(i):
    if a.matrix:
        dec.pushSamplesDevice(mat.data_ptr() + (i % 2) * a.chunk * 8, a.chunk, 2 * a.chunk, a.fs)
    else:
        dec.pushWidebandDevice(cap.data_ptr() + (i % n_slices) * a.chunk * 8, a.chunk, a.fs)
    dec.process_async()


for i in range(5):
    step(i)
dec.collect()
torch.cuda.synchronize()
dec.set_kernel_timing(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for i in range(a.steps):
    step(5 + i)
    if (i + 1) % 16 == 0:
        dec.collect_ready(8)
dec.collect()
e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
if world > 1:
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
k1_ms, k1_n = dec.kernel_timing(0)
rest_ms, rest_n = dec.kernel_timing(1)
chs = total_channels * a.chunk / (ms * 1e-3)
if world > 1:
    dist.barrier()
if rank == 0:
  print(json.dumps({"workload": "wideband %.0f MS/s -> %d NCO channels over %d GPU(s), dec=8, chunk %d" % (a.fs / 1e6, total_channels, world, a.chunk),
                  "ms_per_step": ms, "k1_ms": k1_ms / max(k1_n, 1), "rest_ms": rest_ms / max(rest_n, 1), "channel_MSamples_per_s": chs / 1e6, "capture_MSamples_per_s": a.chunk / (ms * 1e-3) / 1e6,
                  "realtime_factor": a.chunk / (ms * 1e-3) / a.fs}))
if world > 1:
    dist.destroy_process_group()
