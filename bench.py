#!/usr/bin/env python
"""bench.py -- aggregate IQ MSamples/s decoded (BASELINE.json metric) on N B200s of one node.

Workload (BASELINE.json configs[3]): 4096 independent RTTY channels in total, fs 2.048 MS/s, 300 baud
8N2, 425 Hz shift, dec=8 (factor 256), low-pass 1500 Hz, block-partitioned over the ranks (no
collective on the signal path; decoded sentences + AFC stats are all-gathered to rank 0 over NCCL
after the timed region's last step).  One "step" = one Decoder::process() over one chunk of every
channel.  Inputs are synthetic (habdec_b200/synth.py ring workload, periodic continuous-phase FSK +
AWGN), resident in HBM for `value`, in pinned host memory for `e2e`.

  python bench.py --gpus 1 --steps 20 --warmup 3            # our arm
  python bench.py --impl reference --steps 3 --warmup 1     # the reference's CPU Decoder on the host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FS = 2.048e6
BAUD = 300.0
FACTOR = 256
TOTAL_CHANNELS = 4096
SNR_DB = -15.0
ALGO_BYTES_PER_SAMPLE_K1 = 8.0 + 8.0 / 64.0   # cf32 read + stage-1 output write
# DRAM bytes per input sample that K1 really moved in the ncu --set full capture of this exact workload
# (profiles/r1c_k1_ncu_full_raw.csv: dram__bytes_read.sum 2.170465 GB + dram__bytes_write.sum 0.046897 GB per 2^28 samples)
NCU_TRAFFIC_BYTES_PER_SAMPLE_K1 = (2.170465e9 + 0.046897e9) / 268435456.0
METRIC = "aggregate IQ MSamples/s decoded (chars bit-exact)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--channels", type=int, default=None, help="total channels over all ranks (default: 4096 per GPU, weak scaling)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 4096 channels per GPU (BASELINE configs[3] as the per-GPU shard); strong: 4096 channels in total")
    ap.add_argument("--chunk", type=int, default=65536, help="complex samples per channel per step")
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target duration of the cpu_baseline leg")
    ap.add_argument("--ssdv", action="store_true", help="diagnostic: run with SSDV packet sync switched on (one more kernel per step)")
    ap.add_argument("--collect-every", type=int, default=16, help="drain finished calls every this many steps")
    ap.add_argument("--collect-lag", type=int, default=8, help="calls left in flight by the periodic drain")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-kernel-timing", action="store_true", help="diagnostic: no CUDA events around K1 (no roofline numbers)")
    ap.add_argument("--ref-passes", type=int, default=8, help="reference arm: ring passes per host thread per step")
    a = ap.parse_args()
    if a.channels is None:
        world = int(os.environ.get("WORLD_SIZE", "1"))   # ranks actually launched (torchrun); --gpus is informational
        a.channels = TOTAL_CHANNELS * (max(world, 1) if a.scaling == "weak" else 1)
    return a


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons sampled (NVML, every 5 ms) while the timed region runs."""

    def __init__(self, index):
        self.index, self.rows, self._stop, self.th = index, [], False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.mx = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
            return
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def _run(self):
        nv = self.nv
        reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        rs, k = 0, 0
        while not self._stop:
            try:
                clk = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                if k % 3 == 0:          # NVML queries can take tens of ms on a busy box: the clock every time, the reasons every third
                    rs = reasons(self.h)
                k += 1
                self.rows.append((clk, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._stop = True
        if self.th:
            self.th.join(timeout=1.0)
        if not getattr(self, "nv", None) or not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        nv = self.nv
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        reasons = sorted({name for _, rs in self.rows for name, b in bits.items() if rs & b})
        return {"sm_mhz": float(np.median([c for c, _ in self.rows])), "sm_max_mhz": float(self.mx), "reasons": reasons, "samples": len(self.rows)}


def workload_config(args, world, impl):
    return {"workload": "BASELINE configs[3]: %d RTTY channels per GPU (%d in total, %s scaling), 2.048 MS/s cf32, 300 baud 8N2, 425 Hz shift, "
                        "dec=8 (factor 256), lowpass 1500 Hz, chunk %d samples/channel/step" % (args.channels // world, args.channels, args.scaling, args.chunk),
            "channels_total": args.channels, "channels_per_gpu": args.channels // world, "chunk": args.chunk,
            "snr_db_fullband": SNR_DB, "l2_policy": "inputs larger than L2: every step reads a different %.2f GiB slice of a ring resident in HBM"
            % (args.channels // world * args.chunk * 8 / 2**30), "parallelism": "channels block-partitioned, %d rank(s)" % world}


# ---------------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU Decoder (oracle/_ref when it was built, else the restatement) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from habdec_b200 import synth
    from oracle import pyoracle as po
    kind = "ref" if po.available("ref") else "orc"
    cores = os.cpu_count() or 1
    cfg = po.make_config(baud=BAUD, rtty_bits=8, rtty_stops=2.0, lowpass_bw=1500.0, lowpass_trans=0.025, dec_factor=FACTOR, record=False)
    L = synth.ring_length(FS, BAUD)
    iq = np.stack([synth.ring_iq_numpy(c, FS, BAUD, snr_db=SNR_DB) for c in range(cores)])
    # a step = every host thread decodes `ref_passes` ring passes (L samples each) of its own channel, pushed in
    # `chunk` pieces exactly like DECODER_THREAD does (code/websocketServer/main.cpp:235-245)
    P = max(1, args.ref_passes)
    if args.warmup > 0:
        po.bench(kind, cfg, iq, cores, FS, chunk=args.chunk, reps=P * args.warmup)
    t0 = time.time()
    secs, chars = po.bench(kind, cfg, iq, cores, FS, chunk=args.chunk, reps=P * args.steps)
    wall = time.time() - t0
    samples = float(cores) * L * P * args.steps
    value = samples / secs / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "MSamples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": secs * 1e3 / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, max(int(os.environ.get("WORLD_SIZE", "1")), 1), "reference"),
            "cpu_baseline": {"value": value, "unit": "MSamples/s", "cores": cores, "kind": "reference" if kind == "ref" else "port",
                             "sample": "%d host threads x %d steps x %d ring passes of %d samples each (one Decoder per thread), wall %.1f s, %d chars decoded"
                             % (cores, args.steps, P, L, wall, chars)},
            "e2e": {"value": value, "unit": "MSamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from habdec_b200 import api, synth, dist as hdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: habdec_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert args.channels % world == 0
    C = args.channels // world
    ch0 = rank * C
    L = synth.ring_length(FS, BAUD)
    assert L % args.chunk == 0, "chunk must divide the ring length %d" % L
    slices = L // args.chunk

    ring = synth.ring_iq_torch(ch0, C, dev, FS, BAUD, snr_db=SNR_DB)          # [C, L, 2] float32, HBM resident
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()
    dec = api.BatchDecoder(C, device=local_rank, baud=BAUD, rtty_bits=8, rtty_stops=2.0, lowpass_bw=1500.0, lowpass_trans=0.025,
                           dec_factor=FACTOR)
    dec.set_stream(stream.cuda_stream)
    if args.ssdv:
        dec.set_ssdv(True)
    base_ptr = ring.data_ptr()

    def step(i):
        dec.pushSamplesDevice(base_ptr + (i % slices) * args.chunk * 8, args.chunk, L, FS)
        dec.process_async()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_done = 0
    for _ in range(args.warmup):
        step(n_done); n_done += 1
    dec.collect()
    launches0 = dec.kernel_launches()
    dec.set_kernel_timing(not args.no_kernel_timing)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    t_issue = 0.0
    for k in range(args.steps):
        t_h = time.perf_counter()
        step(n_done); n_done += 1
        t_issue += time.perf_counter() - t_h
        if (k + 1) % args.collect_every == 0:
            dec.collect_ready(args.collect_lag)   # drain finished calls; the newest few stay in flight so the GPU never idles
    host_issue_ms = t_issue * 1e3 / max(args.steps, 1)   # host time to enqueue one step, collects excluded (diagnostic)
    t_fc = time.perf_counter()
    dec.collect()                       # results drained (D2H + sentence layer) inside the timed region
    final_collect_ms = (time.perf_counter() - t_fc) * 1e3
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    k1_ms, k1_cnt = dec.kernel_timing(0)
    rest_ms, rest_cnt = dec.kernel_timing(1)
    gaps = {}
    for name, w in (("k1_end_to_next_k1_start_ms", 2), ("k1_end_to_tail_start_ms", 3), ("tail_end_to_k1_plus2_start_ms", 4)):
        g_ms, g_cnt = dec.kernel_timing(w)
        gaps[name] = g_ms / g_cnt if g_cnt else None
    dec.set_kernel_timing(False)
    launches = dec.kernel_launches() - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())

    # results: sentences per channel + AFC stats gathered to rank 0 (NCCL all_gather of packed records)
    local = hdist.collect_local_results(dec, ch0)
    gathered = hdist.gather_to_rank0(local, world, rank, dev)
    total_samples = float(args.channels) * args.chunk * args.steps
    value = total_samples / (ms_max * 1e-3) / 1e6

    # ---- e2e: same metric through the C ABI with HOST buffers (H2D of every step's input + D2H of results in the timed region)
    e2e = None
    if not args.no_e2e:
        n_host = min(3, slices)
        host = [torch.empty((C, args.chunk, 2), dtype=torch.float32, pin_memory=True) for _ in range(n_host)]
        for i in range(n_host):
            host[i].copy_(ring[:, i * args.chunk:(i + 1) * args.chunk, :])
        torch.cuda.synchronize()
        dec2 = api.BatchDecoder(C, device=local_rank, baud=BAUD, rtty_bits=8, rtty_stops=2.0, dec_factor=FACTOR)
        dec2.set_stream(stream.cuda_stream)

        def e2e_step(i):
            # host buffer -> hbd_push_samples_batch (H2D inside, the caller owns the buffer again on return) -> kernels;
            # the results of the previous step are read back (D2H + sentence layer) while this step's kernels run
            h = host[i % n_host]
            dec2._chk(dec2._lib.hbd_push_samples_batch(dec2._h, h.data_ptr(), args.chunk, args.chunk, FS))
            dec2.process_async()
            dec2.collect_ready(1)
        e2e_step(0)
        dec2.collect()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.e2e_steps):
            e2e_step(i + 1)
        dec2.collect()                  # the last step's results
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        d2h = sum(len(dec2.poll_raw_chars(c)) for c in range(C)) / max(args.e2e_steps + 1, 1) + 4 * C
        e2e = {"value": float(args.channels) * args.chunk * args.e2e_steps / float(tt.item()) / 1e6, "unit": "MSamples/s",
               "h2d_bytes_per_step": int(C * args.chunk * 8), "d2h_bytes_per_step": int(d2h), "steps": args.e2e_steps,
               "note": "hbd_push_samples_batch from pinned host memory + hbd_process_async + hbd_collect_ready per step, wall clock, max over ranks; PCIe bound (pinned H2D on this pool: 55.5 GB/s, tools/micro/h2d_bw.py)"}
        dec2.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = load_peaks()
    # K1 is launched once per channel group per step: algorithmic bytes of one launch = step bytes / launches per step
    k1_bytes = ALGO_BYTES_PER_SAMPLE_K1 * C * args.chunk * args.steps / max(k1_cnt, 1)
    k1_avg_ms = k1_ms / max(k1_cnt, 1)
    achieved = k1_bytes / (k1_avg_ms * 1e-3) / 1e9 if k1_cnt else None
    exp_sent = args.steps * args.chunk // L
    got_sent = [len(v["sentences"]) for v in gathered.values()] if gathered else []
    line = {"metric": METRIC, "value": value, "unit": "MSamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args, world, "ours"),
            "roofline": {"bound": "hbm", "kernel": "decim1_kernel<64,348> (K1, stage-1 FIR decimator)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": (NCU_TRAFFIC_BYTES_PER_SAMPLE_K1 * C * args.chunk * args.steps / max(k1_cnt, 1)) if (C == 4096 and args.chunk == 65536) else None,
                         "traffic_source": "ncu --set full capture of this workload, profiles/r1c_k1_ncu_full_raw.csv (bytes per launch)",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": k1_bytes, "avg_launch_ms": k1_avg_ms, "launches_timed": k1_cnt,
                         "k1_share_of_step": (k1_ms / ms) if ms else None, "rest_of_step_ms": rest_ms / max(rest_cnt, 1), "pipeline_gaps": gaps},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "host_issue_ms_per_step": host_issue_ms, "final_collect_ms": final_collect_ms,
            "results": {"channels_gathered": len(gathered) if gathered else 0, "sentences_expected_per_channel_approx": exp_sent,
                        "sentences_min": min(got_sent) if got_sent else None, "sentences_max": max(got_sent) if got_sent else None}}

    if world == 1 and not args.no_cpu_baseline:
        from oracle import pyoracle as po
        kind = "ref" if po.available("ref") else "orc"
        cores = os.cpu_count() or 1
        cfg = po.make_config(baud=BAUD, dec_factor=FACTOR, record=False)
        iq = np.stack([synth.ring_iq_numpy(c, FS, BAUD, snr_db=SNR_DB) for c in range(cores)])
        secs1, _ = po.bench(kind, cfg, iq, cores, FS, chunk=args.chunk, reps=1)
        reps = max(1, int(args.cpu_seconds / max(secs1, 1e-3)))
        secs, chars = po.bench(kind, cfg, iq, cores, FS, chunk=args.chunk, reps=reps)
        line["cpu_baseline"] = {"value": cores * float(L) * reps / secs / 1e6, "unit": "MSamples/s", "cores": cores,
                                "kind": "reference" if kind == "ref" else "port",
                                "sample": "%d threads x %d ring passes of %d samples (one reference Decoder per thread), %.1f s" % (cores, reps, L, secs)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
