"""Multi-GPU plumbing: channels are block-partitioned over ranks (one process per GPU) and never
exchange signal data; only decoded sentences and per-channel AFC/stat records travel, once per
batch, to rank 0 (torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests).
The one exception is the wideband channeliser (BASELINE configs[4]): every rank needs the whole
capture, so the rank that owns the receiver broadcasts each block once (20 MS/s x 8 B = 160 MB/s)
and every rank cuts its own slice of frequency-offset channels out of it."""
from __future__ import annotations

import json

import numpy as np


def shard(n_channels: int, world: int, rank: int) -> range:
    """Block partition: rank r owns [r*C/G, (r+1)*C/G) (remainder spread over the first ranks)."""
    base, rem = divmod(n_channels, world)
    lo = rank * base + min(rank, rem)
    return range(lo, lo + base + (1 if rank < rem else 0))


def broadcast_capture(block, world: int, src: int = 0):
    """One block of the wideband capture (float32 tensor [n, 2], interleaved cf32) from rank `src` to all ranks, in
    place.  NCCL on GPUs (one broadcast over NVLink/NVSwitch per block), gloo on CPU.  Returns the tensor."""
    if world > 1:
        import torch.distributed as dist
        dist.broadcast(block, src=src)
    return block


def wideband_plan(offsets_hz, world: int, rank: int):
    """Rank `rank`'s slice of the frequency-offset channels of one capture: (first global channel, NCO offsets)."""
    mine = shard(len(offsets_hz), world, rank)
    return mine.start, [float(offsets_hz[c]) for c in mine]


def collect_local_results(dec, ch0: int) -> dict:
    """Drain one BatchDecoder: {global_channel: {"sentences": [...], "last": str, "afc": [corr, shift, nf, nv, pl, pr]}}."""
    stats = dec.stats_all()
    out = {}
    for c in range(dec.n_channels):
        out[ch0 + c] = {"sentences": [s.decode("latin1") for s in dec.poll_sentences(c)],
                        "last": dec.getLastSentence(c).decode("latin1"),
                        "afc": [float(x) for x in stats[c]]}
    return out


def gather_to_rank0(local: dict, world: int, rank: int, device) -> dict | None:
    """All ranks call this; rank 0 gets the merged dict, the others None.  Two collectives: sizes, then one
    padded uint8 all_gather (fixed-size records are latency bound, so one message per rank per batch)."""
    if world == 1:
        return {int(k): v for k, v in local.items()}
    import torch
    import torch.distributed as dist
    payload = np.frombuffer(json.dumps(local).encode(), dtype=np.uint8)
    size = torch.tensor([payload.size], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(size) for _ in range(world)]
    dist.all_gather(sizes, size)
    max_len = int(max(int(s.item()) for s in sizes))
    buf = torch.zeros(max_len, dtype=torch.uint8, device=device)
    buf[:payload.size] = torch.from_numpy(payload.copy()).to(device)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    if rank != 0:
        return None
    merged = {}
    for r in range(world):
        raw = bytes(bufs[r][:int(sizes[r].item())].cpu().numpy())
        for k, v in json.loads(raw.decode()).items():
            merged[int(k)] = v
    return merged
