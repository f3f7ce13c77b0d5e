"""Multi-GPU plumbing: channels are block-partitioned over ranks (one process per GPU) and never
exchange signal data; only fixed-size result records (decoded characters, sentences, AFC scalars:
hbd_result_record, include/habdec_b200.h) travel, once per batch, to rank 0 -- inside the library over
NCCL (csrc/dist.cu: hbd_dist_init / hbd_gather_results), or through torch.distributed with
gather_records() below (gloo in the CPU tests).
The one exception is the wideband channeliser (BASELINE configs[4]): every rank needs the whole
capture, so the rank that owns the receiver broadcasts each block once (20 MS/s x 8 B = 160 MB/s)
and every rank cuts its own slice of frequency-offset channels out of it."""
from __future__ import annotations

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


def gather_records(records, world: int, rank: int, device):
    """hbd_result_record blocks (uint8 [n_local, 256], the output of BatchDecoder.pack_results / api.make_records) of
    all ranks, in rank order, through torch.distributed (gloo in the CPU tests; any backend).  Ranks may own different
    numbers of channels.  Rank 0 gets uint8 [n_total, 256] (to be fed to a ResultSink), the others None.
    On GPUs the library moves the records itself over NCCL (hbd_dist_init / hbd_gather_results, csrc/dist.cu); this is
    the same gather for transports the library does not know."""
    import numpy as np
    import torch
    import torch.distributed as dist
    recs = np.ascontiguousarray(records, dtype=np.uint8).reshape(-1, records.shape[-1])
    if world == 1:
        return recs
    n = torch.tensor([recs.shape[0]], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    buf = torch.zeros((max(counts), recs.shape[1]), dtype=torch.uint8, device=device)
    buf[:recs.shape[0]] = torch.from_numpy(recs).to(device)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    if rank != 0:
        return None
    return np.concatenate([bufs[r][:counts[r]].cpu().numpy() for r in range(world)], axis=0)
