"""CPU tier: the N>1 result gather (habdec_b200/dist.py) on a world of 2 gloo ranks."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from habdec_b200 import dist as hdist


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_channels, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = hdist.shard(n_channels, world, rank)
    local = {c: {"sentences": ["C%04d,%d*ABCD" % (c, k) for k in range(c % 3)], "last": "C%04d" % c, "afc": [float(c), 0.5, -60.0, 12.0, c, c + 1]}
             for c in mine}
    merged = hdist.gather_to_rank0(local, world, rank, torch.device("cpu"))
    if rank == 0:
        q.put(merged)
    else:
        assert merged is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_channels", [(2, 10), (2, 7)])
def test_gather_to_rank0_gloo(world, n_channels):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_channels, q)) for r in range(world)]
    for p in procs:
        p.start()
    merged = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert sorted(merged) == list(range(n_channels))
    for c, v in merged.items():
        assert v["last"] == "C%04d" % c and len(v["sentences"]) == c % 3 and v["afc"][4] == c


def _wideband_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(7)
    block = torch.randn((4096, 2), generator=g) if rank == 0 else torch.zeros((4096, 2))
    hdist.broadcast_capture(block, world, src=0)
    offsets = [(c - 5) * 15e3 for c in range(11)]
    ch0, mine = hdist.wideband_plan(offsets, world, rank)
    q.put((rank, float(block.double().sum()), ch0, mine))
    dist.barrier()
    dist.destroy_process_group()


def test_wideband_capture_broadcast_and_channel_plan_gloo():
    """configs[4] on N ranks: every rank ends up with the identical capture block and a disjoint slice of the offsets."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_wideband_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert got[0][1] == got[1][1] != 0.0                       # same samples everywhere
    assert got[0][2] == 0 and got[1][2] == len(got[0][3])      # block partition
    assert got[0][3] + got[1][3] == [(c - 5) * 15e3 for c in range(11)]
