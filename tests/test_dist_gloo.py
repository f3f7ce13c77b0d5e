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
