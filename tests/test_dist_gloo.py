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


def _streams(c):
    """Deterministic per-channel result streams: characters (some channels far more than one record holds), sentences."""
    import random
    rng = random.Random(1000 + c)
    chars = "".join(rng.choice("ABC$*,123 \n") for _ in range(rng.choice([0, 3, 40, 300, 900]))).encode()
    sents = "".join("C%04d,%d,%s*%04X\n" % (c, k, "x" * rng.randint(0, 200), rng.randint(0, 65535)) for k in range(rng.choice([0, 1, 2, 7]))).encode()
    return chars, sents, [float(c), 0.5 * c, -60.0, 12.0, c, c + 1]


def _rounds(channels, ch0):
    """Pack the streams of `channels` into rounds of records (hbd_record_set) until nothing is left."""
    from habdec_b200 import api
    chars, sents, stats = map(list, zip(*[_streams(c) for c in channels])) if channels else ([], [], [])
    out = []
    while True:
        recs, cu, su = api.make_records(ch0, chars, sents, stats)
        out.append(recs)
        chars = [x[k:] for x, k in zip(chars, cu)]
        sents = [x[k:] for x, k in zip(sents, su)]
        if not any(chars) and not any(sents):
            return out


def _worker(rank, world, port, n_channels, q):
    from habdec_b200 import api
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = hdist.shard(n_channels, world, rank)
    rounds = _rounds(list(mine), mine.start)
    n_rounds = torch.tensor([len(rounds)])
    dist.all_reduce(n_rounds, op=dist.ReduceOp.MAX)               # every rank gathers the same number of times
    sink = api.ResultSink(n_channels) if rank == 0 else None
    empty = api.make_records(mine.start, [b""] * len(mine), [b""] * len(mine), [_streams(c)[2] for c in mine])[0]
    for k in range(int(n_rounds.item())):
        allr = hdist.gather_records(rounds[k] if k < len(rounds) else empty, world, rank, torch.device("cpu"))
        if rank == 0:
            assert allr.shape == (n_channels, api.RECORD_BYTES)
            assert sink.feed(allr) == 0
        else:
            assert allr is None
    if rank == 0:
        q.put((sink.hash(), sink.totals(), [(sink.poll_chars(c), sink.poll_sentences(c), list(sink.stats(c))) for c in range(n_channels)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_channels", [(2, 10), (2, 7)])
def test_result_records_gather_to_rank0_gloo(world, n_channels):
    """the gathered output is identical however the channels are sharded.  Two gloo ranks pack their block of
    channels into hbd_result_records (several rounds: most streams exceed one record), rank 0 feeds its sink; contents and
    the sharding-invariant hash have to equal what ONE process gets for all channels."""
    from habdec_b200 import api
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_channels, q)) for r in range(world)]
    for p in procs:
        p.start()
    h, totals, content = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    single = api.ResultSink(n_channels)
    for recs in _rounds(list(range(n_channels)), 0):
        assert single.feed(recs) == 0
    assert h == single.hash()
    assert totals["chars"] == single.totals()["chars"] and totals["sentences"] == single.totals()["sentences"]
    for c in range(n_channels):
        chars, sents, stats = _streams(c)
        assert content[c][0] == chars and content[c][1] == [x for x in sents.split(b"\n") if x]
        assert content[c][2] == pytest.approx(stats)
    # a different stream anywhere changes the hash
    other = api.ResultSink(n_channels)
    recs = _rounds(list(range(n_channels)), 0)
    recs[0][n_channels - 1, 40] ^= 1
    for r in recs:
        other.feed(r)
    assert other.hash() != h or totals["chars"] == 0


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
