"""Shared SSDV test streams: raw character streams with embedded SSDV packets (clean, corrupted, retransmitted,
resolution changes, truncated) between junk with a high density of 0x55 bytes, cut into per-call chunks."""
import numpy as np

from oracle import pyoracle as po

CALLSIGNS = ["HABDEC", "B200", "SP9", "N0CALL"]


def corrupt(pkt: bytes, n_err: int, rng, lo: int = 1) -> bytes:
    q = bytearray(pkt)
    for i in rng.choice(np.arange(lo, 256), n_err, replace=False):
        q[int(i)] ^= int(rng.integers(1, 256))
    return bytes(q)


def random_packet(rng, callsign=None, image_id=None, packet_id=None, fec=None, **kw) -> bytes:
    fec = bool(rng.integers(0, 4)) if fec is None else fec
    payload = rng.integers(0, 256, 237, dtype=np.uint8).tobytes()
    return po.ssdv_make_packet(callsign or CALLSIGNS[int(rng.integers(0, len(CALLSIGNS)))],
                               int(rng.integers(0, 4)) if image_id is None else image_id,
                               int(rng.integers(0, 6)) if packet_id is None else packet_id, payload, fec=fec, **kw)


def junk(rng, n: int, sync_density: float = 0.02) -> bytes:
    b = rng.integers(0, 256, n, dtype=np.uint8)
    b[rng.random(n) < sync_density] = 0x55
    return b.tobytes()


def make_stream(seed: int, n_segments: int = 40):
    """Returns (stream bytes, list of chunks).  Chunk sizes mimic calls of Decoder::process (a few characters each,
    now and then a long one)."""
    rng = np.random.default_rng(seed)
    parts = []
    next_pid = {}
    for _ in range(n_segments):
        kind = int(rng.integers(0, 10))
        if kind < 2:
            parts.append(junk(rng, int(rng.integers(1, 400)), float(rng.choice([0.0, 0.01, 0.05, 0.3]))))
        elif kind == 2:
            parts.append(b"$$CH0001,%d,12:00:00,52.1,21.5,1000*ABCD\n" % int(rng.integers(0, 99)))
        else:
            cs = CALLSIGNS[int(rng.integers(0, len(CALLSIGNS)))]
            iid = int(rng.integers(0, 3))
            pid = next_pid.get((cs, iid), 0)
            roll = int(rng.integers(0, 12))
            kw = {}
            if roll == 0:
                pid = max(pid - 1, 0)                     # retransmission of the previous id
            elif roll == 1:
                kw = dict(width16=int(rng.integers(1, 9)), height16=int(rng.integers(1, 9)))   # resolution change
            elif roll == 2:
                kw = dict(mcu_id=0xFFFF, mcu_offset=250)  # "no MCU starts in this packet": offset is not checked
            elif roll == 3:
                kw = dict(mcu_id=int(rng.integers(0, 40)), flags=int(rng.integers(0, 8)))   # may violate mcu_id < mcu_count
            next_pid[(cs, iid)] = pid + 1
            pkt = random_packet(rng, cs, iid, pid, **kw)
            fate = int(rng.integers(0, 10))
            if fate < 4:
                pass
            elif fate < 7:
                pkt = corrupt(pkt, int(rng.integers(1, 17)), rng)
            elif fate == 7:
                pkt = corrupt(pkt, int(rng.integers(17, 40)), rng)
            elif fate == 8:
                pkt = pkt[:int(rng.integers(2, 255))]     # truncated: the next segment follows immediately
            else:
                pkt = corrupt(pkt, int(rng.integers(0, 10)), rng, lo=0)   # may hit the sync byte
            parts.append(pkt)
    stream = b"".join(parts)
    chunks, o = [], 0
    while o < len(stream):
        n = int(rng.integers(1, 13)) if rng.random() > 0.03 else int(rng.integers(100, 700))
        chunks.append(stream[o:o + n])
        o += n
    return stream, chunks


def accepted_windows(stream: bytes):
    """[(pos, corrected packet, errors)] for every 0x55-started 256-byte window the published packet test accepts."""
    out = []
    a = np.frombuffer(stream, dtype=np.uint8)
    for p in np.flatnonzero(a[:max(len(a) - 255, 0)] == 0x55):
        v, e, c = po.ssdv_is_packet(stream[p:p + 256])
        if v == 0:
            out.append((int(p), c, e))
    return out


def wrapper_events(kind: str, chunks):
    """Transcript of the oracle's SSDV wrapper over the chunks: events + {(callsign, image id): packets}."""
    d = (po.RefDecoder if kind == "ref" else po.PortDecoder)(po.make_config())
    for c in chunks:
        d.ssdv_push(c)
    ev = d.ssdv_events()
    images = {(e[1], e[2]): d.ssdv_image(e[1], e[2]) for e in ev}
    return ev, images
