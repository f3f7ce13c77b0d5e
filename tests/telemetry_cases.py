"""Shared telemetry-layer cases: request lines in the protocol of oracle/ref_telemetry.cpp, our answers through the
C ABI in the same format, and a field-wise comparison (floats are exchanged as C hex floats and compared bit for bit)."""
import os
import subprocess

import numpy as np

from habdec_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "telemetry_ref")

# 2026-10-17 23:59:59 UTC, 2026-10-18 00:01:40, 2026-10-18 12:00:00, a leap day just before midnight, new year's eve
CLOCKS = [1792281599, 1792281700, 1792324800, 1709251199, 1798761599, 0]

HAND = [
    "TIME\t123456", "TIME\t12:34:56", "TIME\t12_34_56", "TIME\t1234", "TIME\t12:34", "TIME\t12:34:56.7", "TIME\t12:34:56.",
    "TIME\t12:34:", "TIME\t12345", "TIME\t1:2:3", "TIME\t", "TIME\t12:34:5", "TIME\t12::34", "TIME\t12:34:56.789012345",
    "TIME\t99x99y99", "TIME\t12.34.56.5", "TIME\t123:45", "TIME\t12345678", "TIME\t12:34.5", "TIME\t1234\r", "TIME\t12:34:56 ",
    "POS\t52.1234", "POS\t121.345", "POS\t-52.1234", "POS\t-121.345", "POS\t5205.5857", "POS\t02112.7309", "POS\t-5205.5857",
    "POS\t-02112.7309", "POS\t", "POS\t-", "POS\t5", "POS\t.5", "POS\t5.", "POS\t1.5", "POS\t12", "POS\t123456.7", "POS\tab.cd",
    "POS\tabcd.ef", "POS\t-ab.cd", "POS\t 2.5", "POS\t52.12.34", "POS\t0000.0000", "POS\t9999.9999", "POS\t18000.0000",
    "POS\t1e10.5", "POS\t12.5e40", "POS\t1234.5e-50", "POS\t+52.5", "POS\t--52.5", "POS\tnan.", "POS\t0x1.8", "POS\t00.0",
    "SENT\t$$CALL,1,12:00:00,52.1234,21.5678,1000", "SENT\tCALL,1,12:00:00,52.1234,21.5678,1000,extra,fields",
    "SENT\tCALL,1,12:00:00,52.1234,21.5678", "SENT\tCALL,x,12:00:00,52.1234,21.5678,1000", "SENT\tCALL,3,12:00:00,0,0,1000",
    "SENT\tCALL,3,12:00:00,0.0,0.0,1000", "SENT\tCALL,3,12:00:00,00.0,00.0,1000", "SENT\tCALL,3,120000,5205.5857,-02112.7309,-12.5",
    "SENT\tCALL,3,bad,52.1,21.5,10", "SENT\tCALL,3,12:00:00,,21.5,10", "SENT\tCALL,3,12:00:00,52.1,21.5,", "SENT\tCALL,3,12:00:00,52.1,21.5,abc",
    "SENT\tab$$cd,3,12:00:00,52.1,21.5,10", "SENT\tabc$$,3,12:00:00,52.1,21.5,10", "SENT\t$,3,12:00:00,52.1,21.5,10",
    "SENT\t$a$b,3,12:00:00,52.1,21.5,10", "SENT\t,3,12:00:00,52.1,21.5,10", "SENT\tCALL,99999999999,12:00:00,52.1,21.5,10",
    "SENT\tCALL, 42abc,12:00:00,52.1,21.5,10", "SENT\tCALL,-7,23:59:59.5,52.1,0,1e3", "SENT\tCALL,7,00:00:01,0,21.5,1e39",
    "SENT\tCALL,7,00:00:01,12,21.5,10", "SENT\tCALL,7,0000,52.5,21.5,10", "SENT\t,,,,,", "SENT\tCALL,1,12:00:00,52.1234,21.5678,1e-46",
    "DIST\t52.0\t21.0\t100\t52.5\t21.9\t30000", "DIST\t0\t0\t0\t0\t0\t0", "DIST\t52\t21\t100\t52\t21\t100", "DIST\t-33.9\t151.2\t20\t51.5\t-0.12\t11000",
    "DIST\t89.99\t0\t0\t-89.99\t180\t0", "DIST\t10\t179.9\t0\t10\t-179.9\t5000", "DIST\t52\t21\t100\t52\t21\t35000", "DIST\t52\t21\t35000\t52\t21\t100",
]


def fuzz_requests(seed: int, n: int) -> list[str]:
    rng = np.random.default_rng(seed)
    digits = "0123456789"
    seps = [":", "", "_", ".", "-", " ", "x", "::"]

    def num(k):
        return "".join(rng.choice(list(digits), k))

    def rand_time():
        r = int(rng.integers(0, 10))
        if r < 6:
            s = num(2) + str(rng.choice(seps)) + num(2)
            if rng.random() < 0.8:
                s += str(rng.choice(seps)) + num(int(rng.choice([2, 2, 2, 1, 3])))
                if rng.random() < 0.4:
                    s += "." + num(int(rng.integers(0, 5)))
            return s
        return "".join(rng.choice(list(digits + ":._ x"), int(rng.integers(0, 10))))

    def rand_pos():
        r = int(rng.integers(0, 10))
        sign = "-" if rng.random() < 0.3 else ""
        if r < 4:
            return sign + "%.*f" % (int(rng.integers(0, 7)), rng.uniform(0, 180))
        if r < 7:
            return sign + "%0*.*f" % (int(rng.choice([9, 10])), 4, rng.uniform(0, 18000))
        if r < 8:
            return sign + num(int(rng.integers(0, 7))) + "." + num(int(rng.integers(0, 6)))
        return "".join(rng.choice(list(digits + ".-+e "), int(rng.integers(0, 9))))

    out = []
    for _ in range(n):
        k = int(rng.integers(0, 10))
        if k < 2:
            out.append("TIME\t" + rand_time())
        elif k < 4:
            out.append("POS\t" + rand_pos())
        elif k < 5:
            out.append("STAMP\t%d\t%d\t%s" % (rng.integers(0, 24), rng.integers(0, 60), rng.choice(["0", "7", "59", "5.5", "12.25", "59.999", "0.5", "33.3333333"])))
        elif k < 6:
            v = [rng.uniform(-90, 90), rng.uniform(-180, 180), rng.uniform(0, 500), rng.uniform(-90, 90), rng.uniform(-180, 180), rng.uniform(0, 40000)]
            if rng.random() < 0.5:      # a payload near the station: the usual case
                v[3] = v[0] + rng.normal(0, 0.5); v[4] = v[1] + rng.normal(0, 0.5)
            out.append("DIST\t" + "\t".join(repr(float(x)) for x in v))
        else:
            call = str(rng.choice(["CALL", "$$CALL", "$$$HAB1", "A", "x$y", "", "Q$"]))
            fields = [call, str(rng.integers(-5, 1000)) if rng.random() < 0.9 else rand_pos(), rand_time() if rng.random() < 0.5 else "12:00:00",
                      rand_pos(), rand_pos(), "%g" % rng.uniform(-100, 40000) if rng.random() < 0.9 else rand_pos()]
            fields += ["x"] * int(rng.integers(0, 3))
            if rng.random() < 0.05:
                fields = fields[:int(rng.integers(0, 6))]
            line = ",".join(fields)
            if rng.random() < 0.5 and len(fields) > 1:
                c = line.index(",")
                out.append("CB\t%s\t%s\t%04X" % (line[:c], line[c + 1:], rng.integers(0, 65536)))
            else:
                out.append("SENT\t" + line)
        if rng.random() < 0.03:
            out.append("NOW\t%d" % int(rng.choice(CLOCKS[:-1])))
        if rng.random() < 0.02:
            out.append("STATION\t%g\t%g\t%g" % ((rng.uniform(-80, 80), rng.uniform(-180, 180), rng.uniform(0, 900)) if rng.random() < 0.8 else (0, 21, 100)))
    return out


def all_requests(seed: int = 1, n_fuzz: int = 1500) -> list[str]:
    req = ["NOW\t%d" % CLOCKS[2]] + list(HAND)
    for t in CLOCKS[:-1]:
        req.append("NOW\t%d" % t)
        for h in (23, 0, 1, 22, 12):
            req.append("STAMP\t%d\t59\t59.25" % h)
        req.append("SENT\tCALL,5,23:59:59,52.1,21.5,10")
        req.append("SENT\tCALL,5,00:00:01,52.1,21.5,10")
    req.append("NOW\t%d" % CLOCKS[2])
    req.append("STATION\t52.0\t21.0\t100")
    for i, (la, lo, al) in enumerate([(52.1234, 21.5678, 1000), (52.3, 21.9, 5000), (52.2, 21.7, 12000), (52.2, 21.7, 12000), (51.0, 20.0, 30000)]):
        req.append("CB\tCALL\t%d,12:00:%02d,%s,%s,%s\tAB%02d" % (i if i != 3 else 2, i, la, lo, al, i))
    req.append("CB\tCALL\tnope\t0000")
    req.append("STATION\t0\t21.0\t100")
    req.append("CB\tCALL\t1,12:00:00,52.1234,21.5678,1000\tABCD")
    return req + ["NOW\t%d" % CLOCKS[2], "STATION\t48.5\t17.25\t250"] + fuzz_requests(seed, n_fuzz)


def run_reference(requests: list[str]) -> list[str]:
    p = subprocess.run([REF_BIN], input=("\n".join(requests) + "\n").encode("latin-1"), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True)
    # parse_sentence prints "Failed parsing time string" on stdout: those lines are not answers
    lines = [ln for ln in p.stdout.decode("latin-1").split("\n")[:-1] if not ln.startswith("Failed parsing time string")]
    assert len(lines) == len(requests), (len(lines), len(requests))
    return lines


def hx(v: float) -> str:
    return float(v).hex()


class Ours:
    """Answers the same protocol through the C ABI (stateless functions + one hbd_tracker per STATION line)."""

    def __init__(self):
        self.now = -1
        self.station = (0.0, 0.0, 0.0)
        self.tracker = api.Tracker()

    def answer(self, line: str) -> str:
        f = line.split("\t")
        b = [x.encode("latin-1") for x in f]
        if f[0] == "NOW":
            self.now = int(f[1]); self.tracker.set_clock(self.now)
            return "OK"
        if f[0] == "TIME":
            rc, v = api.parse_sentence_time(b[1])
            return "THROW" if rc < 0 else "NONE" if rc == 0 else "%d\t%d\t%s" % (v[0], v[1], hx(v[2]))
        if f[0] == "POS":
            rc, v = api.parse_gps_pos(b[1])
            return "THROW" if rc < 0 else hx(v)
        if f[0] == "STAMP":
            return api.timestamp_from_hms(int(f[1]), int(f[2]), float(np.float32(f[3])), self.now).decode()
        if f[0] == "SENT":
            rc, t = api.parse_sentence(b[1], self.now)
            if rc != 1:
                return "THROW" if rc < 0 else "NONE"
            return "\t".join([t["payload_callsign"].decode("latin-1"), t["datetime"].decode(), str(t["frame"]), hx(t["lat"]), hx(t["lon"]), hx(t["alt"]),
                              t["tracking"].decode("latin-1")])
        if f[0] == "DIST":
            g = api.calc_gps_distance(*[float(x) for x in f[1:7]])
            return "\t".join(hx(x) for x in (g.dist_line_, g.dist_circle_, g.dist_radians_, g.elevation_, g.bearing_))
        if f[0] == "STATION":
            self.station = tuple(float(np.float32(x)) for x in f[1:4])
            self.tracker = api.Tracker(station=self.station, now_unix=self.now)
            return "OK"
        if f[0] == "CB":
            rc = self.tracker.push(0, b[1], b[2], b[3])
            if rc != 1:
                return "THROW" if rc < 0 else "NONE"
            st = self.tracker.stats(0)
            d = st.D_
            return "\t".join([str(st.num_ok_)] + [hx(x) for x in (d.dist_line_, d.dist_circle_, d.dist_radians_, d.elevation_, d.bearing_, st.dist_max_, st.elev_min_)]
                             + [self.tracker.stats_payload(0).decode()])
        return "BAD"


def same(a: str, b: str) -> bool:
    """Field-wise equality; fields that parse as C hex floats are compared as numbers (bit exact)."""
    fa, fb = a.split("\t"), b.split("\t")
    if len(fa) != len(fb):
        return False
    for x, y in zip(fa, fb):
        if x == y:
            continue
        try:
            vx, vy = float.fromhex(x), float.fromhex(y)
        except ValueError:
            return False
        if not (vx == vy or (vx != vx and vy != vy)):
            return False
    return True
