"""Snappy pages (float tag 2: few-decimal values or NaN, lib/compress/float.go:77-85,206-208; time tag 3:
lib/encoding/timestamp.go:132-148) open, decode and aggregate like any other page: the loader transcodes them to raw pages
(opengemini_b200/csrc/snappy_load.cuh).  Blocks come from the oracle's greedy encoder (literals + copies) and from a
hand-assembled block with explicit copy elements, so the device decoder is checked independently of that encoder."""
import struct

import numpy as np
import pytest

import oracle
from opengemini_b200 import AggQuery, Shard
from opengemini_b200 import _lib as L

pytestmark = pytest.mark.gpu
T0, SEC = 1_700_000_000_000_000_000, 1_000_000_000
ALL6 = ["count", "sum", "min", "max", "first", "last"]


@pytest.fixture(scope="module", autouse=True)
def _device():
    Shard.init(0)


def _snappy_elements(block):
    """[(kind, length, offset)] of a Snappy block (kind 0 literal, 1/2/3 copies)."""
    block = bytes(block)  # python ints: numpy uint8 arithmetic would wrap
    n, s, sh = 0, 0, 0
    while True:
        b = block[s]; s += 1
        n |= (b & 0x7f) << sh; sh += 7
        if b < 0x80:
            break
    out = []
    while s < len(block):
        tag = block[s]
        k = tag & 3
        if k == 0:
            x = tag >> 2
            if x < 60:
                s += 1
            else:
                nb = x - 59
                x = int.from_bytes(bytes(block[s + 1:s + 1 + nb]), "little"); s += 1 + nb
            out.append((0, x + 1, 0)); s += x + 1
        elif k == 1:
            out.append((1, 4 + ((tag >> 2) & 7), ((tag & 0xe0) << 3) | block[s + 1])); s += 2
        elif k == 2:
            out.append((2, 1 + (tag >> 2), block[s + 1] | (block[s + 2] << 8))); s += 3
        else:
            out.append((3, 1 + (tag >> 2), int.from_bytes(bytes(block[s + 1:s + 5]), "little"))); s += 5
    return n, out


def _shard_of(series_values, n=1000, times=None):
    pages, tpages, tmins, tmaxs, ssb = [], [], [], [], [0]
    for si, v in enumerate(series_values):
        k = (len(v) + n - 1) // n
        for g in range(k):
            pages.append(oracle.field_page_encode(L.TYPE_FLOAT, v[g * n:(g + 1) * n]))
            t = times[si][g * n:(g + 1) * n] if times is not None else T0 + (np.arange(min(n, len(v) - g * n), dtype=np.int64) + g * n) * SEC
            tpages.append(oracle.time_page_encode(t)); tmins.append(t[0]); tmaxs.append(t[-1])
        ssb.append(ssb[-1] + k)
    blob, offs, lens, pos = [], [], [], 0
    for p in pages + tpages:
        offs.append(pos); lens.append(p.size); blob.append(p); pos += p.size
    nseg = ssb[-1]
    sh = Shard.open(np.concatenate(blob), np.arange(1, len(series_values) + 1), ssb, tmins, tmaxs,
                    [("v", L.TYPE_FLOAT, offs[:nseg], lens[:nseg])], offs[nseg:], lens[nseg:])
    return sh, pages, tpages


def _agg_check(sh, sd, tmin, tmax, label, nan=False):
    for iv in (60 * SEC, 7 * SEC, 0):
        for calls in [[(f, 0)] for f in ALL6] + [[(f, 0) for f in ALL6]]:
            for group in ("all", "series"):
                flagsets = (L.Q_STRICT_ORDER, L.Q_NO_FUSED, L.Q_NO_FAST | L.Q_STRICT_ORDER) if group == "all" else (0,)
                for flags in flagsets:
                    q = AggQuery(sh, calls, iv, tmin, tmax, group=group, flags=flags).run()
                    got, ref = q.dense_host(), oracle.scan(sd, q.desc, threads=1)
                    for k, (f, _c) in enumerate(calls):
                        rv = ref["cols"][k]["valid"].astype(bool)
                        assert np.array_equal(got["cols"][k]["valid"].astype(bool), rv), (label, f, iv, group, flags)
                        assert np.array_equal(got["cols"][k]["values"].view(np.uint64)[rv], ref["cols"][k]["values"][rv]), (label, f, iv, group, flags)
                    q.close()


def test_few_decimal_floats_take_the_snappy_route_and_aggregate_bit_exact():
    rng = np.random.default_rng(31)
    rows = 3000
    series = [np.round(20.0 + rng.random(rows) * 5 + (s % 3), 2) for s in range(40)]   # <= 2 decimals -> lessDecimal -> Snappy
    series[3][1200:1210] = 21.25                                                      # a run: long copies
    sh, pages, _ = _shard_of(series)
    n_snappy = 0
    for p in pages:
        assert p[0] == 31 and (p[5] >> 4) == 2, "the oracle encoder is expected to route 2-decimal values to Snappy"
        _n, els = _snappy_elements(p[6:])
        n_snappy += any(k != 0 for k, _l, _o in els)
    assert n_snappy > len(pages) // 2, "blocks with copy elements expected"
    sd = oracle.shard_desc_from_export(sh.export())  # transcoded pages: the oracle reads the same raw bytes the kernels read
    for seg in (0, 7, 119):
        rec = sh.decode_segment(seg)
        s_, g = divmod(seg, 3)
        assert np.array_equal(rec["cols"][0]["values"], series[s_][g * 1000:(g + 1) * 1000])
    _agg_check(sh, sd, T0, T0 + (rows - 1) * SEC, "2-decimal")
    q = AggQuery(sh, [("sum", 0), ("count", 0), ("max", 0)], 60 * SEC, T0, T0 + (rows - 1) * SEC).run()
    assert q.stats()["path"] == 3 and q.stats()["general_segments"] == 0  # transcoded pages run on the fused kernel (XOR-delta lanes)
    ref = oracle.scan(sd, q.desc, threads=1)
    got = q.dense_host()
    assert np.allclose(got["cols"][0]["values"], ref["cols"][0]["values"].view(np.float64), rtol=1e-12, atol=0)
    assert np.array_equal(got["cols"][1]["values"].view(np.uint64), ref["cols"][1]["values"])
    assert np.array_equal(got["cols"][2]["values"].view(np.uint64), ref["cols"][2]["values"])
    q.close(); sh.close()


def test_nan_segments_take_the_snappy_route():
    rng = np.random.default_rng(32)
    rows = 2000
    series = []
    for s in range(12):
        v = np.round(100.0 + rng.random(rows), 2)  # few decimals keep the Snappy block under 90 % of raw; full-mantissa NaN segments end up raw
        v[rng.integers(0, rows, 40)] = np.nan
        series.append(v)
    sh, pages, _ = _shard_of(series)
    assert all((p[5] >> 4) == 2 for p in pages)
    sd = oracle.shard_desc_from_export(sh.export())
    for iv in (60 * SEC, 0):
        for f in ALL6:
            for flags in (L.Q_STRICT_ORDER, L.Q_NO_FUSED):
                q = AggQuery(sh, [(f, 0)], iv, T0, T0 + (rows - 1) * SEC, flags=flags).run()
                got, ref = q.dense_host(), oracle.scan(sd, q.desc, threads=1)
                rv = ref["cols"][0]["valid"].astype(bool)
                assert np.array_equal(got["cols"][0]["valid"].astype(bool), rv)
                assert np.array_equal(got["cols"][0]["values"].view(np.uint64)[rv], ref["cols"][0]["values"][rv]), (f, iv, flags)
                q.close()
    sh.close()


def test_time_pages_with_huge_deltas_take_the_snappy_route():
    rows = 1000
    t = (np.arange(rows, dtype=np.int64) * SEC) - 4_000_000_000_000_000_000
    t[500:] += 8_000_000_000_000_000_000  # one delta above 2^60: neither const-delta nor Simple8b
    v = 100.0 + np.random.default_rng(5).random(rows)
    sh, _pages, tpages = _shard_of([v], times=[t])
    assert tpages[0][0] == 32 and (tpages[0][5] >> 4) == 3, "time page expected on the Snappy route"
    rec = sh.decode_segment(0)
    assert np.array_equal(rec["times"], t)
    sd = oracle.shard_desc_from_export(sh.export())
    for calls in ([("sum", 0), ("count", 0)], [("first", 0)], [("last", 0)]):
        q = AggQuery(sh, calls, 3600 * SEC, int(t[0]), int(t[-1]), flags=L.Q_STRICT_ORDER).run()
        got, ref = q.dense_host(), oracle.scan(sd, q.desc, threads=1)
        for k in range(len(calls)):
            rv = ref["cols"][k]["valid"].astype(bool)
            assert np.array_equal(got["cols"][k]["valid"].astype(bool), rv)
            assert np.array_equal(got["cols"][k]["values"].view(np.uint64)[rv], ref["cols"][k]["values"][rv])
        q.close()
    sh.close()


def test_hand_assembled_block_with_every_copy_form():
    """Independent of any encoder: literal + 1-byte-offset copy + 2-byte-offset copy (overlapping, pattern repeat) + 4-byte-offset
    copy + a 61-tag literal, assembled here byte by byte from the Snappy format description."""
    n = 64
    base = (100.0 + np.arange(8) * 0.5).astype("<f8").tobytes()          # 64 bytes, values 100.0 .. 103.5
    want = bytearray(base)
    block = bytearray()
    block += bytes([0x80 | ((8 * n) & 0x7f), (8 * n) >> 7])              # uvarint 512
    block += bytes([(60 << 2) | 0, 63]) + base                           # literal, length 64 via the 1-byte length form (tag 60)
    block += bytes([((64 >> 8) << 5) | ((8 - 4) << 2) | 1, 64])           # copy1: length 8, offset 64  -> value 100.0 again
    want += want[-64:-56]
    block += bytes([((64 - 1) << 2) | 2, 8, 0])                          # copy2: length 64, offset 8: overlapping -> 100.0 x 8
    for _ in range(64):
        want.append(want[-8])
    block += bytes([((56 - 1) << 2) | 3]) + struct.pack("<I", 136)       # copy4: length 56, offset 136
    for _ in range(56):
        want.append(want[-136])
    lit = (7.25 + np.arange(40) * 0.25).astype("<f8").tobytes()          # 320 bytes: literal with a 2-byte length (tag 61)
    block += bytes([(61 << 2) | 0]) + struct.pack("<H", len(lit) - 1) + lit
    want += lit
    assert len(want) == 8 * n
    vals = np.frombuffer(bytes(want), "<f8")
    page = np.frombuffer(bytes([31]) + struct.pack(">I", n) + bytes([0x20]) + bytes(block), np.uint8)
    t = T0 + np.arange(n, dtype=np.int64) * SEC
    tp = oracle.time_page_encode(t)
    data = np.concatenate([page, tp])
    sh = Shard.open(data, [1], [0, 1], [int(t[0])], [int(t[-1])], [("v", L.TYPE_FLOAT, [0], [page.size])], [page.size], [tp.size])
    rec = sh.decode_segment(0)
    assert np.array_equal(rec["cols"][0]["values"].view(np.uint64), vals.view(np.uint64))
    q = AggQuery(sh, [("sum", 0), ("count", 0), ("max", 0)], 10 * SEC, int(t[0]), int(t[-1]), flags=L.Q_STRICT_ORDER).run()
    d = q.dense_host()
    b = (t - d["start"]) // (10 * SEC)
    want_sum = np.zeros(d["n_buckets"]); want_max = np.full(d["n_buckets"], -np.inf)
    for i in range(n):  # sequential adds, like the reducer
        want_sum[b[i]] = want_sum[b[i]] + vals[i]; want_max[b[i]] = max(want_max[b[i]], vals[i])
    assert np.array_equal(d["cols"][0]["values"], want_sum) and np.array_equal(d["cols"][2]["values"], want_max)
    q.close(); sh.close()
    # a truncated / inconsistent block is a corrupt page, not a crash
    bad = page.copy(); bad[8] = 0xff
    with pytest.raises(L.OgpuError) as ei:
        Shard.open(np.concatenate([bad, tp]), [1], [0, 1], [int(t[0])], [int(t[-1])], [("v", L.TYPE_FLOAT, [0], [bad.size])], [bad.size], [tp.size])
    assert ei.value.status == L.OG_E_CORRUPT
