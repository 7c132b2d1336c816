"""Two small rows of the hot-path table:
* count() on a STRING column (legal SQL; lib/encoding/string.go pages): only the page header is ever read — ValidCount of the
  null bitmap — so the GPU path serves it without decoding the payload; every other call on strings is refused cleanly.
* descending materialisation: og_decode_segment_ex(OG_DECODE_DESCENDING) hands a segment over reversed (values, validity bits,
  times), as reader.go:516-519,1035-1042 does for ORDER BY time DESC scans."""
import struct

import numpy as np
import pytest

import oracle
from opengemini_b200 import AggQuery, Shard
from opengemini_b200 import _lib as L

pytestmark = pytest.mark.gpu
T0, SEC = 1_700_000_000_000_000_000, 1_000_000_000
TYPE_STRING = 4


@pytest.fixture(scope="module", autouse=True)
def _device():
    Shard.init(0)


def _string_page(valid, payload=b"\x10opaque-string-block-bytes"):
    """A string column page as EncodeColumnHeader frames it: Full (34) / Empty (44) / normal (4) header + an opaque block."""
    rows = len(valid)
    nil = int(rows - valid.sum())
    if nil == 0:
        return np.frombuffer(bytes([34]) + struct.pack(">I", rows) + payload, np.uint8)
    if nil == rows:
        return np.frombuffer(bytes([44]) + struct.pack(">I", rows), np.uint8)
    bm = np.packbits(valid.astype(np.uint8), bitorder="little").tobytes()
    return np.frombuffer(bytes([TYPE_STRING]) + struct.pack(">I", len(bm)) + bm + struct.pack(">II", 0, nil) + payload, np.uint8)


def test_count_on_a_string_column():
    rng = np.random.default_rng(3)
    n_series, segs, n = 5, 4, 700
    pages, fpages, tpages, tmins, tmaxs, ssb, valids = [], [], [], [], [], [0], []
    for s in range(n_series):
        for g in range(segs):
            shape = (s + g) % 3
            valid = np.ones(n, bool) if shape == 0 else np.zeros(n, bool) if shape == 1 else rng.random(n) > 0.3
            valids.append(valid)
            pages.append(_string_page(valid))
            fpages.append(oracle.field_page_encode(L.TYPE_FLOAT, 100 + rng.random(n)))
            t = T0 + (np.arange(n, dtype=np.int64) + g * n) * SEC
            tpages.append(oracle.time_page_encode(t)); tmins.append(t[0]); tmaxs.append(t[-1])
        ssb.append(ssb[-1] + segs)
    blob, offs, lens, pos = [], [], [], 0
    for p in pages + fpages + tpages:
        offs.append(pos); lens.append(p.size); blob.append(p); pos += p.size
    nseg = ssb[-1]
    sh = Shard.open(np.concatenate(blob), np.arange(1, n_series + 1), ssb, tmins, tmaxs,
                    [("name", TYPE_STRING, offs[:nseg], lens[:nseg]), ("v", L.TYPE_FLOAT, offs[nseg:2 * nseg], lens[nseg:2 * nseg])],
                    offs[2 * nseg:], lens[2 * nseg:])
    tmax = T0 + (segs * n - 1) * SEC
    for group in ("all", "series"):
        q = AggQuery(sh, [("count", 0), ("count", 1)], 60 * SEC, T0, tmax, group=group).run()
        d = q.dense_host()
        nb = d["n_buckets"]
        want = np.zeros(d["n_groups"] * nb, np.int64); wantf = np.zeros_like(want)
        for s in range(n_series):
            for g in range(segs):
                b = ((np.arange(n) + g * n) * SEC + T0 - d["start"]) // (60 * SEC)
                base = (s * nb) if group == "series" else 0
                np.add.at(want, base + b[valids[s * segs + g]], 1)
                np.add.at(wantf, base + b, 1)
        assert np.array_equal(d["cols"][0]["values"] * d["cols"][0]["valid"], want)
        assert np.array_equal(d["cols"][0]["valid"].astype(bool), want > 0)   # a window without a non-null string is NULL, not 0
        assert np.array_equal(d["cols"][1]["values"] * d["cols"][1]["valid"], wantf)
        q.close()
    for calls in ([("max", 0)], [("first", 0)], [("sum", 0)]):
        with pytest.raises(L.OgpuError) as ei:
            AggQuery(sh, calls, 60 * SEC, T0, tmax)
        assert ei.value.status in (L.OG_E_UNSUPPORTED, L.OG_E_INVAL)
    sh.close()


def test_descending_materialisation_reverses_values_bitmaps_and_times():
    cols = [(L.TYPE_FLOAT, L.SYNTH_F_HI, 0), (L.TYPE_INT, L.SYNTH_INT_WALK, 200), (L.TYPE_BOOL, L.SYNTH_BOOL, 100), (L.TYPE_FLOAT, L.SYNTH_F_LO, 1000)]
    hs = oracle.HostShard(2, 1777, cols, t0=T0, dt=SEC, seed=21)
    sh = Shard.open_desc(hs.desc, keepalive=hs)
    for seg in range(hs.desc.n_segments):
        asc, desc = sh.decode_segment(seg), sh.decode_segment(seg, descending=True)
        assert np.array_equal(desc["times"], asc["times"][::-1])
        for c in range(len(cols)):
            a, dd = asc["cols"][c], desc["cols"][c]
            assert dd["len"] == a["len"] and dd["nil_count"] == a["nil_count"]
            assert np.array_equal(dd["valid"], a["valid"][::-1]), (seg, c)
            assert np.ascontiguousarray(dd["values"]).tobytes() == np.ascontiguousarray(a["values"][::-1]).tobytes(), (seg, c)
    sh.close()


def test_scan_cursor_emits_time_filtered_records_in_both_orders():
    """ScanCursor = KeyCursor.Next() of a plain scan: records per qualifying segment, cut to the range, ascending and descending;
    checked against the oracle's decode of the same pages."""
    from opengemini_b200 import ScanCursor
    cols = [(L.TYPE_FLOAT, L.SYNTH_F_HI, 100), (L.TYPE_INT, L.SYNTH_INT_WALK, 0), (L.TYPE_BOOL, L.SYNTH_BOOL, 300)]
    hs = oracle.HostShard(3, 3210, cols, t0=T0, dt=SEC, seed=31)
    sh = Shard.open_desc(hs.desc, keepalive=hs)
    tmin, tmax = T0 + 1500 * SEC + 1, T0 + 2999 * SEC
    want = []
    d = hs.desc
    for s in range(d.n_series):
        for g in range(d.series_seg_begin[s], d.series_seg_begin[s + 1]):
            ref = dict(times=oracle.time_page_decode(hs.page(len(cols), g)), cols=[])
            for c, (typ, _d, _n) in enumerate(cols):
                v, valid = oracle.field_page_decode(typ, hs.page(c, g))
                ref["cols"].append(dict(values=v, valid=valid))
            keep = (ref["times"] >= tmin) & (ref["times"] <= tmax)
            if keep.any():
                want.append((int(d.sids[s]), g, ref, keep))
    got = list(ScanCursor(sh, tmin, tmax))
    assert [(r["sid"], r["segment"]) for r in got] == [(w[0], w[1]) for w in want]
    for r, (_, g, ref, keep) in zip(got, want):
        assert np.array_equal(r["times"], ref["times"][keep]) and r["rows"] == int(keep.sum())
        for c in range(len(cols)):
            valid = ref["cols"][c]["valid"].astype(bool)
            assert np.array_equal(r["cols"][c]["valid"], valid[keep]), (g, c)
            assert r["cols"][c]["values"].tobytes() == ref["cols"][c]["values"][keep[valid]].tobytes(), (g, c)
    desc = list(ScanCursor(sh, tmin, tmax, ascending=False))
    assert len(desc) == len(got)
    by_key = {(r["sid"], r["segment"]): r for r in got}
    last = {}
    for r in desc:
        a = by_key[(r["sid"], r["segment"])]
        assert np.array_equal(r["times"], a["times"][::-1])
        for c in range(len(cols)):
            assert np.array_equal(r["cols"][c]["valid"], a["cols"][c]["valid"][::-1])
            assert r["cols"][c]["values"].tobytes() == a["cols"][c]["values"][::-1].tobytes()
        assert last.get(r["sid"], 1 << 62) > r["segment"]  # a series' segments come latest first
        last[r["sid"]] = r["segment"]
    sh.close()
