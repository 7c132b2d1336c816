"""TSSP container reader (csrc/tssp.cpp, og_tssp_parse / og_tssp_desc) against files written by tests/tssp_file.py — a plain
Python restatement of the Go marshal code (msbuilder.go Flush, ChunkMeta/ColumnMeta/MetaIndex/Trailer marshal).  The parse is
host code: these tests run without a GPU, except the last one, which opens the parsed file as a shard and queries it."""
import ctypes as C
import random

import numpy as np
import pytest

import oracle
import tssp_file
from opengemini_b200 import _lib as L

T0, SEC = 1_700_000_000_000_000_000, 1_000_000_000
COLS = [(L.TYPE_FLOAT, L.SYNTH_F_HI, 0), (L.TYPE_INT, L.SYNTH_INT_WALK, 100), (L.TYPE_BOOL, L.SYNTH_BOOL, 0)]
NAMES = [b"a_float", b"b_int", b"c_bool"]


def chunks_of(hs, names=NAMES, cols=COLS, lacking=()):
    """(series, column) pairs in `lacking` are left out of that series' chunk: series of one file may have different fields."""
    d = hs.desc
    out = []
    for s in range(d.n_series):
        g0, g1 = d.series_seg_begin[s], d.series_seg_begin[s + 1]
        use = [c for c in range(len(cols)) if (s, c) not in lacking]
        out.append(dict(sid=int(d.sids[s]), tmin=[d.seg_tmin[g] for g in range(g0, g1)], tmax=[d.seg_tmax[g] for g in range(g0, g1)],
                        columns=[(names[c], cols[c][0], [hs.page(c, g).tobytes() for g in range(g0, g1)]) for c in use],
                        time=[hs.page(len(cols), g).tobytes() for g in range(g0, g1)]))
    return out


def parse(blob):
    buf = np.frombuffer(blob, dtype=np.uint8) if len(blob) else np.zeros(1, np.uint8)
    h = C.c_void_p()
    rc = L.lib().og_tssp_parse(buf.ctypes.data, len(blob), C.byref(h))
    return rc, h, buf


@pytest.fixture(scope="module")
def shard():
    hs = oracle.HostShard(7, 2500, COLS, t0=T0, dt=SEC, seed=5)
    return hs, chunks_of(hs, lacking={(3, 1)})


@pytest.mark.parametrize("per_block", [1, 3, 100])
def test_directory_matches_what_was_written(shard, per_block):
    hs, chunks = shard
    d = hs.desc
    blob, directory = tssp_file.build(chunks, measurement=b"cpu_0001", metas_per_block=per_block)
    rc, h, keep = parse(blob)
    assert rc == L.OG_OK, L.lib().og_last_error()
    sd = L.ShardDesc()
    assert L.lib().og_tssp_desc(h, C.byref(sd)) == L.OG_OK
    assert L.lib().og_tssp_measurement(h) == b"cpu_0001"
    lo, hi = C.c_int64(), C.c_int64()
    assert L.lib().og_tssp_time_range(h, C.byref(lo), C.byref(hi)) == L.OG_OK
    assert (lo.value, hi.value) == (T0, T0 + 2499 * SEC)
    assert (sd.n_series, sd.n_segments, sd.n_columns) == (d.n_series, d.n_segments, 3)
    assert sd.data_len == len(blob)
    for s in range(d.n_series):
        assert sd.sids[s] == d.sids[s] and sd.series_seg_begin[s] == d.series_seg_begin[s]
        for k, g in enumerate(range(d.series_seg_begin[s], d.series_seg_begin[s + 1])):
            assert (sd.seg_tmin[g], sd.seg_tmax[g]) == (d.seg_tmin[g], d.seg_tmax[g])
            assert (sd.time_page_off[g], sd.time_page_len[g]) == directory[sd.sids[s]][b"time"][k]
            for c in range(3):
                assert sd.columns[c].name == NAMES[c] and sd.columns[c].type == COLS[c][0]
                off, ln = sd.columns[c].page_off[g], sd.columns[c].page_len[g]
                if (s, c) == (3, 1):
                    assert ln == 0
                else:
                    assert (off, ln) == directory[sd.sids[s]][NAMES[c]][k]
                    assert blob[off:off + ln] == hs.page(c, g).tobytes()
    assert sd.series_seg_begin[d.n_series] == d.n_segments
    L.lib().og_tssp_free(h)


def test_columns_come_out_sorted_by_name_whatever_the_chunk_order(shard):
    hs, _ = shard
    names = [b"zeta", b"alpha", b"mid"]                     # the writer sorts a chunk's columns by name, like the reference
    chunks = chunks_of(hs, names=names)
    for ch in chunks:
        ch["columns"].sort(key=lambda t: t[0])
    rc, h, keep = parse(tssp_file.build(chunks)[0])
    assert rc == L.OG_OK
    sd = L.ShardDesc(); L.lib().og_tssp_desc(h, C.byref(sd))
    assert [sd.columns[c].name for c in range(3)] == [b"alpha", b"mid", b"zeta"]
    assert [sd.columns[c].type for c in range(3)] == [L.TYPE_INT, L.TYPE_BOOL, L.TYPE_FLOAT]
    L.lib().og_tssp_free(h)


def test_trailer_forms(shard):
    _, chunks = shard
    for kw, want in ((dict(compress_flag=1), L.OG_E_UNSUPPORTED), (dict(legacy_extra=b"\x01"), L.OG_OK), (dict(legacy_extra=b"\x01\x00"), L.OG_OK),
                     (dict(legacy_extra=b"\x01\x02"), L.OG_E_UNSUPPORTED), (dict(bloom_bytes=0, idtime_bytes=0), L.OG_OK)):
        rc, h, keep = parse(tssp_file.build(chunks, **kw)[0])
        assert rc == want, (kw, L.lib().og_last_error())
        if rc == L.OG_OK:
            L.lib().og_tssp_free(h)


def test_damaged_files_are_refused_not_crashed_on(shard):
    _, chunks = shard
    blob, _ = tssp_file.build(chunks)
    rnd = random.Random(1)
    for cut in list(range(0, 40)) + [rnd.randrange(len(blob)) for _ in range(300)]:
        rc, h, keep = parse(blob[:cut])
        assert rc in (L.OG_E_CORRUPT, L.OG_E_UNSUPPORTED), cut
    assert parse(b"53ac2022" + blob[8:])[0] == L.OG_E_CORRUPT
    assert parse(blob[:8] + (3).to_bytes(8, "big") + blob[16:])[0] == L.OG_E_UNSUPPORTED
    # single-bit damage in the metadata tail: either refused, or a directory whose every page still lies inside the data region
    tail0 = len(blob) - 600
    for _ in range(2000):
        b2 = bytearray(blob)
        b2[rnd.randrange(tail0, len(blob))] ^= 1 << rnd.randrange(8)
        rc, h, keep = parse(bytes(b2))
        if rc == L.OG_OK:
            sd = L.ShardDesc(); L.lib().og_tssp_desc(h, C.byref(sd))
            for g in range(sd.n_segments):
                assert sd.time_page_off[g] + sd.time_page_len[g] <= len(b2)
                for c in range(sd.n_columns):
                    assert sd.columns[c].page_off[g] + sd.columns[c].page_len[g] <= len(b2)
            L.lib().og_tssp_free(h)
    # chunks out of series-id order
    rc, h, keep = parse(tssp_file.build([chunks[1], chunks[0]] + chunks[2:])[0])
    assert rc == L.OG_E_CORRUPT


@pytest.mark.gpu
def test_query_over_a_parsed_file_equals_the_oracle(shard):
    from opengemini_b200 import AggQuery, Shard
    hs, chunks = shard
    # every series carries every column here: the oracle scans hs.desc, the GPU the parsed file
    blob, _ = tssp_file.build(chunks_of(hs), metas_per_block=2)
    Shard.init(0)
    sh = Shard.open_tssp(blob)
    assert sh.measurement == "mst_0000" and sh.columns == [("a_float", L.TYPE_FLOAT), ("b_int", L.TYPE_INT), ("c_bool", L.TYPE_BOOL)]
    info = sh.info()
    assert info["n_series"] == 7 and info["n_rows"] == 7 * 2500
    calls = [("sum", 0), ("count", 1), ("max", 0), ("count", 2), ("min", 1)]
    for flags in (L.Q_STRICT_ORDER, L.Q_STRICT_ORDER | L.Q_NO_FUSED):
        q = AggQuery(sh, calls, 60 * SEC, T0 + 17 * SEC, T0 + 2400 * SEC, flags=flags).run()
        got, ref = q.dense_host(), oracle.scan(hs.desc, q.desc, threads=1)
        for k in range(len(calls)):
            m = ref["cols"][k]["valid"].astype(bool)
            assert np.array_equal(got["cols"][k]["valid"].astype(bool), m)
            assert np.array_equal(got["cols"][k]["values"].view(np.uint64)[m], ref["cols"][k]["values"][m]), (flags, k)
        q.close()
    for seg in (0, 4, 20):
        a = sh.decode_segment(seg)
        s = next(i for i in range(7) if hs.desc.series_seg_begin[i + 1] > seg)
        assert a["times"][0] == hs.desc.seg_tmin[seg] and a["times"][-1] == hs.desc.seg_tmax[seg], s
    sh.close()
    # a series that lacks a column contributes nulls
    blob2, _ = tssp_file.build(chunks, metas_per_block=2)
    sh2 = Shard.open_tssp(blob2)
    q = AggQuery(sh2, [("count", 1), ("count", 0)], 0, T0, T0 + 2499 * SEC, group="series", flags=L.Q_STRICT_ORDER).run()
    d = q.dense_host()
    cnt_int = d["cols"][0]["values"].astype(np.int64) * d["cols"][0]["valid"]
    cnt_f = d["cols"][1]["values"].astype(np.int64) * d["cols"][1]["valid"]
    assert cnt_int.reshape(-1)[3] == 0 and cnt_f.reshape(-1)[3] == 2500 and int(cnt_f.sum()) == 7 * 2500
    q.close(); sh2.close()
