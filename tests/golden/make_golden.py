"""Writes tests/golden/*.npz: fixed page bytes + decoded values, and a small shard + expected aggregates.

The bytes come from the oracle's restated encoders (oracle/codec.cpp, record.cpp) — the reference holds no golden encoded bytes
(SURVEY §8c), so these fixtures pin the ORACLE against regressions and give the GPU decoders fixed inputs; they are not an
independent pin of the reference's wire format.  Three hand-derived vectors in tests/test_golden.py (time const-delta page, bool
bit-pack page, one-row page) come from the format description in SURVEY App. A instead.

usage: python tests/golden/make_golden.py     (needs only the CPU oracle; commit the .npz it writes)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.dirname(os.path.dirname(HERE)), os.path.dirname(HERE)]
import oracle  # noqa: E402
from opengemini_b200 import _lib as L  # noqa: E402

T0, SEC = 1_700_000_000_000_000_000, 1_000_000_000


def page_cases():
    rng = np.random.default_rng(20240917)
    n = 300
    cases = []

    def add(name, typ, cells, valid=None, times=None):
        times = T0 + np.arange(len(cells), dtype=np.int64) * SEC if times is None else times
        cases.append((name, typ, np.ascontiguousarray(cells), valid, np.ascontiguousarray(times, np.int64)))

    add("float_gorilla_noise", L.TYPE_FLOAT, 100.0 + rng.random(n))
    add("float_gorilla_walk", L.TYPE_FLOAT, 1000.0 + np.cumsum(rng.integers(-2, 3, n)).astype(np.float64))
    add("float_same", L.TYPE_FLOAT, np.full(n, 3.25))
    add("float_same_zero", L.TYPE_FLOAT, np.zeros(n))
    add("float_rle", L.TYPE_FLOAT, np.repeat(np.array([1.5, 0.0, -2.0, 7.0]), n // 4))
    add("float_raw_short", L.TYPE_FLOAT, np.array([1.0, -2.5, 3.75, 1e300]))
    add("float_raw_incompressible", L.TYPE_FLOAT, rng.integers(0, 2**63, n).view(np.float64) % 1e300)
    v = 50.0 + rng.random(n); ok = (rng.random(n) > 0.2).astype(np.uint8); ok[0] = 0; ok[-1] = 1
    add("float_nulls", L.TYPE_FLOAT, v, ok)
    add("float_all_null", L.TYPE_FLOAT, np.zeros(n), np.zeros(n, np.uint8))
    add("float_one_row", L.TYPE_FLOAT, np.array([42.5]))
    add("int_const_delta", L.TYPE_INT, (7 + 13 * np.arange(n)).astype(np.int64))
    add("int_simple8b", L.TYPE_INT, np.cumsum(rng.integers(-1000, 1001, n)).astype(np.int64))
    add("int_raw_pair", L.TYPE_INT, np.array([5, -9], np.int64))
    add("int_big_delta", L.TYPE_INT, np.where(np.arange(n) % 2 == 0, 1 << 40, -(1 << 40)).astype(np.int64))
    iv = np.cumsum(rng.integers(-5, 6, n)).astype(np.int64); iok = (rng.random(n) > 0.5).astype(np.uint8)
    add("int_nulls", L.TYPE_INT, iv, iok)
    add("bool_bits", L.TYPE_BOOL, (rng.random(n) > 0.5).astype(np.uint8))
    add("bool_nulls", L.TYPE_BOOL, (rng.random(n) > 0.5).astype(np.uint8), (rng.random(n) > 0.3).astype(np.uint8))
    jitter = T0 + np.cumsum(rng.integers(1, 5, n)).astype(np.int64) * 1_000_000
    add("time_simple8b_scaled", L.TYPE_INT, np.arange(n, dtype=np.int64), None, jitter)
    wild = T0 + np.cumsum(rng.integers(1, 2**40, n)).astype(np.int64)
    add("time_irregular", L.TYPE_INT, np.arange(n, dtype=np.int64), None, wild)
    return cases


def main():
    out = {}
    names = []
    for name, typ, cells, valid, times in page_cases():
        page = oracle.field_page_encode(typ, cells, valid)
        tpage = oracle.time_page_encode(times)
        vals, ok = oracle.field_page_decode(typ, page, cap=max(8, len(cells) + 8))
        names.append(name)
        out[f"{name}/type"] = np.int32(typ)
        out[f"{name}/page"] = page
        out[f"{name}/time_page"] = tpage
        out[f"{name}/values"] = vals.view(np.uint8) if typ != L.TYPE_BOOL else vals
        out[f"{name}/valid"] = ok.astype(np.uint8)
        out[f"{name}/times"] = times
        # the encoder/decoder pair must round-trip before the bytes become golden
        want = cells[valid.astype(bool)] if valid is not None else cells
        assert np.array_equal(vals.view(np.uint8), np.ascontiguousarray(want).view(np.uint8)), name
        assert np.array_equal(oracle.time_page_decode(tpage, cap=len(times) + 8), times), name
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "pages_v1.npz"), **out)

    # a small two-column shard (float noise + int walk with 5% nulls) and the oracle's aggregates over it
    cols = [(L.TYPE_FLOAT, L.SYNTH_F_HI, 0), (L.TYPE_INT, L.SYNTH_INT_WALK, 50)]
    hs = oracle.HostShard(6, 2500, cols, t0=T0, dt=SEC, seed=11)
    d = hs.desc
    ns, ng, nc = d.n_series, d.n_segments, d.n_columns
    sh = {
        "data": np.ctypeslib.as_array(d.data, shape=(d.data_len,)).copy(),
        "sids": np.ctypeslib.as_array(d.sids, shape=(ns,)).copy(),
        "series_seg_begin": np.ctypeslib.as_array(d.series_seg_begin, shape=(ns + 1,)).copy(),
        "seg_tmin": np.ctypeslib.as_array(d.seg_tmin, shape=(ng,)).copy(),
        "seg_tmax": np.ctypeslib.as_array(d.seg_tmax, shape=(ng,)).copy(),
        "col_types": np.array([d.columns[c].type for c in range(nc)], np.int32),
        "page_off": np.stack([np.ctypeslib.as_array(d.columns[c].page_off, shape=(ng,)) for c in range(nc)] + [np.ctypeslib.as_array(d.time_page_off, shape=(ng,))]).copy(),
        "page_len": np.stack([np.ctypeslib.as_array(d.columns[c].page_len, shape=(ng,)) for c in range(nc)] + [np.ctypeslib.as_array(d.time_page_len, shape=(ng,))]).copy(),
    }
    queries = [
        ("mean_max_1m", [(L.AGG_SUM, 0), (L.AGG_COUNT, 0), (L.AGG_MAX, 0)], 60 * SEC, L.GROUP_ALL),
        ("int_5m_per_series", [(L.AGG_COUNT, 1), (L.AGG_SUM, 1), (L.AGG_MIN, 1), (L.AGG_FIRST, 1), (L.AGG_LAST, 1)], 300 * SEC, L.GROUP_PER_SERIES),
        ("single_min_selector", [(L.AGG_MIN, 0)], 600 * SEC, L.GROUP_ALL),
    ]
    for qn, calls, ivl, gm in queries:
        ca = (L.Call * len(calls))(*calls)
        qd = L.QueryDesc(ivl, 0, T0 + 17 * SEC, T0 + 2400 * SEC, 1, len(calls), ca, 0, None, gm, ns if gm == L.GROUP_PER_SERIES else 1, None, 0, 0)
        r = oracle.scan(d, qd, threads=1)
        sh[f"q/{qn}/calls"] = np.array(calls, np.int32)
        sh[f"q/{qn}/params"] = np.array([ivl, T0 + 17 * SEC, T0 + 2400 * SEC, gm, r["n_groups"], r["n_buckets"], r["start"]], np.int64)
        for k, c in enumerate(r["cols"]):
            sh[f"q/{qn}/{k}/values"], sh[f"q/{qn}/{k}/valid"], sh[f"q/{qn}/{k}/times"] = c["values"], c["valid"], c["times"]
    sh["query_names"] = np.array([q[0] for q in queries])
    np.savez_compressed(os.path.join(HERE, "shard_v1.npz"), **sh)
    print("wrote", os.path.join(HERE, "pages_v1.npz"), os.path.join(HERE, "shard_v1.npz"))


if __name__ == "__main__":
    main()
