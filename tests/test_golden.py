"""Committed fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py).

CPU: the oracle still decodes the fixed bytes to the fixed values and still encodes the values to the same bytes, and three
hand-derived vectors (from the wire-format description, SURVEY App. A / reference encoders) match the oracle's encoders.
GPU: the C-ABI decoders read the same fixed bytes to the same values; aggregates over the fixed shard equal the stored ones.
"""
import os
import struct

import numpy as np
import pytest

import oracle
from opengemini_b200 import _lib as L

HERE = os.path.dirname(os.path.abspath(__file__))
PAGES = np.load(os.path.join(HERE, "golden", "pages_v1.npz"))
SHARD = np.load(os.path.join(HERE, "golden", "shard_v1.npz"))
NAMES = [str(n) for n in PAGES["names"]]
T0, SEC = 1_700_000_000_000_000_000, 1_000_000_000


def _case(name):
    typ = int(PAGES[f"{name}/type"])
    vals = PAGES[f"{name}/values"]
    if typ == L.TYPE_FLOAT:
        vals = vals.view(np.float64)
    elif typ == L.TYPE_INT:
        vals = vals.view(np.int64)
    return typ, PAGES[f"{name}/page"], PAGES[f"{name}/time_page"], vals, PAGES[f"{name}/valid"].astype(bool), PAGES[f"{name}/times"]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_decodes_golden_pages(name):
    typ, page, tpage, vals, valid, times = _case(name)
    got_v, got_ok = oracle.field_page_decode(typ, page, cap=len(valid) + 8)
    assert np.array_equal(got_ok, valid)
    assert np.array_equal(got_v.view(np.uint8), vals.view(np.uint8))
    assert np.array_equal(oracle.time_page_decode(tpage, cap=len(times) + 8), times)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_encoders_reproduce_golden_bytes(name):
    typ, page, tpage, vals, valid, times = _case(name)
    cells = np.zeros(len(valid), vals.dtype)
    cells[valid] = vals
    again = oracle.field_page_encode(typ, cells, None if valid.all() else valid.astype(np.uint8))
    assert np.array_equal(again, page)
    assert np.array_equal(oracle.time_page_encode(times), tpage)


def test_hand_derived_vectors():
    """Bytes written down from the format description, not produced by the oracle."""
    # time column, 1000 rows, 1 s cadence: [BlockIntegerFull=32][u32 BE rows][const-delta tag 1<<4][u64 BE t0][uvarint delta][uvarint n-1]
    # (encoding.go:29-65, column_builder.go:446-486, timestamp.go:60-188; SURVEY §8d quotes the same 21 bytes)
    t = T0 + np.arange(1000, dtype=np.int64) * SEC
    want = bytes([32]) + struct.pack(">I", 1000) + bytes([0x10]) + struct.pack(">Q", T0) + bytes([0x80, 0x94, 0xEB, 0xDC, 0x03]) + bytes([0xE7, 0x07])
    assert bytes(oracle.time_page_encode(t)) == want and len(want) == 21
    # bool column without nulls, 9 rows 1,0,1,1,0,0,0,1,1: [BlockBooleanFull=33][u32 BE rows] + [1<<4][u32 BE n][bits MSB-first] (bool.go:40-61)
    b = np.array([1, 0, 1, 1, 0, 0, 0, 1, 1], np.uint8)
    want = bytes([33]) + struct.pack(">I", 9) + bytes([0x10]) + struct.pack(">I", 9) + bytes([0b10110001, 0b10000000])
    assert bytes(oracle.field_page_encode(L.TYPE_BOOL, b)) == want
    # one-row float column: [BlockFloat64One=17][8 bytes LE] (column_builder.go:488-502)
    want = bytes([17]) + struct.pack("<d", 42.5)
    assert bytes(oracle.field_page_encode(L.TYPE_FLOAT, np.array([42.5]))) == want
    # int const-delta block: [tag 1<<4][u64 BE zigzag(v0)][uvarint zigzag(delta)][uvarint n-1] after the Full header (int.go:91-134)
    v = (7 + 13 * np.arange(300)).astype(np.int64)
    want = bytes([32]) + struct.pack(">I", 300) + bytes([0x10]) + struct.pack(">Q", 14) + bytes([26]) + bytes([0xAB, 0x02])
    assert bytes(oracle.field_page_encode(L.TYPE_INT, v)) == want


def _golden_shard_desc():
    ex = {k: SHARD[k] for k in ("data", "sids", "series_seg_begin", "seg_tmin", "seg_tmax", "col_types", "page_off", "page_len")}
    ex["data"] = np.concatenate([ex["data"], np.zeros(1024, np.uint8)])[:len(ex["data"])].copy()
    return oracle.shard_desc_from_export(ex), ex


def test_oracle_reproduces_golden_aggregates():
    d, keep = _golden_shard_desc()
    for qn in [str(x) for x in SHARD["query_names"]]:
        calls = [tuple(int(v) for v in c) for c in SHARD[f"q/{qn}/calls"]]
        ivl, tmin, tmax, gm, ng, nb, start = (int(v) for v in SHARD[f"q/{qn}/params"])
        ca = (L.Call * len(calls))(*calls)
        qd = L.QueryDesc(ivl, 0, tmin, tmax, 1, len(calls), ca, 0, None, gm, d.n_series if gm == L.GROUP_PER_SERIES else 1, None, 0, 0)
        r = oracle.scan(d, qd, threads=1)  # one cursor = the reference's fold order; more workers re-associate float sums
        assert (r["n_groups"], r["n_buckets"], r["start"]) == (ng, nb, start)
        for k, c in enumerate(r["cols"]):
            ok = SHARD[f"q/{qn}/{k}/valid"].astype(bool)
            assert np.array_equal(c["valid"].astype(bool), ok), (qn, k)
            assert np.array_equal(c["values"][ok], SHARD[f"q/{qn}/{k}/values"][ok]), (qn, k)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_decodes_golden_pages(name):
    from opengemini_b200 import Shard
    typ, page, tpage, vals, valid, times = _case(name)
    data = np.concatenate([page, tpage])
    sh = Shard.open(data, [1], [0, 1], [int(times[0])], [int(times[-1])], [("v", typ, [0], [page.size])], [page.size], [tpage.size])
    rec = sh.decode_segment(0)
    assert np.array_equal(rec["times"], times)
    col = rec["cols"][0]
    assert np.array_equal(col["valid"], valid)
    assert np.array_equal(np.ascontiguousarray(col["values"]).view(np.uint8), vals.view(np.uint8))
    sh.close()


@pytest.mark.gpu
def test_gpu_reproduces_golden_aggregates():
    from opengemini_b200 import AggQuery, Shard
    d, keep = _golden_shard_desc()
    sh = Shard.open_desc(d, keepalive=keep)
    names = {L.AGG_COUNT: "count", L.AGG_SUM: "sum", L.AGG_MIN: "min", L.AGG_MAX: "max", L.AGG_FIRST: "first", L.AGG_LAST: "last"}
    for qn in [str(x) for x in SHARD["query_names"]]:
        calls = [(names[int(f)], int(c)) for f, c in SHARD[f"q/{qn}/calls"]]
        ivl, tmin, tmax, gm, ng, nb, start = (int(v) for v in SHARD[f"q/{qn}/params"])
        # the fixtures hold the reference's summation order: OG_Q_STRICT_ORDER (the default one-tagset order folds 32 series first)
        q = AggQuery(sh, calls, ivl, tmin, tmax, group={L.GROUP_ALL: "all", L.GROUP_PER_SERIES: "series"}[gm], flags=L.Q_STRICT_ORDER).run()
        got = q.dense_host()
        assert (got["n_groups"], got["n_buckets"], got["start"]) == (ng, nb, start)
        for k in range(len(calls)):
            ok = SHARD[f"q/{qn}/{k}/valid"].astype(bool)
            assert np.array_equal(got["cols"][k]["valid"].astype(bool), ok), (qn, k)
            assert np.array_equal(got["cols"][k]["values"].view(np.uint64)[ok], SHARD[f"q/{qn}/{k}/values"][ok]), (qn, k)  # bit-exact incl. float sums
            if got["cols"][k]["times"] is not None:
                assert np.array_equal(got["cols"][k]["times"][ok], SHARD[f"q/{qn}/{k}/times"][ok]), (qn, k)
        q.close()
    sh.close()
