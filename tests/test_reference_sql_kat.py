"""Known answers from the reference's SQL-level tests, replayed through the scan/aggregate path.

tests/server_test.go of the reference writes eight points (`intmany` / `floatmany`: eight series host=server01..08, one point each,
10 s apart from 2000-01-01T00:00:00Z, values 2,4,4,4,5,5,7,9) and states the expected result of aggregate queries over them:
  :2319-2340  MEAN ... GROUP BY time(10m) = 5; FIRST = 2 @ 00:00:00; LAST = 9 @ 00:01:10; SPREAD (= max - min) = 7
  :2480-2530  max GROUP BY time(10s) and time(30s), min / first / last GROUP BY time(15s) over [00:00:00, 00:01:14]
The same rows are encoded as one-row TSSP pages (one series per host, as the writes create), then both the CPU oracle (always) and
the GPU library (-m gpu) must return those published values.  This pins the oracle's whole scan pipeline — page decode, FilterByTime,
window alignment, per-series reduce, tagset merge — to answers the reference itself holds, not just to its own encoders.
"""
import numpy as np
import pytest

import oracle
from opengemini_b200 import _lib as L

SEC = 1_000_000_000
T2000 = 946_684_800 * SEC  # 2000-01-01T00:00:00Z
VALUES = [2.0, 4.0, 4.0, 4.0, 5.0, 5.0, 7.0, 9.0]


def _export():
    pages, tpages, times = [], [], []
    for k, v in enumerate(VALUES):
        t = np.array([T2000 + 10 * k * SEC], np.int64)
        pages.append(oracle.field_page_encode(L.TYPE_FLOAT, np.array([v])))
        tpages.append(oracle.time_page_encode(t))
        times.append(int(t[0]))
    blob, offs, lens, pos = [], [], [], 0
    for p in pages + tpages:
        offs.append(pos); lens.append(p.size); blob.append(p); pos += p.size
    n = len(VALUES)
    return dict(data=np.concatenate(blob + [np.zeros(1024, np.uint8)])[:pos].copy(), sids=np.arange(1, n + 1, dtype=np.uint64),
                series_seg_begin=np.arange(n + 1, dtype=np.uint32), seg_tmin=np.array(times, np.int64), seg_tmax=np.array(times, np.int64),
                col_types=np.array([L.TYPE_FLOAT], np.int32), page_off=np.array([offs[:n], offs[n:]], np.uint64),
                page_len=np.array([lens[:n], lens[n:]], np.uint32))


# (calls, interval, tmin, tmax, expected per call: list over buckets of value or None)
RANGE = (T2000, T2000 + 74 * SEC)
CASES = [
    ("max by 10s", [("max", 0)], 10 * SEC, RANGE, [[2, 4, 4, 4, 5, 5, 7, 9]]),
    ("max by 30s", [("max", 0)], 30 * SEC, RANGE, [[4, 5, 9]]),
    ("min by 15s", [("min", 0)], 15 * SEC, RANGE, [[2, 4, 4, 5, 7]]),
    ("first by 15s", [("first", 0)], 15 * SEC, RANGE, [[2, 4, 4, 5, 7]]),
    ("last by 15s", [("last", 0)], 15 * SEC, RANGE, [[4, 4, 5, 5, 9]]),
    ("mean by 10m", [("sum", 0), ("count", 0)], 600 * SEC, (T2000, T2000 + 120 * SEC - 1), [[40], [8]]),
    ("spread", [("max", 0), ("min", 0)], 0, (T2000 - 5 * SEC, T2000 + 500 * SEC), [[9], [2]]),
]
FUNCS = {"count": L.AGG_COUNT, "sum": L.AGG_SUM, "min": L.AGG_MIN, "max": L.AGG_MAX, "first": L.AGG_FIRST, "last": L.AGG_LAST}


def _check(result, calls, expected, label):
    for k, ((f, _c), want) in enumerate(zip(calls, expected)):
        col = result["cols"][k]
        assert len(want) == result["n_buckets"], (label, result["n_buckets"])
        for b, w in enumerate(want):
            assert bool(col["valid"][b]) == (w is not None), (label, f, b)
            if w is None:
                continue
            got = int(col["values"][b]) if f == "count" else float(np.array(col["values"][b:b + 1]).view(np.float64)[0])
            assert got == w, (label, f, b, got, w)


def _oracle_scan(ex, calls, interval, tr):
    d = oracle.shard_desc_from_export(ex)
    ca = (L.Call * len(calls))(*[(FUNCS[f], c) for f, c in calls])
    qd = L.QueryDesc(interval, 0, tr[0], tr[1], 1, len(calls), ca, 0, None, L.GROUP_ALL, 1, None, 0, 0)
    return oracle.scan(d, qd, threads=1)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_returns_the_references_published_answers(case):
    label, calls, interval, tr, expected = case
    _check(_oracle_scan(_export(), calls, interval, tr), calls, expected, label)


def test_oracle_selector_times_are_the_points_times():
    """FIRST(value) -> 2 at 2000-01-01T00:00:00Z, LAST(value) -> 9 at 2000-01-01T00:01:10Z (server_test.go:2325-2340)."""
    ex = _export()
    r = _oracle_scan(ex, [("first", 0)], 0, (T2000 - SEC, T2000 + 1000 * SEC))
    assert r["cols"][0]["values"].view(np.float64)[0] == 2.0 and r["cols"][0]["times"][0] == T2000
    r = _oracle_scan(ex, [("last", 0)], 0, (T2000 - SEC, T2000 + 1000 * SEC))
    assert r["cols"][0]["values"].view(np.float64)[0] == 9.0 and r["cols"][0]["times"][0] == T2000 + 70 * SEC


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_gpu_returns_the_references_published_answers(case):
    from opengemini_b200 import AggQuery, Shard
    label, calls, interval, tr, expected = case
    ex = _export()
    sh = Shard.open(ex["data"], ex["sids"], ex["series_seg_begin"], ex["seg_tmin"], ex["seg_tmax"],
                    [("value", L.TYPE_FLOAT, ex["page_off"][0], ex["page_len"][0])], ex["page_off"][1], ex["page_len"][1])
    q = AggQuery(sh, calls, interval, tr[0], tr[1]).run()
    got = q.dense_host()
    _check(dict(n_buckets=got["n_buckets"], cols=[dict(values=np.ascontiguousarray(c["values"]).view(np.uint64), valid=c["valid"]) for c in got["cols"]]),
           calls, expected, label)
    q.close()
    if label == "spread":
        for f, v, t in (("first", 2.0, T2000), ("last", 9.0, T2000 + 70 * SEC)):
            q = AggQuery(sh, [(f, 0)], 0, T2000 - SEC, T2000 + 1000 * SEC).run()
            g = q.dense_host()
            assert g["cols"][0]["values"][0] == v and g["cols"][0]["times"][0] == t
            q.close()
    sh.close()


def test_oracle_descending_scan_keeps_the_aggregates():
    """ORDER BY time DESC changes the order rows are emitted in, not the windows' aggregates (server_test.go:2585-2590)."""
    d = oracle.shard_desc_from_export(_export())
    ca = (L.Call * 1)((L.AGG_MAX, 0))
    res = []
    for asc in (1, 0):
        qd = L.QueryDesc(10 * SEC, 0, T2000, T2000 + 60 * SEC, asc, 1, ca, 0, None, L.GROUP_ALL, 1, None, 0, 0)
        res.append(oracle.scan(d, qd, threads=1))
    assert np.array_equal(res[0]["cols"][0]["values"], res[1]["cols"][0]["values"])
    assert list(res[1]["cols"][0]["values"].view(np.float64)) == [2, 4, 4, 4, 5, 5, 7]


@pytest.mark.gpu
def test_gpu_order_by_time_desc():
    """`SELECT max(value) ... group by time(10s) order by time desc` -> 7,5,5,4,4,4,2 from 00:01:00 down to 00:00:00 (server_test.go:2585-2590)."""
    from opengemini_b200 import AggQuery, Shard
    ex = _export()
    sh = Shard.open(ex["data"], ex["sids"], ex["series_seg_begin"], ex["seg_tmin"], ex["seg_tmax"],
                    [("value", L.TYPE_FLOAT, ex["page_off"][0], ex["page_len"][0])], ex["page_off"][1], ex["page_len"][1])
    for chunk in (1024, 3):
        q = AggQuery(sh, [("max", 0), ("count", 0)], 10 * SEC, T2000, T2000 + 60 * SEC, ascending=False, chunk_size=chunk).run()
        vals, times = [], []
        for rec in q.records():
            assert rec["rows"] <= chunk
            vals += list(rec["cols"][0]["values"]); times += list(rec["times"])
        assert vals == [7, 5, 5, 4, 4, 4, 2]
        assert times == [T2000 + k * 10 * SEC for k in range(6, -1, -1)]
        q.close()
    sh.close()
