"""The CPU-baseline leg (oracle/fast_scan.cpp: the reference's batch Gorilla decode with a 64-bit cached bit reader) must give
exactly what the checker's pull loop gives: bitwise with one worker, 1e-12 on float sums with several (worker partials merge)."""
import time

import numpy as np
import pytest

import oracle
from opengemini_b200 import _lib as L

T0, SEC = 1_700_000_000_000_000_000, 1_000_000_000


def _qd(calls, interval, tmin, tmax, offset=0):
    ca = (L.Call * len(calls))()
    for i, (f, c) in enumerate(calls):
        ca[i].func, ca[i].column = f, c
    d = L.QueryDesc()
    d.interval, d.offset, d.tmin, d.tmax, d.ascending = interval, offset, tmin, tmax, 1
    d.n_calls, d.calls, d.n_filter, d.group_mode, d.chunk_size = len(calls), ca, 0, L.GROUP_ALL, 1024
    d._keep = ca
    return d


@pytest.mark.parametrize("dist", [L.SYNTH_F_HI, L.SYNTH_F_LO])
def test_fast_scan_equals_checker(dist):
    rows = 30_000
    hs = oracle.HostShard(37, rows, [(L.TYPE_FLOAT, dist, 0)], t0=T0, dt=SEC, seed=5)
    cases = [([(L.AGG_SUM, 0), (L.AGG_COUNT, 0), (L.AGG_MAX, 0)], 60 * SEC, T0, T0 + (rows - 1) * SEC, 0),
             ([(L.AGG_SUM, 0), (L.AGG_COUNT, 0), (L.AGG_MIN, 0), (L.AGG_MAX, 0)], 7 * SEC, T0 + 1234 * SEC + 5, T0 + 20_000 * SEC, 3 * SEC),
             ([(L.AGG_SUM, 0)], 3600 * SEC, T0 - 100 * SEC, T0 + 10 * rows * SEC, 0),
             ([(L.AGG_COUNT, 0)], 61 * SEC, T0 + 999 * SEC, T0 + 1001 * SEC, -13 * SEC)]
    for calls, iv, tmin, tmax, off in cases:
        qd = _qd(calls, iv, tmin, tmax, off)
        ref = oracle.scan(hs.desc, qd, threads=1)
        for threads in (1, 4):
            got = oracle.scan(hs.desc, qd, threads=threads, fast=True)
            assert (got["n_buckets"], got["start"], got["rows_decoded"], got["page_bytes"]) == (ref["n_buckets"], ref["start"], ref["rows_decoded"], ref["page_bytes"])
            for k, (f, _c) in enumerate(calls):
                rv = ref["cols"][k]["valid"].astype(bool)
                assert np.array_equal(got["cols"][k]["valid"].astype(bool), rv)
                g, r = got["cols"][k]["values"][rv], ref["cols"][k]["values"][rv]
                if f == L.AGG_SUM and threads > 1:
                    assert np.allclose(g.view(np.float64), r.view(np.float64), rtol=1e-12, atol=0)
                else:
                    assert np.array_equal(g, r), (f, threads)


def test_fast_scan_rejects_what_it_does_not_restate():
    hs = oracle.HostShard(3, 2000, [(L.TYPE_FLOAT, L.SYNTH_F_HI, 100), (L.TYPE_INT, L.SYNTH_INT_WALK, 0)], t0=T0, dt=SEC, seed=1)
    for calls in ([(L.AGG_MAX, 0)], [(L.AGG_SUM, 1)], [(L.AGG_SUM, 0)]):  # selector with time; int column; pages with nulls
        with pytest.raises(ValueError):
            oracle.scan(hs.desc, _qd(calls, 60 * SEC, T0, T0 + 1999 * SEC), fast=True)


def test_fast_scan_speed_is_in_the_range_the_reference_reports():
    """batch_float.go:303-306 reports 320-340 MB/s of decoded float64 per core (a 2016 laptop): the baseline leg should be at
    least in that league on this host, and clearly faster than the bit-serial checker."""
    rows = 200_000
    hs = oracle.HostShard(8, rows, [(L.TYPE_FLOAT, L.SYNTH_F_HI, 0)], t0=T0, dt=SEC, seed=9)
    qd = _qd([(L.AGG_SUM, 0), (L.AGG_COUNT, 0), (L.AGG_MAX, 0)], 60 * SEC, T0, T0 + (rows - 1) * SEC)
    oracle.scan(hs.desc, qd, fast=True)
    t = time.perf_counter(); oracle.scan(hs.desc, qd, fast=True); fast = time.perf_counter() - t
    t = time.perf_counter(); oracle.scan(hs.desc, qd); slow = time.perf_counter() - t
    mbs = 8 * rows * 8 / fast / 1e6
    print(f"fast leg {mbs:.0f} MB/s decoded per thread; checker {8 * rows * 8 / slow / 1e6:.0f} MB/s")
    assert fast < slow and mbs > 150
