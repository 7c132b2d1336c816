"""ctypes access to the CPU oracle (oracle/_build/libogoracle.so) — test infrastructure only.

Nothing under opengemini_b200/ imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from opengemini_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "_build", "libogoracle.so")
KAT = os.path.join(ORACLE_DIR, "_build", "kat_tests")


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        for name in ("ogo_float_encode", "ogo_float_decode", "ogo_gorilla_encode", "ogo_int_encode", "ogo_int_decode", "ogo_time_encode",
                     "ogo_time_decode", "ogo_bool_encode", "ogo_bool_decode", "ogo_snappy_roundtrip", "ogo_field_page_encode",
                     "ogo_field_page_decode", "ogo_time_page_encode", "ogo_time_page_decode", "ogo_s8b_encode"):
            getattr(_lib, name).restype = C.c_long
        _lib.ogo_synth_build.restype = C.c_void_p
        _lib.ogo_synth_build.argtypes = [C.POINTER(L.SynthDesc), C.POINTER(C.c_int)]
        _lib.ogo_shard_desc.argtypes = [C.c_void_p, C.POINTER(L.ShardDesc)]
        _lib.ogo_shard_free.argtypes = [C.c_void_p]
        _lib.ogo_scan.restype = C.c_void_p
        _lib.ogo_fast_scan.restype = C.c_void_p
        _lib.ogo_fast_scan.argtypes = [C.POINTER(L.ShardDesc), C.POINTER(L.QueryDesc), C.c_int, C.c_uint32, C.c_uint32, C.POINTER(C.c_int)]
        _lib.ogo_scan.argtypes = [C.POINTER(L.ShardDesc), C.POINTER(L.QueryDesc), C.c_int, C.c_uint32, C.c_uint32, C.POINTER(C.c_int)]
        _lib.ogo_scan_dims.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        _lib.ogo_scan_col.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.ogo_scan_free.argtypes = [C.c_void_p]
        _lib.ogo_window.argtypes = [C.c_int64] * 5 + [L.i64p, L.i64p]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _enc(fn, arr, cap):
    out = np.empty(cap, np.uint8)
    n = getattr(lib(), fn)(_p(arr), C.c_size_t(arr.size), _p(out), C.c_size_t(cap))
    if n < 0:
        raise ValueError(f"{fn} rc={n}")
    return out[:n].copy()


def _dec(fn, page, dtype, cap):
    page = np.ascontiguousarray(page, np.uint8)
    out = np.empty(cap, dtype)
    n = getattr(lib(), fn)(_p(page), C.c_size_t(page.size), _p(out), C.c_size_t(cap))
    if n < 0:
        raise ValueError(f"{fn} rc={n}")
    return out[:n].copy()


def float_encode(v):
    v = np.ascontiguousarray(v, np.float64)
    return _enc("ogo_float_encode", v, v.size * 10 + 64)


def float_decode(b, cap=100000):
    return _dec("ogo_float_decode", b, np.float64, cap)


def int_encode(v):
    v = np.ascontiguousarray(v, np.int64)
    return _enc("ogo_int_encode", v, v.size * 10 + 64)


def int_decode(b, cap=100000):
    return _dec("ogo_int_decode", b, np.int64, cap)


def time_encode(v):
    v = np.ascontiguousarray(v, np.int64)
    return _enc("ogo_time_encode", v, v.size * 10 + 64)


def time_decode(b, cap=100000):
    return _dec("ogo_time_decode", b, np.int64, cap)


def bool_encode(v):
    v = np.ascontiguousarray(v, np.uint8)
    return _enc("ogo_bool_encode", v, v.size + 64)


def bool_decode(b, cap=100000):
    return _dec("ogo_bool_decode", b, np.uint8, cap)


def field_page_encode(typ, cells, valid=None):
    """cells: per-row array (float64/int64/uint8); valid: per-row 0/1 or None."""
    cells = np.ascontiguousarray(cells)
    rows = cells.size
    out = np.empty(rows * 10 + 256, np.uint8)
    vp = _p(np.ascontiguousarray(valid, np.uint8)) if valid is not None else None
    vkeep = np.ascontiguousarray(valid, np.uint8) if valid is not None else None
    n = lib().ogo_field_page_encode(typ, _p(cells), _p(vkeep) if vkeep is not None else None, C.c_size_t(rows), _p(out), C.c_size_t(out.size))
    del vp
    if n < 0:
        raise ValueError(f"ogo_field_page_encode rc={n}")
    return out[:n].copy()


def field_page_decode(typ, page, cap=2000):
    page = np.ascontiguousarray(page, np.uint8)
    vals = np.empty(cap, np.uint8 if typ == L.TYPE_BOOL else np.uint64)
    valid = np.empty(cap, np.uint8)
    nil = C.c_int()
    n = lib().ogo_field_page_decode(typ, _p(page), C.c_size_t(page.size), _p(vals), C.c_size_t(cap), _p(valid), C.c_size_t(cap), C.byref(nil))
    if n < 0:
        raise ValueError(f"ogo_field_page_decode rc={n}")
    nv = n - nil.value
    v = vals[:nv].copy()
    if typ == L.TYPE_FLOAT:
        v = v.view(np.float64)
    elif typ == L.TYPE_INT:
        v = v.view(np.int64)
    return v, valid[:n].astype(bool)


def time_page_encode(t):
    t = np.ascontiguousarray(t, np.int64)
    return _enc("ogo_time_page_encode", t, t.size * 10 + 64)


def time_page_decode(page, cap=2000):
    return _dec("ogo_time_page_decode", page, np.int64, cap)


def window(interval, offset, tmin, tmax, t):
    s, e = C.c_int64(), C.c_int64()
    lib().ogo_window(interval, offset, tmin, tmax, t, C.byref(s), C.byref(e))
    return s.value, e.value


class HostShard:
    """Synthetic shard built on the host with the oracle's restated encoders."""

    def __init__(self, n_series, rows_per_series, columns, t0=1_700_000_000_000_000_000, dt=1_000_000_000, seed=1, rows_per_segment=1000,
                 threads=1, series_base=0):
        cols = (L.SynthColumn * len(columns))()
        for i, (t, dist, npm) in enumerate(columns):
            cols[i].type, cols[i].dist, cols[i].null_permille = t, dist, npm
        self._cols = cols
        d = L.SynthDesc(n_series, rows_per_series, rows_per_segment, t0, dt, seed, len(columns), cols, series_base)
        st = C.c_int()
        lib().ogo_synth_build_mt.restype = C.c_void_p
        lib().ogo_synth_build_mt.argtypes = [C.POINTER(L.SynthDesc), C.c_int, C.POINTER(C.c_int)]
        self.h = lib().ogo_synth_build_mt(C.byref(d), threads, C.byref(st))
        if not self.h:
            raise ValueError(f"ogo_synth_build rc={st.value}")
        self.desc = L.ShardDesc()
        lib().ogo_shard_desc(self.h, C.byref(self.desc))

    def page(self, col, seg):
        """col == n_columns selects the time column."""
        d = self.desc
        if col == d.n_columns:
            off, ln = d.time_page_off[seg], d.time_page_len[seg]
        else:
            off, ln = d.columns[col].page_off[seg], d.columns[col].page_len[seg]
        return np.ctypeslib.as_array(d.data, shape=(d.data_len,))[off:off + ln].copy()

    def __del__(self):
        if getattr(self, "h", None):
            lib().ogo_shard_free(self.h)
            self.h = None


def shard_desc_from_export(ex):
    """Build an L.ShardDesc over arrays exported from a device shard (Shard.export())."""
    nc = ex["col_types"].size
    cds = (L.ColumnDesc * max(1, nc))()
    keep = [ex]
    names = []
    for c in range(nc):
        names.append(f"f{c}".encode())
        cds[c].name = names[-1]
        cds[c].type = int(ex["col_types"][c])
        po = np.ascontiguousarray(ex["page_off"][c])
        pl = np.ascontiguousarray(ex["page_len"][c])
        keep += [po, pl]
        cds[c].page_off = po.ctypes.data_as(L.u64p)
        cds[c].page_len = pl.ctypes.data_as(L.u32p)
    tpo = np.ascontiguousarray(ex["page_off"][nc])
    tpl = np.ascontiguousarray(ex["page_len"][nc])
    keep += [tpo, tpl, cds, names]
    d = L.ShardDesc()
    d.data = ex["data"].ctypes.data_as(L.u8p)
    d.data_len = ex["data"].size
    d.n_series = ex["sids"].size
    d.sids = ex["sids"].ctypes.data_as(L.u64p)
    d.series_seg_begin = ex["series_seg_begin"].ctypes.data_as(L.u32p)
    d.n_segments = ex["seg_tmin"].size
    d.seg_tmin = ex["seg_tmin"].ctypes.data_as(L.i64p)
    d.seg_tmax = ex["seg_tmax"].ctypes.data_as(L.i64p)
    d.n_columns = nc
    d.columns = cds
    d.time_page_off = tpo.ctypes.data_as(L.u64p)
    d.time_page_len = tpl.ctypes.data_as(L.u32p)
    d._keep = keep
    return d


def scan(shard_desc, query_desc, threads=1, s0=0, s1=0xFFFFFFFF, fast=False):
    """Run the reference-structured CPU pipeline. Returns dict like AggQuery.dense_host().
    fast=True: the CPU-baseline leg (oracle/fast_scan.cpp: batch Gorilla decode with a 64-bit cached bit reader) for the
    headline query shape; use it on HostShard descriptors (it may read a few bytes past the last page)."""
    st = C.c_int()
    fn = lib().ogo_fast_scan if fast else lib().ogo_scan
    h = fn(C.byref(shard_desc), C.byref(query_desc), threads, s0, s1, C.byref(st))
    if not h:
        raise ValueError(f"{'ogo_fast_scan' if fast else 'ogo_scan'} rc={st.value}")
    try:
        ng, nb = C.c_uint32(), C.c_uint32()
        start, iv = C.c_int64(), C.c_int64()
        rows, segs, by = C.c_uint64(), C.c_uint64(), C.c_uint64()
        lib().ogo_scan_dims(h, C.byref(ng), C.byref(nb), C.byref(start), C.byref(iv), C.byref(rows), C.byref(segs), C.byref(by))
        n = ng.value * nb.value
        cols = []
        for k in range(query_desc.n_calls):
            pv, pk, pt = C.c_void_p(), C.c_void_p(), C.c_void_p()
            lib().ogo_scan_col(h, k, C.byref(pv), C.byref(pk), C.byref(pt))
            vals = np.ctypeslib.as_array(C.cast(pv, L.u64p), shape=(n,)).copy()
            valid = np.ctypeslib.as_array(C.cast(pk, L.u8p), shape=(n,)).copy()
            times = np.ctypeslib.as_array(C.cast(pt, L.i64p), shape=(n,)).copy()
            cols.append(dict(values=vals, valid=valid, times=times))
        return dict(n_groups=ng.value, n_buckets=nb.value, start=start.value, interval=iv.value, cols=cols,
                    rows_decoded=rows.value, segments=segs.value, page_bytes=by.value)
    finally:
        lib().ogo_scan_free(h)
