/*
 * agg_ops.cuh — the reducer algebra shared by every aggregate kernel.
 *
 *  Acc::add_row      per-record-window reduce kernels  engine/series_agg_func.gen.go:24-274 + lib/record/column_util.go:23-278
 *  series_merge      cross-record stitch (prevBuf (+) currBuf)  series_agg_func.gen.go:44-46,62-64,92-98,140-146,188,233
 *  group_update      tagset-level interval-record update  lib/record/reccord_functions.go:47-786
 *
 * Values travel as raw 64-bit cells: double bits, int64, or bool 0/1.  Types are influx.Field_Type_* (1 int, 3 float, 5 bool).
 */
#pragma once
#include <cstdint>

#include "../../include/ogpu.h"

namespace ogpu {

__device__ __forceinline__ double u2d(uint64_t u) { return __longlong_as_double((long long)u); }
__device__ __forceinline__ uint64_t d2u(double d) { return (uint64_t)__double_as_longlong(d); }

__device__ __forceinline__ bool v_lt(int type, uint64_t a, uint64_t b) {
    if (type == OG_TYPE_FLOAT) return u2d(a) < u2d(b);
    if (type == OG_TYPE_INT) return (int64_t)a < (int64_t)b;
    return a != b && !a; /* bool: false < true (series_agg_func.gen.go:124-130) */
}
__device__ __forceinline__ bool v_gt(int type, uint64_t a, uint64_t b) {
    if (type == OG_TYPE_FLOAT) return u2d(a) > u2d(b);
    if (type == OG_TYPE_INT) return (int64_t)a > (int64_t)b;
    return a != b && a;
}
__device__ __forceinline__ bool v_eq(int type, uint64_t a, uint64_t b) { return type == OG_TYPE_FLOAT ? u2d(a) == u2d(b) : a == b; }
__device__ __forceinline__ bool v_le(int type, uint64_t a, uint64_t b) {
    if (type == OG_TYPE_FLOAT) return u2d(a) <= u2d(b);
    if (type == OG_TYPE_INT) return (int64_t)a <= (int64_t)b;
    return a == b || !a;
}
__device__ __forceinline__ bool v_ge(int type, uint64_t a, uint64_t b) {
    if (type == OG_TYPE_FLOAT) return u2d(a) >= u2d(b);
    if (type == OG_TYPE_INT) return (int64_t)a >= (int64_t)b;
    return a == b || a;
}

struct Part { uint64_t v; int64_t t; uint32_t ok; }; /* one partial aggregate: value, carried row time, validity */

/* accumulate one surviving, non-null row into a per-record-window partial */
__device__ __forceinline__ void acc_row(int func, int type, Part &p, uint64_t v, int64_t t) {
    switch (func) {
    case OG_AGG_COUNT: p.v += 1; p.ok = 1; break;                          /* *CountReduce = ValidCount */
    case OG_AGG_SUM:                                                       /* floatSumReduce :48 / integerSumReduce :66: sequential from 0 */
        if (type == OG_TYPE_FLOAT) p.v = d2u(u2d(p.v) + u2d(v)); else p.v += v;
        p.ok = 1; break;
    case OG_AGG_MIN:                                                       /* minValue: first value seeds, strict > replaces */
        if (!p.ok || v_gt(type, p.v, v)) { p.v = v; p.t = t; }
        p.ok = 1; break;
    case OG_AGG_MAX:
        if (!p.ok || v_lt(type, p.v, v)) { p.v = v; p.t = t; }
        p.ok = 1; break;
    case OG_AGG_FIRST: if (!p.ok) { p.v = v; p.t = t; p.ok = 1; } break;  /* firstValue */
    default: p.v = v; p.t = t; p.ok = 1; break;                            /* lastValue */
    }
}
__device__ __forceinline__ Part part_empty() { Part p; p.v = 0; p.t = 0; p.ok = 0; return p; }

/* prev (+) curr for the same (series, window), prev being earlier in time */
__device__ __forceinline__ Part series_merge(int func, int type, const Part &prev, const Part &curr) {
    if (!curr.ok) return prev;
    if (!prev.ok) return curr;
    Part r = prev;
    switch (func) {
    case OG_AGG_COUNT: r.v = prev.v + curr.v; break;                                   /* integerCountMerge */
    case OG_AGG_SUM: r.v = type == OG_TYPE_FLOAT ? d2u(u2d(prev.v) + u2d(curr.v)) : prev.v + curr.v; break;
    case OG_AGG_MIN: if (v_lt(type, curr.v, prev.v)) r = curr; break;                  /* floatMinMerge: strict */
    case OG_AGG_MAX: if (v_gt(type, curr.v, prev.v)) r = curr; break;
    case OG_AGG_FIRST: break;                                                          /* floatFirstMerge keeps prev */
    default: r = curr; break;                                                          /* floatLastMerge takes curr */
    }
    return r;
}

/* tagset-level update of the dense interval record cell `a` with one series partial `p`.
 * a.t is the row's time column for single-call selectors (initialised to the window start by
 * BuildEmptyIntervalRec) or RecMeta.Times (initialised to 0) for multi-call first/last. */
__device__ __forceinline__ void group_update(int func, int type, bool multi, Part &a, const Part &p) {
    if (!p.ok) return; /* every Update* returns on a nil partial */
    switch (func) {
    case OG_AGG_COUNT: a.v = (uint64_t)((int64_t)p.v + (int64_t)a.v); a.ok = 1; return;           /* updateCountImpl :757 */
    case OG_AGG_SUM:                                                                              /* update{Integer,Float}SumImpl :712,:730 */
        a.v = type == OG_TYPE_FLOAT ? d2u(u2d(p.v) + u2d(a.v)) : p.v + a.v; a.ok = 1; return;
    case OG_AGG_MIN:
    case OG_AGG_MAX: {
        bool is_min = func == OG_AGG_MIN;
        if (multi) { /* update*Column{Min,Max}Impl :586-660 */
            if ((is_min ? v_le(type, a.v, p.v) : v_ge(type, a.v, p.v)) && a.ok) return;
            a.v = p.v; a.ok = 1; return;
        }
        if ((is_min ? v_lt(type, a.v, p.v) : v_gt(type, a.v, p.v)) && a.ok) return;                /* update*{Min,Max}Impl :429-560 */
        /* tie -> earlier time; float columns read the time column through FloatValue() (:487-488) */
        bool t_le = type == OG_TYPE_FLOAT ? (__longlong_as_double(a.t) <= __longlong_as_double(p.t)) : (a.t <= p.t);
        if (v_eq(type, a.v, p.v) && t_le && a.ok) return;
        a = p; a.ok = 1; return;                                                                   /* UpdateIntervalRecRow */
    }
    default: { /* first / last: update*FirstLastImp :47-227, update*ColumnFirstLastImp :229-420 (same control flow) */
        bool is_first = func == OG_AGG_FIRST;
        int64_t t1 = a.t, t2 = p.t;
        if (is_first ? (t1 > t2) : (t1 < t2)) { a = p; a.ok = 1; return; }
        if (a.ok && (is_first ? (t2 > t1) : (t2 < t1))) return;
        if (a.ok && v_ge(type, a.v, p.v)) return;
        a = p; a.ok = 1; return;
    }
    }
}

/* ProcessorOptions.Window bucket index of t relative to `start` (window start of the query's first bucket).
 * Valid for t >= start, which holds for every in-range row (start = Window(tmin).start). */
__device__ __forceinline__ uint32_t bucket_of(int64_t t, int64_t start, int64_t interval) {
    return (uint32_t)((uint64_t)(t - start) / (uint64_t)interval);
}

} // namespace ogpu
