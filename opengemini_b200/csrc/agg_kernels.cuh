/*
 * agg_kernels.cuh — the scan/aggregate kernels.
 *
 * Data flow of one query over one chunk of series (all arrays in HBM):
 *
 *   pages ──k_fused_segment──┐                       (thread per segment: decode + window partials, nothing materialised)
 *   pages ──k_decode_tile──> tile ──k_filter_tile──> keep ──k_window_reduce──┤   (generic path: any column mix, WHERE filters)
 *                                                                            v
 *                         interior windows -> cells[series][bucket]   first/last window of a segment -> edges[2*seg]
 *                                                                            │
 *                         k_fix_edges: ordered stitch of windows that span segments (prevBuf/currBuf, series_agg_reducer.gen.go:228-266)
 *                                                                            v
 *                         k_merge_groups: per (group, bucket) sequential fold over the group's series in shard order
 *                                         (AggTagSetCursor.RecordInit agg_tagset_cursor.go:1069) -> dense interval record
 */
#pragma once
#include "agg_ops.cuh"
#include "decode.cuh"
#include "internal.h"

namespace ogpu {

#define OG_NO_BUCKET 0xFFFFFFFFu

struct DirP { /* device directory */
    const uint8_t *data;
    const uint64_t *page_off; /* [(n_columns+1)*n_segments] */
    const uint32_t *page_len;
    const uint32_t *seg_series, *seg_rows, *series_seg_begin;
    const int64_t *seg_tmin, *seg_tmax;
    uint32_t n_segments, n_columns;
};

struct ChunkP { /* one chunk of whole series */
    uint32_t series_begin, series_end; /* global series range */
    uint32_t seg_begin, seg_end;       /* global segment range (contiguous) */
    Tri cells[OG_MAX_CALLS];           /* [ cell_idx(ch, series, b) ]: bucket-major, series contiguous — a warp whose lanes are
                                          32 consecutive series writes one bucket as one 256-byte run */
    uint32_t cell_sb;                  /* cells per bucket row = series of the chunk rounded up to 32 */
    Tri edges[OG_MAX_CALLS];           /* [ 2 * (seg - seg_begin) + {0 head, 1 tail} ] */
    uint32_t *edge_bucket;             /* [ 2 * (seg - seg_begin) ] OG_NO_BUCKET = absent */
    int *err;                          /* [0] first error code, [1] segment */
};

__device__ __forceinline__ size_t cell_idx(const ChunkP &ch, uint32_t series, uint32_t b) { return (size_t)b * ch.cell_sb + (series - ch.series_begin); }

__device__ __forceinline__ void report_err(int *err, int code, uint32_t seg) {
    if (atomicCAS(&err[0], 0, code) == 0) err[1] = (int)seg;
}

__device__ __forceinline__ void store_part(const Tri &a, size_t i, const Part &p) {
    a.val[i] = p.v; a.ok[i] = (uint8_t)p.ok;
    if (a.tim) a.tim[i] = p.t;
}
__device__ __forceinline__ Part load_part(const Tri &a, size_t i) {
    Part p; p.ok = a.ok[i]; p.v = a.val[i]; p.t = a.tim ? a.tim[i] : 0; return p;
}

/* ------------------------------------------------------------------------------------------------------------
 * shard open: row counts, codec support and framing validation (one thread per segment)
 * ------------------------------------------------------------------------------------------------------------ */
__global__ void k_validate(DirP d, const int32_t *col_types, uint32_t *seg_rows, unsigned long long *totals /*[0]=rows [1]=page bytes*/,
                           uint32_t *max_rows, int *err) {
    uint32_t seg = blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= d.n_segments) return;
    size_t ti = (size_t)d.n_columns * d.n_segments + seg;
    TimeDesc t;
    int rc = parse_time_page(d.data + d.page_off[ti], d.page_len[ti], t);
    if (rc != D_OK) { report_err(err, rc, seg); seg_rows[seg] = 0; return; }
    seg_rows[seg] = t.rows;
    unsigned long long bytes = d.page_len[ti];
    for (uint32_t c = 0; c < d.n_columns; c++) {
        size_t pi = (size_t)c * d.n_segments + seg;
        uint32_t len = d.page_len[pi];
        if (len == 0) continue;
        bytes += len;
        PageHdr h;
        rc = parse_field_header(d.data + d.page_off[pi], len, col_types[c], t.rows, h);
        if (rc == D_OK && h.rows != t.rows) rc = D_CORRUPT;
        if (rc == D_OK && !h.one_row && h.nil_count < h.rows) {
            if (h.block_len < 1) rc = D_CORRUPT;
            else {
                int tag = __ldg(h.block) >> 4, ty = col_types[c];
                if (ty == OG_TYPE_FLOAT) rc = (tag == 0 || tag == 3 || tag == 4 || tag == 5) ? D_OK : (tag == 1 || tag == 2 || tag == 6) ? D_UNSUPPORTED : D_CORRUPT;
                else if (ty == OG_TYPE_INT) rc = (tag == 1 || tag == 2 || tag == 4) ? D_OK : tag == 3 ? D_UNSUPPORTED : D_CORRUPT;
                else if (ty == OG_TYPE_BOOL) rc = tag == 1 ? D_OK : D_CORRUPT;
                else rc = D_UNSUPPORTED;
            }
        }
        if (rc != D_OK) { report_err(err, rc, seg); return; }
    }
    atomicAdd(&totals[0], (unsigned long long)t.rows);
    atomicAdd(&totals[1], bytes);
    atomicMax(max_rows, t.rows);
}

__global__ void k_fill_seg_series(const uint32_t *series_seg_begin, uint32_t n_series, uint32_t *seg_series) {
    uint32_t s = blockIdx.x;
    if (s >= n_series) return;
    for (uint32_t g = series_seg_begin[s] + threadIdx.x; g < series_seg_begin[s + 1]; g += blockDim.x) seg_series[g] = s;
}

/* ------------------------------------------------------------------------------------------------------------
 * generic path, step 1: materialise a tile of segments (thread per page; grid.y = column slot, last slot = time)
 *   vals[slot][(seg-tile_begin)*R + row]  expanded to one cell per row, okb = validity byte per row
 * ------------------------------------------------------------------------------------------------------------ */
struct TileP {
    uint32_t tile_begin, tile_end, R; /* R = rows reserved per segment */
    uint32_t S;                       /* segments per row of the tile (tile size rounded up to 32): cell (segment sl, row r)
                                         lives at r*S + sl, so threads that own consecutive segments and walk their rows in
                                         step read and write consecutive addresses */
    uint64_t *vals[OG_MAX_COLS];
    uint8_t *okb[OG_MAX_COLS];
    int64_t *times;
    uint8_t *keep;
};

struct ExpandEmit {
    uint64_t *out; uint8_t *okb; const PageHdr *h; uint32_t row; size_t stride;
    __device__ __forceinline__ void operator()(uint32_t, uint64_t bits) {
        while (row < h->rows && !hdr_row_valid(*h, row)) { out[row * stride] = 0; okb[row * stride] = 0; row++; }
        if (row < h->rows) { out[row * stride] = bits; okb[row * stride] = 1; row++; }
    }
};
struct TimeStore { int64_t *out; size_t stride; __device__ __forceinline__ void operator()(uint32_t i, int64_t t) { out[i * stride] = t; } };

__global__ void k_decode_tile(DirP d, QueryP q, TileP tp, int *err) {
    uint32_t seg = tp.tile_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= tp.tile_end) return;
    uint32_t slot = blockIdx.y;
    const size_t base = seg - tp.tile_begin, S = tp.S;
    uint32_t rows = d.seg_rows[seg];
    if (slot == q.n_cols) {
        size_t ti = (size_t)d.n_columns * d.n_segments + seg;
        TimeDesc t;
        int rc = parse_time_page(d.data + d.page_off[ti], d.page_len[ti], t);
        if (rc == D_OK) { TimeStore ts{tp.times + base, S}; rc = decode_time_values(t, ts); }
        if (rc != D_OK) report_err(err, rc, seg);
        return;
    }
    int col = q.col_index[slot], type = q.col_type[slot];
    size_t pi = (size_t)col * d.n_segments + seg;
    uint64_t *out = tp.vals[slot] + base; uint8_t *okb = tp.okb[slot] + base;
    uint32_t len = d.page_len[pi];
    if (len == 0) { for (uint32_t i = 0; i < rows; i++) { out[i * S] = 0; okb[i * S] = 0; } return; }
    PageHdr h;
    int rc = parse_field_header(d.data + d.page_off[pi], len, type, rows, h);
    if (rc == D_OK) {
        ExpandEmit em{out, okb, &h, 0, S};
        rc = decode_block(type, h, em);
        for (uint32_t i = em.row; i < rows; i++) { out[i * S] = 0; okb[i * S] = 0; }
    }
    if (rc != D_OK) report_err(err, rc, seg);
}

/* step 2: row mask = inside [tmin,tmax] AND WHERE RPN (one thread per row; SURVEY App.B.12 semantics) */
__device__ __forceinline__ bool term_pass(const FilterP &f, uint64_t raw) {
    if (f.type == OG_TYPE_FLOAT || (f.type == OG_TYPE_INT && f.const_is_float)) {
        double v = f.type == OG_TYPE_FLOAT ? u2d(raw) : (double)(int64_t)raw;
        double c = f.const_is_float ? f.fval : (double)f.ival;
        switch (f.op) {
        case OG_OP_LT: return !(v >= c);
        case OG_OP_LTE: return !(v > c);
        case OG_OP_GT: return !(v <= c);
        case OG_OP_GTE: return !(v < c);
        case OG_OP_EQ: return !(v != c);
        default: return !(v == c);
        }
    }
    int64_t v = f.type == OG_TYPE_BOOL ? (int64_t)(raw != 0) : (int64_t)raw, c = f.ival;
    switch (f.op) {
    case OG_OP_LT: return !(v >= c);
    case OG_OP_LTE: return !(v > c);
    case OG_OP_GT: return !(v <= c);
    case OG_OP_GTE: return !(v < c);
    case OG_OP_EQ: return !(v != c);
    default: return !(v == c);
    }
}

__global__ void k_filter_tile(DirP d, QueryP q, TileP tp) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)tp.S * tp.R;
    if (idx >= total) return;
    uint32_t row = (uint32_t)(idx / tp.S), sl = (uint32_t)(idx % tp.S);
    if (sl >= tp.tile_end - tp.tile_begin) return;
    if (row >= d.seg_rows[tp.tile_begin + sl]) { tp.keep[idx] = 0; return; }
    int64_t t = tp.times[idx];
    bool keep = t >= q.tmin && t <= q.tmax;
    if (keep && q.n_filter) {
        uint32_t stack = 0; int sp = 0;
        for (uint32_t i = 0; i < q.n_filter; i++) {
            const FilterP &f = q.filter[i];
            if (f.kind == OG_F_TERM) {
                bool pass = tp.okb[f.col_slot][idx] && term_pass(f, tp.vals[f.col_slot][idx]);
                stack |= (uint32_t)pass << sp; sp++;
            } else {
                uint32_t b = (stack >> (sp - 1)) & 1, a = (stack >> (sp - 2)) & 1;
                uint32_t r = f.kind == OG_F_AND ? (a & b) : (a | b);
                sp -= 2; stack &= ~(3u << sp); stack |= r << sp; sp++;
            }
        }
        keep = stack & 1;
    }
    tp.keep[idx] = keep;
}

/* where does a segment's partial for bucket b (w-th window of nwin) go */
__device__ __forceinline__ void emit_window(const QueryP &q, const ChunkP &ch, uint32_t seg, uint32_t series, uint32_t b,
                                            bool is_head, bool is_tail, int call, const Part &p) {
    if (is_head) store_part(ch.edges[call], 2 * (size_t)(seg - ch.seg_begin), p);
    else if (is_tail) store_part(ch.edges[call], 2 * (size_t)(seg - ch.seg_begin) + 1, p);
    else if (p.ok) store_part(ch.cells[call], cell_idx(ch, series, b), p);
}

/* step 3: one thread per segment walks its rows in time order (threads of a warp own consecutive segments and move row by row
 * together, so every load is coalesced in the r*S + sl layout); the rows of a window are accumulated left to right, which keeps
 * float sums in the reference's order (series_agg_func.gen.go:48-60).  Windows are those of the rows inside [tmin, tmax]; rows
 * removed by the WHERE mask do not contribute, a window whose rows were all removed yields an invalid partial. */
__global__ void k_window_reduce(DirP d, QueryP q, TileP tp, ChunkP ch) {
    uint32_t sl = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t seg = tp.tile_begin + sl;
    if (seg >= tp.tile_end) return;
    const uint32_t rows = d.seg_rows[seg], series = d.seg_series[seg];
    const size_t S = tp.S, e = 2 * (size_t)(seg - ch.seg_begin);
    Part parts[OG_MAX_CALLS];
    uint32_t cur_b = OG_NO_BUCKET, head_b = OG_NO_BUCKET; bool head_done = false;
    int64_t we = 0;
    auto flush = [&](bool final) {
        if (cur_b == OG_NO_BUCKET) return;
        for (uint32_t c = 0; c < q.n_calls; c++) {
            if (!head_done) store_part(ch.edges[c], e, parts[c]);
            else if (final) store_part(ch.edges[c], e + 1, parts[c]);
            else if (parts[c].ok) store_part(ch.cells[c], cell_idx(ch, series, cur_b), parts[c]);
        }
        if (!head_done) { head_done = true; head_b = cur_b; }
    };
    for (uint32_t r = 0; r < rows; r++) {
        const size_t ix = (size_t)r * S + sl;
        const int64_t t = tp.times[ix];
        if (t < q.tmin) continue;
        if (t > q.tmax) break;
        if (cur_b == OG_NO_BUCKET || t >= we) {
            flush(false);
            cur_b = bucket_of(t, q.start, q.interval);
            we = q.start + (int64_t)(cur_b + 1) * q.interval;
#pragma unroll
            for (uint32_t c = 0; c < OG_MAX_CALLS; c++) parts[c] = part_empty();
        }
        if (!tp.keep[ix]) continue;
#pragma unroll
        for (uint32_t c = 0; c < OG_MAX_CALLS; c++) {
            if (c >= q.n_calls) break;
            const CallP &cp = q.calls[c];
            if (tp.okb[cp.col_slot][ix]) acc_row(cp.func, cp.type, parts[c], tp.vals[cp.col_slot][ix], t);
        }
    }
    flush(true);
    ch.edge_bucket[e] = head_b;
    ch.edge_bucket[e + 1] = (head_b == OG_NO_BUCKET || cur_b == head_b) ? OG_NO_BUCKET : cur_b;
}

/* ------------------------------------------------------------------------------------------------------------
 * ordered stitch of windows that span segment boundaries (thread per segment)
 * ------------------------------------------------------------------------------------------------------------ */
__global__ void k_fix_edges(DirP d, QueryP q, ChunkP ch) {
    uint32_t seg = ch.seg_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= ch.seg_end) return;
    uint32_t series = d.seg_series[seg];
    uint32_t s_first = d.series_seg_begin[series], s_end = d.series_seg_begin[series + 1];
    const uint32_t *eb = ch.edge_bucket;
    size_t e = 2 * (size_t)(seg - ch.seg_begin);
    uint32_t hb = eb[e], tb = eb[e + 1];
    if (hb == OG_NO_BUCKET) return; /* no in-range rows */
    /* is head(seg) the leader of its run?  previous edge = tail(seg-1) if present else head(seg-1) */
    bool head_leader = true;
    if (seg > s_first) {
        size_t pe = e - 2;
        uint32_t pb = eb[pe + 1] != OG_NO_BUCKET ? eb[pe + 1] : eb[pe];
        if (pb != OG_NO_BUCKET && pb == hb) head_leader = false;
    }
    for (int which = 0; which < 2; which++) {
        uint32_t b = which == 0 ? hb : tb;
        if (which == 0 && !head_leader) continue;
        if (which == 1 && tb == OG_NO_BUCKET) continue;
        /* the run continues into later segments only from the last edge of this segment */
        bool can_extend = which == 1 || tb == OG_NO_BUCKET;
        for (uint32_t c = 0; c < q.n_calls; c++) {
            const CallP &cp = q.calls[c];
            Part acc = load_part(ch.edges[c], e + which);
            if (can_extend) {
                for (uint32_t nx = seg + 1; nx < s_end; nx++) {
                    size_t ne = 2 * (size_t)(nx - ch.seg_begin);
                    if (eb[ne] != b) break; /* includes OG_NO_BUCKET */
                    acc = series_merge(cp.func, cp.type, acc, load_part(ch.edges[c], ne));
                    if (eb[ne + 1] != OG_NO_BUCKET) break; /* that segment has a distinct tail window: run ends at its head */
                }
            }
            if (acc.ok) store_part(ch.cells[c], cell_idx(ch, series, b), acc);
        }
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * tagset merge: thread per (group, bucket); series of the group in shard order (CSR), restricted to the chunk
 * ------------------------------------------------------------------------------------------------------------ */
struct GroupP {
    const uint32_t *grp_begin;  /* [n_groups+1] */
    const uint32_t *grp_series; /* series ids sorted by (group, series) */
    uint32_t n_groups;
    Tri dense[OG_MAX_CALLS];    /* [g * n_buckets + b] accumulators (persist across chunks) */
};

__global__ void k_merge_groups(QueryP q, ChunkP ch, GroupP gp) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)gp.n_groups * q.n_buckets;
    if (idx >= total) return;
    uint32_t g = (uint32_t)(idx / q.n_buckets), b = (uint32_t)(idx % q.n_buckets);
    uint32_t lo = gp.grp_begin[g], hi = gp.grp_begin[g + 1];
    { /* first member >= series_begin */
        uint32_t a = lo, z = hi;
        while (a < z) { uint32_t m = (a + z) >> 1; if (gp.grp_series[m] < ch.series_begin) a = m + 1; else z = m; }
        lo = a;
    }
    if (lo >= hi || gp.grp_series[lo] >= ch.series_end) return;
    { /* one thread per (group, bucket, call): blockIdx.y = call */
        const uint32_t c = blockIdx.y;
        const CallP &cp = q.calls[c];
        Part acc = load_part(gp.dense[c], idx);
        const Tri cells = ch.cells[c];
        /* The fold is strictly sequential in series order (that is the reference's order, reccord_functions.go:730-733),
         * but the loads do not depend on it: fetch a batch of U partials first, so each thread keeps U independent
         * loads in flight (16,667 bucket threads alone cannot hide DRAM latency). */
        constexpr int U = 16;
        for (uint32_t i = lo; i < hi; i += U) {
            uint32_t okv[U]; uint64_t vv[U]; int64_t tt[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                uint32_t ii = i + u;
                uint32_t s = ii < hi ? gp.grp_series[ii] : 0xffffffffu;
                bool in = s < ch.series_end;
                size_t ci = in ? cell_idx(ch, s, b) : 0;
                okv[u] = in ? cells.ok[ci] : 0;
                vv[u] = in ? cells.val[ci] : 0;
                tt[u] = (in && cells.tim) ? cells.tim[ci] : 0;
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (!okv[u]) continue;
                Part p; p.ok = 1; p.v = vv[u]; p.t = tt[u];
                group_update(cp.func, cp.out_type == OG_TYPE_INT && cp.func == OG_AGG_COUNT ? OG_TYPE_INT : cp.type, q.multi != 0, acc, p);
            }
        }
        store_part(gp.dense[c], idx, acc);
    }
}

/* k_merge_groups for the one-tagset case (OG_GROUP_ALL) when no call carries a time: the cell matrix is bucket-major with
 * the series contiguous, so a block stages a [128 buckets][32 series] tile through shared memory with fully coalesced 256-byte
 * row loads, then thread t folds row t left to right — the same strictly sequential series order as k_merge_groups. */
__global__ void __launch_bounds__(128) k_merge_all(QueryP q, ChunkP ch, GroupP gp) {
    __shared__ uint64_t sv[128 * 33];
    __shared__ uint8_t sk[128 * 36];
    const uint32_t c = blockIdx.y, b0 = blockIdx.x * 128, t = threadIdx.x, lane = t & 31, w = t >> 5;
    const uint32_t nS = ch.series_end - ch.series_begin, nb = q.n_buckets;
    const CallP &cp = q.calls[c];
    const int ftype = cp.out_type == OG_TYPE_INT && cp.func == OG_AGG_COUNT ? OG_TYPE_INT : cp.type;
    const Tri cells = ch.cells[c];
    Part acc = part_empty();
    if (b0 + t < nb) acc = load_part(gp.dense[c], b0 + t);
    for (uint32_t s0 = 0; s0 < nS; s0 += 32) {
        const uint32_t sr = s0 + lane;
#pragma unroll 16
        for (uint32_t i = 0; i < 32; i++) {
            const uint32_t r = w + 4 * i, b = b0 + r;
            const bool in = b < nb && sr < nS;
            const size_t ci = in ? (size_t)b * ch.cell_sb + sr : 0;
            const uint8_t k = in ? cells.ok[ci] : (uint8_t)0;
            sv[r * 33 + lane] = in ? cells.val[ci] : 0;
            sk[r * 36 + lane] = k;
        }
        __syncthreads();
        if (b0 + t < nb) {
#pragma unroll 8
            for (uint32_t k = 0; k < 32; k++) {
                if (!sk[t * 36 + k]) continue;
                Part p; p.ok = 1; p.v = sv[t * 33 + k]; p.t = 0;
                group_update(cp.func, ftype, q.multi != 0, acc, p);
            }
        }
        __syncthreads();
    }
    if (b0 + t < nb) store_part(gp.dense[c], b0 + t, acc);
}

/* OG_GROUP_PER_SERIES: every tagset is one series, so the merge is a transposition of the bucket-major cell matrix into the
 * series-major dense record — done through a 32x32 shared-memory tile so both sides are coalesced.  Each element still goes
 * through group_update on an empty accumulator (same value/time rules as the general merge); k_init_dense is not needed. */
__global__ void __launch_bounds__(256) k_merge_per_series(QueryP q, ChunkP ch, GroupP gp) {
    __shared__ uint64_t sv[32][33];
    __shared__ int64_t st[32][33];
    __shared__ uint8_t sk[32][36];
    const uint32_t c = blockIdx.z, b0 = blockIdx.x * 32, s0 = blockIdx.y * 32, tx = threadIdx.x, ty = threadIdx.y;
    const uint32_t nS = ch.series_end - ch.series_begin, nb = q.n_buckets;
    const CallP &cp = q.calls[c];
    const int ftype = cp.out_type == OG_TYPE_INT && cp.func == OG_AGG_COUNT ? OG_TYPE_INT : cp.type;
    const Tri cells = ch.cells[c];
    for (uint32_t j = ty; j < 32; j += 8) {
        const uint32_t b = b0 + j, sr = s0 + tx;
        const bool in = b < nb && sr < nS;
        const size_t ci = in ? (size_t)b * ch.cell_sb + sr : 0;
        sk[j][tx] = in ? cells.ok[ci] : (uint8_t)0;
        sv[j][tx] = in ? cells.val[ci] : 0;
        st[j][tx] = (in && cells.tim) ? cells.tim[ci] : 0;
    }
    __syncthreads();
    for (uint32_t j = ty; j < 32; j += 8) {
        const uint32_t sr = s0 + j, b = b0 + tx;
        if (sr >= nS || b >= nb) continue;
        Part p; p.ok = sk[tx][j]; p.v = sv[tx][j]; p.t = st[tx][j];
        Part a; a.v = 0; a.ok = 0; a.t = q.multi ? 0 : q.start + (int64_t)b * q.interval;
        group_update(cp.func, ftype, q.multi != 0, a, p);
        store_part(gp.dense[c], (size_t)(ch.series_begin + sr) * nb + b, a);
    }
}

/* dense initialisation: values 0, valid 0, times = window start (single-call selectors) or 0 (RecMeta.Times) */
__global__ void k_init_dense(QueryP q, GroupP gp) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)gp.n_groups * q.n_buckets;
    if (idx >= total) return;
    uint32_t b = (uint32_t)(idx % q.n_buckets);
    for (uint32_t c = 0; c < q.n_calls; c++) {
        gp.dense[c].val[idx] = 0; gp.dense[c].ok[idx] = 0;
        if (gp.dense[c].tim) gp.dense[c].tim[idx] = q.multi ? 0 : q.start + (int64_t)b * q.interval;
    }
}

/* merge another shard's dense partial (same geometry) into ours: the cross-shard step for selector aggregates */
__global__ void k_merge_dense(QueryP q, GroupP mine, GroupP other) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)mine.n_groups * q.n_buckets;
    if (idx >= total) return;
    for (uint32_t c = 0; c < q.n_calls; c++) {
        const CallP &cp = q.calls[c];
        Part a = load_part(mine.dense[c], idx), p = load_part(other.dense[c], idx);
        group_update(cp.func, cp.func == OG_AGG_COUNT ? OG_TYPE_INT : cp.type, q.multi != 0, a, p);
        store_part(mine.dense[c], idx, a);
    }
}

} // namespace ogpu
