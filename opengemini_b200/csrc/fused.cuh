/*
 * fused.cuh — K5: fused decode + time-bucket + reduce, one thread per segment, nothing materialised.
 *
 * Replaces, for queries that touch one field column and have no WHERE on fields:
 *   tsspFileReader.readSegmentRecord (tssp_file.go:369) + decodeColumnData (reader.go:674) + FilterByTime (reader.go:754)
 *   + aggregateCursor.getIntervalIndex/reduce (aggregate_cursor.go:306-356) + the per-window reduce kernels
 *   (series_agg_func.gen.go:24-274).
 * The thread walks the rows of its segment in order: time comes from a TimeIter over the time page (closed form for
 * const-delta pages), values from the column block decoder; per-window partials stay in registers and are flushed
 * when the row's time leaves the current window.  First/last window of the segment go to the edge arrays (they may
 * be shared with the neighbouring segments of the series), interior windows straight to cells[series][bucket].
 */
#pragma once
#include "agg_kernels.cuh"

namespace ogpu {

struct TimeIter {
    TimeDesc d; uint64_t cur; uint32_t idx; uint32_t w; unsigned k, n, bits; uint64_t word;
    __device__ __forceinline__ void init(const TimeDesc &t) { d = t; cur = (uint64_t)t.t0; idx = 0; w = 0; k = 0; n = 0; bits = 0; word = 0; }
    __device__ __forceinline__ int64_t next() { /* time of row idx, then advance */
        int64_t t;
        if (d.kind == 0 || d.kind == 3) { t = (int64_t)cur; cur += d.delta; }
        else if (d.kind == 2) { t = zigzag_dec(ld_be64(d.words + 8ull * idx)); }
        else {
            if (idx != 0) {
                while (k == n) { /* next simple8b word */
                    if (w >= d.n_words) { idx++; return (int64_t)cur; }
                    word = ld_be64(d.words + 8ull * w); w++;
                    s8b_sel((unsigned)(word >> 60), n, bits); k = 0;
                }
                uint64_t dv = bits == 0 ? 1ull : ((word >> (k * bits)) & ((1ull << bits) - 1));
                k++;
                cur += dv * d.delta;
            }
            t = (int64_t)cur;
        }
        idx++;
        return t;
    }
};

template <int NC>
struct FusedState {
    const QueryP &q; const ChunkP &ch; const PageHdr &h;
    TimeIter ti;
    Part parts[NC];
    uint32_t seg, series, row, cur_b, head_b;
    int64_t we; /* end of the current window */
    bool head_done;

    __device__ __forceinline__ FusedState(const QueryP &q_, const ChunkP &ch_, const PageHdr &h_) : q(q_), ch(ch_), h(h_) {}

    __device__ __forceinline__ void flush(bool final) {
        if (cur_b == OG_NO_BUCKET) return;
        size_t e = 2 * (size_t)(seg - ch.seg_begin);
        if (!head_done) {
#pragma unroll
            for (int c = 0; c < NC; c++) store_part(ch.edges[c], e, parts[c]);
            head_done = true; head_b = cur_b;
        } else if (final) {
#pragma unroll
            for (int c = 0; c < NC; c++) store_part(ch.edges[c], e + 1, parts[c]);
        } else {
#pragma unroll
            for (int c = 0; c < NC; c++) if (parts[c].ok) store_cell(ch, c, series, cur_b, parts[c]);
        }
    }
    __device__ __forceinline__ void row_step(bool valid, uint64_t bits) {
        int64_t t = ti.next();
        row++;
        if (t < q.tmin || t > q.tmax) return;
        if (cur_b == OG_NO_BUCKET || t >= we) {
            flush(false);
            cur_b = bucket_of(t, q.start, q.interval);
            if (cur_b >= q.n_buckets) { report_err(ch.err, D_CORRUPT, seg); cur_b = OG_NO_BUCKET; return; } /* cannot happen on a validated shard */
            we = q.start + (int64_t)(cur_b + 1) * q.interval;
#pragma unroll
            for (int c = 0; c < NC; c++) parts[c] = part_empty();
        }
        if (valid) {
#pragma unroll
            for (int c = 0; c < NC; c++) acc_row(q.calls[c].func, q.calls[c].type, parts[c], bits, t);
        }
    }
    /* decoder callback: value i belongs to the next valid row */
    __device__ __forceinline__ void operator()(uint32_t, uint64_t bits) {
        while (row < h.rows && !hdr_row_valid(h, row)) row_step(false, 0);
        if (row < h.rows) row_step(true, bits);
    }
};

template <int NC>
/* `list` (sorted segment ids, `n` of them) selects the segments of this chunk that k_fused_fast did not take, compacted so
 * that warps stay full; list == nullptr: every segment of the chunk. */
__global__ void __launch_bounds__(128) k_fused_segment(DirP d, QueryP q, ChunkP ch, const uint32_t *list, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t seg = list ? list[i] : ch.seg_begin + i;
    size_t e = 2 * (size_t)(seg - ch.seg_begin);
    uint32_t rows = d.seg_rows[seg];
    /* segment pruning by directory time range (location.go:276-280) */
    if (d.seg_tmax[seg] < q.tmin || d.seg_tmin[seg] > q.tmax || rows == 0) {
        ch.edge_bucket[e] = OG_NO_BUCKET; ch.edge_bucket[e + 1] = OG_NO_BUCKET; return;
    }
    size_t ti_idx = (size_t)d.n_columns * d.n_segments + seg;
    TimeDesc td;
    int rc = parse_time_page(d.data + d.page_off[ti_idx], d.page_len[ti_idx], td);
    if (rc != D_OK) { report_err(ch.err, rc, seg); ch.edge_bucket[e] = OG_NO_BUCKET; ch.edge_bucket[e + 1] = OG_NO_BUCKET; return; }
    int col = q.col_index[0], type = q.col_type[0];
    size_t pi = (size_t)col * d.n_segments + seg;
    uint32_t len = d.page_len[pi];
    PageHdr h;
    if (len == 0) { h.rows = rows; h.nil_count = rows; h.bitmap = nullptr; h.bm_off = 0; h.block = nullptr; h.block_len = 0; h.one_row = 0; }
    else {
        rc = parse_field_header(d.data + d.page_off[pi], len, type, rows, h);
        if (rc != D_OK) { report_err(ch.err, rc, seg); ch.edge_bucket[e] = OG_NO_BUCKET; ch.edge_bucket[e + 1] = OG_NO_BUCKET; return; }
    }
    FusedState<NC> st(q, ch, h);
    st.ti.init(td);
    st.seg = seg; st.series = d.seg_series[seg]; st.row = 0; st.cur_b = OG_NO_BUCKET; st.head_b = OG_NO_BUCKET; st.we = 0; st.head_done = false;
    if (h.nil_count < h.rows) {
        rc = decode_block(type, h, st);
        if (rc != D_OK) report_err(ch.err, rc, seg);
    }
    while (st.row < rows) st.row_step(false, 0); /* trailing null rows still define windows */
    uint32_t last_b = st.cur_b;
    bool single = !st.head_done;
    st.flush(true);
    ch.edge_bucket[e] = st.head_b;
    ch.edge_bucket[e + 1] = (single || st.head_b == OG_NO_BUCKET) ? OG_NO_BUCKET : last_b;
}

} // namespace ogpu
