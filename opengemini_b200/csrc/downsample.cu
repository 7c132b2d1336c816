/*
 * downsample.cu — og_downsample: the read-aggregate-write pass of a downsample / level compaction behind ONE C-ABI call
 * (configs[4]).  Replaces, for one field column of one shard: engine/record_plan.go:494-830 (FileSequenceAggregator pulls
 * records, newProcessor reduces them per series and window) feeding engine/immutable/stream_downsample.go:454-600 (the
 * downsampled columns go through the ordinary column builders, column_builder.go:151-349, chunkdata_builder.go:65-97).
 *
 *   1. og_query_run with OG_GROUP_PER_SERIES and the six calls min, max, sum, count, first, last  -> dense [series][window]
 *   2. k_ds_count / k_ds_scatter: windows without rows are dropped (TransIntervalRec2Rec, lib/record/record.go:1298-1365), the
 *      kept windows of a series are packed front to back (stable) and cut into 1000-row segments (lib/util/util.go:72); the row
 *      time is the window start
 *   3. og_encode_pages per output column and for the time column
 *   4. the directory of the new shard (ChunkMeta contents: segment time ranges, page offsets / sizes) is assembled on the host
 *
 * Everything heavy stays on the device; the host sees one u32 per series, and two i64 plus seven (offset, size) pairs per
 * output segment.  The page bytes stay in HBM: og_downsampled_desc describes them with OG_SHARD_DEVICE_DATA, so the new shard can
 * be opened and queried in place, or copied out with og_downsampled_export to be written to a file.
 */
#include <cub/block/block_reduce.cuh>
#include <cub/block/block_scan.cuh>

#include <cstring>

#include <string>
#include <vector>

#include "internal.h"

using namespace ogpu;

namespace {

constexpr uint32_t DS_ROWS = 1000; /* rows per segment of the output (lib/util/util.go:72) */
constexpr int DS_COLS = 6;         /* min, max, sum, count, first, last */
constexpr int DS_THREADS = 256;

/* rows_out[s] = windows of series s that hold rows */
__global__ void k_ds_count(const uint8_t *keep, uint32_t nb, uint32_t *rows_out) {
    typedef cub::BlockReduce<uint32_t, DS_THREADS> Reduce;
    __shared__ typename Reduce::TempStorage tmp;
    const uint8_t *k = keep + (size_t)blockIdx.x * nb;
    uint32_t n = 0;
    for (uint32_t b = threadIdx.x; b < nb; b += DS_THREADS) n += k[b] != 0;
    n = Reduce(tmp).Sum(n);
    if (threadIdx.x == 0) rows_out[blockIdx.x] = n;
}

struct DsSrc { const uint64_t *val[DS_COLS]; };
struct DsDst { uint64_t *val[DS_COLS]; int64_t *time; };

/* block per series: kept windows, in time order, to cells [cell_base[s] + rank] of every output column */
__global__ void k_ds_scatter(DsSrc src, const uint8_t *keep, uint32_t nb, int64_t start, int64_t interval, const uint64_t *cell_base, DsDst dst) {
    typedef cub::BlockScan<uint32_t, DS_THREADS> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ uint32_t carry;
    const size_t row0 = (size_t)blockIdx.x * nb;
    const uint64_t base = cell_base[blockIdx.x];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t b0 = 0; b0 < nb; b0 += DS_THREADS) {
        const uint32_t b = b0 + threadIdx.x;
        const uint32_t k = (b < nb && keep[row0 + b]) ? 1u : 0u;
        uint32_t rank, total;
        Scan(tmp).ExclusiveSum(k, rank, total);
        const uint32_t before = carry;
        if (k) {
            const uint64_t at = base + before + rank;
#pragma unroll
            for (int c = 0; c < DS_COLS; c++) dst.val[c][at] = src.val[c][row0 + b];
            dst.time[at] = start + (int64_t)b * interval;
        }
        __syncthreads(); /* everyone has read carry and is done with tmp */
        if (threadIdx.x == 0) carry = before + total;
        __syncthreads();
    }
}

/* first / last row time of every output segment */
__global__ void k_ds_seg_times(const int64_t *time, const uint32_t *seg_rows, uint32_t n_seg, int64_t *tmin, int64_t *tmax) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_seg) return;
    tmin[g] = time[(size_t)g * DS_ROWS];
    tmax[g] = time[(size_t)g * DS_ROWS + seg_rows[g] - 1];
}

template <class T> int dsalloc(T **p, size_t n) {
    *p = nullptr;
    cudaError_t e = dev_malloc((void **)p, (n ? n : 1) * sizeof(T));
    if (e != cudaSuccess) { set_error("device allocation of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(e)); return e == cudaErrorMemoryAllocation ? OG_E_NOMEM : OG_E_CUDA; }
    return OG_OK;
}

struct Frees { /* device buffers released when the call returns */
    std::vector<void *> p;
    ~Frees() { for (void *q : p) dev_free(q); }
    template <class T> int get(T **out, size_t n) { int rc = dsalloc(out, n); if (rc == OG_OK) p.push_back(*out); return rc; }
};

} // namespace

struct og_downsampled {
    uint8_t *d_data = nullptr; uint64_t data_len = 0;
    uint64_t rows = 0;
    std::vector<uint64_t> sids; std::vector<uint32_t> ssb;
    std::vector<int64_t> tmin, tmax;
    std::vector<uint64_t> off[DS_COLS + 1]; std::vector<uint32_t> len[DS_COLS + 1]; /* time last */
    std::vector<std::string> names; int32_t types[DS_COLS];
    std::vector<og_column_desc> cols;
};

#define DS_CU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rc = cuda_fail(e__, #call, __FILE__, __LINE__); goto done; } } while (0)
#define DS_RC(call) do { rc = (call); if (rc != OG_OK) goto done; } while (0)

extern "C" {

OG_API void og_downsampled_free(og_downsampled *d) {
    if (!d) return;
    dev_free(d->d_data);
    delete d;
}

OG_API int og_downsample(og_shard *s, uint32_t column, int64_t interval, int64_t tmin, int64_t tmax, og_downsampled **out) {
    if (!s || !out || interval <= 0) { set_error("bad argument (interval must be > 0)"); return OG_E_INVAL; }
    *out = nullptr;
    static const int funcs[DS_COLS] = {OG_AGG_MIN, OG_AGG_MAX, OG_AGG_SUM, OG_AGG_COUNT, OG_AGG_FIRST, OG_AGG_LAST};
    static const char *fnames[DS_COLS] = {"min", "max", "sum", "count", "first", "last"};
    int rc = OG_OK;
    og_query *q = nullptr;
    og_downsampled *r = nullptr;
    Frees tmp;
    og_shard_layout lay;
    if ((rc = og_shard_layout_get(s, &lay))) return rc;
    if (column >= lay.n_columns) { set_error("column %u out of range", column); return OG_E_INVAL; }
    std::vector<int32_t> col_types(lay.n_columns);
    std::vector<uint64_t> sids(lay.n_series);
    if ((rc = og_shard_export(s, nullptr, sids.data(), nullptr, nullptr, nullptr, nullptr, nullptr, col_types.data()))) return rc;
    const int32_t ctype = col_types[column];
    if (ctype != OG_TYPE_FLOAT && ctype != OG_TYPE_INT) { set_error("downsample of a column of type %d", ctype); return OG_E_UNSUPPORTED; }

    og_call calls[DS_COLS];
    for (int c = 0; c < DS_COLS; c++) { calls[c].func = funcs[c]; calls[c].column = (int32_t)column; }
    og_query_desc qd{};
    qd.interval = interval; qd.tmin = tmin; qd.tmax = tmax; qd.ascending = 1; qd.n_calls = DS_COLS; qd.calls = calls;
    qd.group_mode = OG_GROUP_PER_SERIES;
    og_dense_view dv;
    uint32_t ns = 0, nb = 0, n_seg = 0;
    std::vector<uint32_t> rows_s, seg_rows;
    std::vector<uint64_t> cell_base;
    uint32_t *d_rows_s = nullptr, *d_seg_rows = nullptr; uint64_t *d_cell_base = nullptr;
    int64_t *d_time = nullptr, *d_tmin = nullptr, *d_tmax = nullptr;
    uint8_t *d_pages[DS_COLS + 1]; uint64_t page_bytes[DS_COLS + 1];
    uint64_t *d_off = nullptr; uint32_t *d_len = nullptr;
    DsSrc src; DsDst dst;
    uint64_t total = 0, cells = 0, pos = 0;

    DS_RC(og_query_create(s, &qd, &q));
    DS_RC(og_query_run(q));
    DS_RC(og_query_dense(q, &dv));
    ns = dv.n_groups; nb = dv.n_buckets;
    r = new og_downsampled;
    r->sids = sids;
    r->ssb.assign((size_t)ns + 1, 0);
    if (ns == 0 || nb == 0) goto directory;

    /* rows per series -> segments per series -> cell offsets */
    DS_RC(tmp.get(&d_rows_s, ns));
    k_ds_count<<<ns, DS_THREADS>>>(dv.cols[3].valid, nb, d_rows_s);
    rows_s.resize(ns);
    DS_CU(cudaMemcpy(rows_s.data(), d_rows_s, (size_t)ns * 4, cudaMemcpyDeviceToHost));
    cell_base.resize(ns);
    for (uint32_t i = 0; i < ns; i++) {
        const uint32_t segs = (rows_s[i] + DS_ROWS - 1) / DS_ROWS;
        cell_base[i] = (uint64_t)n_seg * DS_ROWS;
        for (uint32_t g = 0; g < segs; g++) seg_rows.push_back(g + 1 < segs ? DS_ROWS : rows_s[i] - g * DS_ROWS);
        if ((uint64_t)n_seg + segs > 0xfffffff0ull) { set_error("too many output segments"); rc = OG_E_UNSUPPORTED; goto done; }
        n_seg += segs; r->ssb[i + 1] = n_seg; r->rows += rows_s[i];
    }
    if (n_seg == 0) goto directory;
    cells = (uint64_t)n_seg * DS_ROWS;
    DS_RC(tmp.get(&d_cell_base, ns));
    DS_RC(tmp.get(&d_seg_rows, n_seg));
    DS_CU(cudaMemcpy(d_cell_base, cell_base.data(), (size_t)ns * 8, cudaMemcpyHostToDevice));
    DS_CU(cudaMemcpy(d_seg_rows, seg_rows.data(), (size_t)n_seg * 4, cudaMemcpyHostToDevice));
    for (int c = 0; c < DS_COLS; c++) {
        DS_RC(tmp.get(&dst.val[c], cells));
        DS_CU(cudaMemset(dst.val[c], 0, cells * 8)); /* cells past the last row of a series' last segment are never read, but keep them defined */
        src.val[c] = (const uint64_t *)dv.cols[c].values;
    }
    DS_RC(tmp.get(&d_time, cells));
    DS_CU(cudaMemset(d_time, 0, cells * 8));
    dst.time = d_time;
    k_ds_scatter<<<ns, DS_THREADS>>>(src, dv.cols[3].valid, nb, dv.start, dv.interval, d_cell_base, dst);
    DS_CU(cudaGetLastError());

    /* segment time ranges */
    DS_RC(tmp.get(&d_tmin, n_seg)); DS_RC(tmp.get(&d_tmax, n_seg));
    k_ds_seg_times<<<(n_seg + 255) / 256, 256>>>(d_time, d_seg_rows, n_seg, d_tmin, d_tmax);
    r->tmin.resize(n_seg); r->tmax.resize(n_seg);
    DS_CU(cudaMemcpy(r->tmin.data(), d_tmin, (size_t)n_seg * 8, cudaMemcpyDeviceToHost));
    DS_CU(cudaMemcpy(r->tmax.data(), d_tmax, (size_t)n_seg * 8, cudaMemcpyDeviceToHost));

    /* encode: six value columns, then time */
    DS_RC(tmp.get(&d_off, n_seg)); DS_RC(tmp.get(&d_len, n_seg));
    for (int c = 0; c <= DS_COLS; c++) {
        const bool is_time = c == DS_COLS;
        const int32_t typ = is_time ? OG_TYPE_INT : (funcs[c] == OG_AGG_COUNT ? OG_TYPE_INT : ctype);
        const uint64_t cap = (uint64_t)n_seg * 8800; /* a 1000-row page never exceeds 8 B per row + headers */
        DS_RC(tmp.get(&d_pages[c], cap));
        DS_RC(og_encode_pages(typ, is_time ? 1 : 0, is_time ? (const void *)d_time : (const void *)dst.val[c], nullptr, d_seg_rows, n_seg, DS_ROWS,
                              d_pages[c], cap, d_off, d_len, &page_bytes[c]));
        r->off[c].resize(n_seg); r->len[c].resize(n_seg);
        DS_CU(cudaMemcpy(r->off[c].data(), d_off, (size_t)n_seg * 8, cudaMemcpyDeviceToHost));
        DS_CU(cudaMemcpy(r->len[c].data(), d_len, (size_t)n_seg * 4, cudaMemcpyDeviceToHost));
        for (uint32_t g = 0; g < n_seg; g++) r->off[c][g] += total;
        total += page_bytes[c];
        if (!is_time) r->types[c] = typ;
    }
    /* one buffer: the columns back to back + the slack word-granular readers need behind the last page */
    {
        uint8_t *all = nullptr;
        cudaError_t e = dev_malloc((void **)&all, total + 1024);
        if (e != cudaSuccess) { rc = cuda_fail(e, "downsample output", __FILE__, __LINE__); goto done; }
        r->d_data = all; r->data_len = total;
        for (int c = 0; c <= DS_COLS; c++) { DS_CU(cudaMemcpy(all + pos, d_pages[c], page_bytes[c], cudaMemcpyDeviceToDevice)); pos += page_bytes[c]; }
        DS_CU(cudaMemset(all + total, 0, 1024));
    }

directory:
    if (!r->d_data) { /* nothing survived: an empty shard still has a valid (zero-length) data region */
        uint8_t *all = nullptr;
        cudaError_t e = dev_malloc((void **)&all, 1024);
        if (e != cudaSuccess) { rc = cuda_fail(e, "downsample output", __FILE__, __LINE__); goto done; }
        cudaMemset(all, 0, 1024);
        r->d_data = all; r->data_len = 0;
        for (int c = 0; c < DS_COLS; c++) r->types[c] = funcs[c] == OG_AGG_COUNT ? OG_TYPE_INT : ctype;
    }
    for (int c = 0; c < DS_COLS; c++) r->names.push_back(std::string(fnames[c]) + "_f" + std::to_string(column));
    for (int c = 0; c < DS_COLS; c++) {
        og_column_desc cd; cd.name = r->names[c].c_str(); cd.type = r->types[c]; cd.page_off = r->off[c].data(); cd.page_len = r->len[c].data();
        r->cols.push_back(cd);
    }
    DS_CU(cudaDeviceSynchronize());
    *out = r; r = nullptr;

done:
    if (q) og_query_destroy(q);
    if (r) og_downsampled_free(r);
    return rc;
}

OG_API int og_downsampled_desc(const og_downsampled *d, og_shard_desc *desc, uint64_t *rows) {
    if (!d || !desc) { set_error("null argument"); return OG_E_INVAL; }
    memset(desc, 0, sizeof *desc);
    desc->data = d->d_data; desc->data_len = d->data_len;
    desc->n_series = (uint32_t)d->sids.size(); desc->sids = d->sids.data(); desc->series_seg_begin = d->ssb.data();
    desc->n_segments = (uint32_t)d->tmin.size(); desc->seg_tmin = d->tmin.data(); desc->seg_tmax = d->tmax.data();
    desc->n_columns = DS_COLS; desc->columns = d->cols.data();
    desc->time_page_off = d->off[DS_COLS].data(); desc->time_page_len = d->len[DS_COLS].data();
    desc->flags = OG_SHARD_DEVICE_DATA;
    if (rows) *rows = d->rows;
    return OG_OK;
}

OG_API int og_downsampled_export(const og_downsampled *d, uint8_t *host_data) {
    if (!d || !host_data) { set_error("null argument"); return OG_E_INVAL; }
    if (d->data_len == 0) return OG_OK;
    cudaError_t e = cudaMemcpy(host_data, d->d_data, d->data_len, cudaMemcpyDeviceToHost);
    return e == cudaSuccess ? OG_OK : cuda_fail(e, "og_downsampled_export", __FILE__, __LINE__);
}

} // extern "C"
