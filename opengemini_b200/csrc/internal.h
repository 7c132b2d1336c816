/*
 * internal.h — host-side handle layouts and kernel parameter blocks shared by the .cu files of libogpu.so.
 */
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <mutex>
#include <vector>

#include "../../include/ogpu.h"

#define OG_MAX_CALLS 8
#define OG_MAX_FILTER 16
#define OG_MAX_COLS 8 /* distinct field columns one query may touch */

namespace ogpu {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);
int ensure_device(); /* bind the calling thread to the library's device (og_init(0) on first use) */

/* Device memory comes from the device's stream-ordered pool (cudaMallocAsync on the legacy stream) with its release threshold
 * raised at og_init: shards and queries that are opened and closed in a loop get their buffers back from the pool instead of
 * paying cudaMalloc/cudaFree (tens of ms per GB-sized buffer) every time.  Every buffer is released only after the stream that
 * used it has been synchronised (og_query_run / og_shard_open return synchronised), so reuse across streams is ordered. */
inline cudaError_t dev_malloc(void **p, size_t bytes) { return cudaMallocAsync(p, bytes ? bytes : 1, (cudaStream_t)0); }
inline void dev_free(void *p) { if (p) cudaFreeAsync(p, (cudaStream_t)0); }
/* free device memory as a budget sees it: what the driver reports plus what the pool holds but does not use */
cudaError_t dev_mem_info(size_t *free_b, size_t *total_b);
#define CU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return ::ogpu::cuda_fail(e__, #call, __FILE__, __LINE__); } while (0)

struct DevBuf { /* RAII-less helper: explicit free */
    void *p = nullptr; size_t bytes = 0;
};

} // namespace ogpu

struct og_shard {
    int device = 0;
    uint8_t *d_data = nullptr; uint64_t data_len = 0; bool owns_data = true;
    uint32_t n_series = 0, n_segments = 0, n_columns = 0;
    uint32_t max_seg_rows = 0;
    uint64_t irregular_time_pages = 0; /* time pages that are neither const-delta nor one-row (Simple8b / raw times) */
    uint64_t n_rows = 0, page_bytes = 0;
    uint64_t snappy_pages = 0, snappy_bytes_in = 0, snappy_bytes_out = 0; /* Snappy pages transcoded to raw at open */
    int64_t tmin = 0, tmax = 0;
    std::vector<uint64_t> sids;
    std::vector<int32_t> col_types;
    std::vector<std::string> col_names;
    std::vector<uint32_t> h_series_seg_begin; /* always mirrored on host (n_series+1) */
    /* device directory (SoA) */
    uint32_t *d_series_seg_begin = nullptr; /* [n_series+1] */
    uint32_t *d_seg_series = nullptr;       /* [n_segments] */
    uint32_t *d_seg_rows = nullptr;         /* [n_segments] rows per segment (from the time page) */
    int64_t *d_tmin = nullptr, *d_tmax = nullptr; /* [n_segments] */
    uint64_t *d_page_off = nullptr;         /* [(n_columns+1) * n_segments], time column last */
    uint32_t *d_page_len = nullptr;
    uint64_t *d_sids = nullptr;
    /* materialise scratch for og_decode_segment (host pinned + device) */
    void *h_seg_buf = nullptr; size_t h_seg_buf_bytes = 0;
    void *d_seg_buf = nullptr; size_t d_seg_buf_bytes = 0;
    std::vector<og_colval_view> seg_views;
    /* lane-interleaved, length-binned stream copy per column for the fused Gorilla kernel (fused_il.cuh), built on first use */
    struct IlCol { int state = 0; /* 0 not built, 1 ready, -1 no eligible segment, -2 not enough device memory (general kernel serves the column) */
                   uint32_t *words = nullptr; uint64_t *grp_off = nullptr; uint32_t *grp_rows = nullptr, *grp_col = nullptr; uint8_t *ok = nullptr;
                   uint32_t *lane_seg = nullptr, *lane_rows = nullptr, *lane_series = nullptr; int64_t *lane_t0 = nullptr; uint64_t *lane_dt = nullptr;
                   uint32_t *gen_list = nullptr; std::vector<uint32_t> gen_host; /* segments the fused kernel does not take (ascending), device + host */
                   uint32_t n_groups = 0, J = 0; /* J != 0: regular shard, lane groups share a segment index */
                   uint32_t n_super = 1, cols_per_super = 0; std::vector<uint32_t> super_grp_first; /* [n_super+1] first lane group of each block of OG_IL_SUPER series */
                   uint64_t n_words = 0; double build_ms = 0; };
    std::vector<IlCol> il; /* [n_columns] */
    std::mutex il_mu;      /* queries of one shard may be planned from different threads: the build is serialised */
};

namespace ogpu {

struct CallP { int32_t func, col_slot, type, out_type; };
struct FilterP { int32_t kind, col_slot, op, type, const_is_float; double fval; int64_t ival; };

/* per-call cell/edge/dense array triple */
struct Tri { uint64_t *val; int64_t *tim; uint8_t *ok; };

struct QueryP { /* passed by value to kernels */
    int64_t tmin, tmax, start, interval; /* interval = window length (end-start of Window()); >0 always on device */
    uint32_t n_buckets, n_calls, n_filter, n_cols;
    int32_t multi; /* callCount > 1 */
    CallP calls[OG_MAX_CALLS];
    FilterP filter[OG_MAX_FILTER];
    int32_t col_index[OG_MAX_COLS]; /* shard column of each slot */
    int32_t col_type[OG_MAX_COLS];
};

} // namespace ogpu

struct og_query {
    og_shard *sh = nullptr;
    og_query_desc desc{};
    std::vector<og_call> calls;
    std::vector<og_filter_item> filter;
    std::vector<uint32_t> series_group;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    ogpu::QueryP qp{};
    uint32_t n_groups = 0;
    bool ran = false; volatile int aborted = 0;
    /* dense result */
    ogpu::Tri dense[OG_MAX_CALLS]{};
    og_dense_col dense_cols[OG_MAX_CALLS]{};
    /* group membership CSR on device (series sorted by group, stable) */
    uint32_t *d_group_of_series = nullptr; /* [n_series] */
    /* scratch */
    std::vector<void *> scratch;
    og_stats stats{};
    /* host staging for og_query_next */
    std::vector<uint64_t> h_val[OG_MAX_CALLS];
    std::vector<uint8_t> h_ok[OG_MAX_CALLS];
    std::vector<int64_t> h_tim[OG_MAX_CALLS];
    bool host_ready = false;
    uint32_t next_group = 0, next_row = 0;
    /* record view backing store */
    std::vector<std::vector<uint8_t>> rv_val, rv_bitmap;
    std::vector<std::vector<int64_t>> rv_coltimes;
    std::vector<int64_t> rv_times;
    std::vector<og_colval_view> rv_cols;
    int path_used = 0; /* 0 generic tile path, 1 fused (general kernel), 2 fused Gorilla kernel, per-series cells, 3 fused Gorilla kernel, folded cells, 4 pull-iterator kernel k_fused_multi, 5 column-at-a-time kernel k_fused_cols */
    bool cells_dirty = true; /* per-series cell validity bytes need clearing before the next run */
    /* execution plan + scratch, built by the first og_query_run and reused by later runs */
    bool planned = false;
    uint32_t chunk_series = 0, tile_segs = 0;
    int *d_err = nullptr;
    void *plan = nullptr; /* ogpu::Plan (agg_kernels.cuh types) */
    std::vector<cudaEvent_t> main_ev; /* event pairs around the dominant decode+reduce kernels */
    void *merge_state = nullptr;      /* og_merge_state of the last og_query_allreduce (comm.cu) */
};
