/*
 * fused_fast.cuh — K5 specialised for the headline shape: float64 Gorilla pages (tag 3, no nulls) with const-delta
 * time pages.  One thread per segment; a warp = one "lane group" of 32 segments: segment j of 32 CONSECUTIVE SERIES on
 * regular shards (every series has the same number of segments), so that the lanes share window boundaries, flush together and
 * write one bucket of 32 series as one 256-byte run; otherwise 32 consecutive segments.
 *
 *   layout    Gorilla decode is serial per stream, so a warp reads 32 different pages.  To make those reads coalesced
 *             the shard keeps, next to the pages, a LANE-INTERLEAVED copy of every eligible stream (built once per
 *             shard and column by k_il_repack, on first use): the streams of a lane group are cut into 32-bit
 *             big-endian words and word w of lane l is stored at  il[grp_off + w*32 + l].  One 128-byte row therefore
 *             holds word w of all 32 lanes: a warp-wide 4-byte access is one fully used line, in HBM and in shared
 *             memory alike, whatever the lanes' individual positions are.
 *   staging   each lane copies ITS words with 4-byte cp.async (LDGSTS) into the warp's shared-memory window
 *             [OG_IL_NW rows + 2 mirror rows][32 lanes]; bank = lane, so neither the copies nor the loads ever conflict.
 *             cp.async groups are per thread: no mbarrier, no cross-lane signalling, no uniform-datapath waterfall.
 *             Refill runs on a fixed schedule (every K records): a record consumes <= 77 bits, so a 32-word window
 *             refilled every 4 records always holds two service periods of look-ahead (see the proof at the loop;
 *             measured best of 32/4, 64/8, 64/4, 128/16 because it leaves the most warps resident).
 *   decode    stateless bit addressing: the 64 bits at bit position p are three LDS.32 at immediate row offsets + two
 *             funnel shifts (words are pre-swapped to native order by the repack); the '10' (window reuse) record —
 *             >95% of records on noisy-mantissa data — is then one shift + mask + xor and p += 2+m.
 *   reduce    window boundaries are row countdowns derived from the const-delta time page (no time decode and no
 *             division in the loop); when no lane reaches a boundary within the next K records — 14 rounds of 15 on
 *             regular shards — the records run without the per-record boundary test.  Partials stay in registers and
 *             are flushed to the same edge/cell arrays the general kernel uses, so k_fix_edges / k_merge_* are shared
 *             and float sums keep the reference's left-to-right order.
 * k_fused_raw takes the raw pages of the same columns (Gorilla output above 90 % of raw): warp per segment, lane per window.
 *
 * Earlier staging designs (per-lane TMA bulk copies, warp-cooperative cp.async + mbarriers) read the pages where they
 * lie; both spent more instructions on staging than on decoding — see profiles/r01_fast_kernel_history.md.
 *
 * Replaces for eligible segments: tsm1.FloatArrayDecodeAll (batch_float.go:278-514) + Time.constDeltaDecoding
 * (timestamp.go:190) + FilterByTime (reader.go:754) + getIntervalIndex/reduce (aggregate_cursor.go:306-356) +
 * float{Sum,Min,Max,First,Last}Reduce / *CountReduce (series_agg_func.gen.go:24-274).
 * Segments that are not eligible (other codecs, nulls, irregular time pages) are left to k_fused_segment.
 */
#pragma once
#include "agg_kernels.cuh"

namespace ogpu {

/* ---- PTX wrappers ---- */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t opaque(uint32_t x) { asm volatile("mov.u32 %0, %0;" : "+r"(x)); return x; } /* keeps a value in a register instead of being rematerialised */
/* 4-byte asynchronous global->shared copy (LDGSTS.32); groups are per thread */
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory"); }
template <int DOFF, int SOFF> __device__ __forceinline__ void cp_async4o(uint32_t dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0+%2], [%1+%3], 4;" ::"r"(dst), "l"(src), "n"(DOFF), "n"(SOFF) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
template <int OFF> __device__ __forceinline__ uint32_t lds32o(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF)); return v; }

enum { FM_COUNT = 1, FM_SUM = 2, FM_MIN = 4, FM_MAX = 8, FM_FIRST = 16, FM_LAST = 32 };
enum { SEG_GENERAL = 0, SEG_FAST = 1, SEG_RAW = 2 };  /* per-segment classes computed by k_classify */

#define OG_FAST_THREADS 128
#ifndef OG_FAST_MINB
#define OG_FAST_MINB 1
#endif
#ifndef OG_IL_NW
#define OG_IL_NW 32u            /* window rows (words per lane) */
#endif
#define OG_IL_ROWS (OG_IL_NW + 2u) /* + 2 mirror rows that repeat rows 0,1 so that three consecutive rows never wrap */
#ifndef OG_IL_K
#define OG_IL_K 4u              /* records between two refills */
#endif
#ifndef OG_IL_BATCH
#define OG_IL_BATCH 4u          /* words per refill batch (4 or 8) */
#endif
#define OG_IL_PAD_WORDS 6u      /* words appended to every stream: the decoder may touch 77 + 64 bits past the last record */
#define OG_RAW_STAGE 8208u     /* k_fused_raw staging bytes per warp: 1024 rows + alignment slack */
#define OG_IL_HDR 7u            /* page = [31][rows u32][0x30][0x10] | stream: first value 8 B BE, records... */

/* lane-interleaved stream copy of one column (owned by the shard, built lazily) */
struct IlP {
    const uint32_t *words;     /* il[grp_off[g] + w*32 + lane] */
    const uint64_t *grp_off;   /* [n_groups32] in words */
    const uint32_t *grp_words; /* [n_groups32] words per lane in this group (0: no eligible lane) */
    const uint8_t *ok;         /* [n_segments] static class (SEG_*) by codec + header shape */
    const uint32_t *lane_seg;  /* [n_groups32 * 32] segment of every lane slot, OG_IL_NONE = empty.  Regular shards (every
                                  series has the same number of segments) put segment j of 32 CONSECUTIVE SERIES in one group:
                                  the lanes then share window boundaries (flushes coincide, no divergence) and write one
                                  bucket of 32 consecutive series = one 256-byte run of the cell matrix.  Otherwise a group
                                  is 32 consecutive segments. */
};
#define OG_IL_NONE 0xffffffffu

/* static eligibility + stream length in words (one thread per segment; header bytes only) */
__global__ void k_il_scan(DirP d, int col, int col_type, uint8_t *ok, uint32_t *seg_words) {
    uint32_t seg = blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= d.n_segments) return;
    uint8_t c = 0; uint32_t nw = 0;
    uint32_t rows = d.seg_rows[seg];
    if (rows >= 2 && col_type == OG_TYPE_FLOAT) {
        size_t pi = (size_t)col * d.n_segments + seg, ti = (size_t)d.n_columns * d.n_segments + seg;
        const uint8_t *p = d.data + d.page_off[pi], *t = d.data + d.page_off[ti];
        uint32_t len = d.page_len[pi], tlen = d.page_len[ti];
        /* value page: [31][u32 rows][0x30][0x10][8 B first]...; time page: [32][u32 rows][0x10][t0][uvarint dt][uvarint n-1] */
        if (len >= 16 && tlen >= 16 && __ldg(p) == 31 && (__ldg(p + 5) >> 4) == 3 && __ldg(t) == 32 && (__ldg(t + 5) >> 4) == 1) {
            TimeDesc td;
            if (parse_time_page(t, tlen, td) == D_OK && td.kind == 0 && td.delta > 0 && td.delta < (1ull << 40) && ld_be32(p + 1) == rows) {
                c = SEG_FAST; nw = ((len - OG_IL_HDR + 3) / 4 + OG_IL_PAD_WORDS + 7) & ~7u; /* refill copies 8-word batches */
            }
        } else if (len == 6 + 8 * (size_t)rows && tlen >= 16 && __ldg(p) == 31 && (__ldg(p + 5) >> 4) == 0 && __ldg(t) == 32 && (__ldg(t + 5) >> 4) == 1) {
            /* raw page (Gorilla output above 90% of raw, float.go:96-99): [31][u32 rows][0x00][rows x 8 B LE] -> k_fused_raw */
            TimeDesc td;
            if (parse_time_page(t, tlen, td) == D_OK && td.kind == 0 && td.delta > 0 && td.delta < (1ull << 40) && ld_be32(p + 1) == rows) c = SEG_RAW;
        }
    }
    ok[seg] = c; seg_words[seg] = nw;
}

/* words per lane of every lane group = max over its eligible lanes (one warp per group) */
__global__ void k_il_group_words(uint32_t n_groups, const uint32_t *lane_seg, const uint32_t *seg_words, uint32_t *grp_words) {
    uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g >= n_groups) return;
    uint32_t seg = lane_seg[(size_t)g * 32 + lane];
    uint32_t w = seg != OG_IL_NONE ? seg_words[seg] : 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
    if (lane == 0) grp_words[g] = w;
}

/* the repack: word w of lane l -> il[grp_off + w*32 + l], big-endian stream words stored in native order (one warp per group) */
__global__ void k_il_repack(DirP d, int col, const uint8_t *ok, const uint32_t *lane_seg, const uint64_t *grp_off, const uint32_t *grp_words,
                            uint32_t n_groups, uint32_t *il) {
    uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g >= n_groups) return;
    uint32_t nw = grp_words[g];
    if (nw == 0) return;
    uint32_t seg = lane_seg[(size_t)g * 32 + lane];
    bool live = seg != OG_IL_NONE && ok[seg] == SEG_FAST;
    const uint32_t *base = nullptr; uint32_t sh = 0, own_words = 0;
    if (live) {
        size_t pi = (size_t)col * d.n_segments + seg;
        const uint8_t *s = d.data + d.page_off[pi] + OG_IL_HDR;
        base = (const uint32_t *)((uintptr_t)s & ~(uintptr_t)3);
        sh = (uint32_t)((uintptr_t)s & 3);
        own_words = (d.page_len[pi] - OG_IL_HDR + 3) / 4 + OG_IL_PAD_WORDS; /* bytes past the page are the next page or the shard's tail padding */
    }
    uint32_t *out = il + grp_off[g] + lane;
    /* bytes s[4w..4w+3] big-endian: from aligned words a=base[w], b=base[w+1] (little-endian loads) */
    const uint32_t sel = sh == 0 ? 0x0123u : sh == 1 ? 0x1234u : sh == 2 ? 0x2345u : 0x3456u;
    uint32_t a = live ? __ldg(base) : 0;
    for (uint32_t w = 0; w < nw; w++) {
        uint32_t v = 0;
        if (w < own_words) {
            uint32_t b = __ldg(base + w + 1);
            v = __byte_perm(a, b, sel);
            a = b;
        }
        out[(size_t)w * 32] = v;
    }
}

/* per-query classes: the static class (IlP.ok) for segments that overlap the query's time range, else SEG_GENERAL */
__global__ void k_classify(DirP d, QueryP q, const uint8_t *ok, uint8_t *cls) {
    uint32_t seg = blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= d.n_segments) return;
    cls[seg] = (d.seg_tmax[seg] < q.tmin || d.seg_tmin[seg] > q.tmax) ? (uint8_t)SEG_GENERAL : ok[seg];
}

/* the segments of class `klass`, appended in any order (sorted on the host afterwards) */
__global__ void k_list_class(const uint8_t *cls, uint32_t n_segments, uint8_t klass, uint32_t *list, uint32_t *count) {
    uint32_t seg = blockIdx.x * blockDim.x + threadIdx.x;
    bool g = seg < n_segments && cls[seg] == klass;
    uint32_t m = __ballot_sync(0xffffffffu, g), lane = threadIdx.x & 31;
    if (m == 0) return;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (g) list[base + __popc(m & ((1u << lane) - 1))] = seg;
}

/* Raw float pages with const-delta time: every row is addressable, so one warp takes a segment and each lane reduces one
 * window (rows of a window sequentially: float sums keep the reference order).  Same outputs as k_fused_segment. */
template <int FM, bool TIMES>
__global__ void __launch_bounds__(128) k_fused_raw(DirP d, QueryP q, ChunkP ch, const uint32_t *list, uint32_t n) {
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n) return;
    const uint32_t seg = list[i];
    const size_t e = 2 * (size_t)(seg - ch.seg_begin);
    const uint32_t rows = d.seg_rows[seg], series = d.seg_series[seg];
    const size_t ti_idx = (size_t)d.n_columns * d.n_segments + seg;
    const uint8_t *tp = d.data + d.page_off[ti_idx];
    const int64_t t0 = (int64_t)ld_be64(tp + 6);
    uint64_t dtu = 1; ld_uvarint(tp + 14, d.page_len[ti_idx] - 14, &dtu);
    const int64_t dt = (int64_t)dtu;
    uint32_t r_lo = 0, r_hi = rows - 1;
    if (t0 < q.tmin) { uint64_t k = ((uint64_t)(q.tmin - t0) + dtu - 1) / dtu; r_lo = k > rows ? rows : (uint32_t)k; }
    { int64_t t_last = t0 + (int64_t)(rows - 1) * dt; if (t_last > q.tmax) { if (q.tmax < t0) r_lo = rows; else r_hi = (uint32_t)((uint64_t)(q.tmax - t0) / dtu); } }
    if (r_lo > r_hi || r_lo >= rows) { if (lane == 0) { ch.edge_bucket[e] = OG_NO_BUCKET; ch.edge_bucket[e + 1] = OG_NO_BUCKET; } return; }
    const uint8_t *vals = d.data + d.page_off[(size_t)q.col_index[0] * d.n_segments + seg] + 6;
    /* stage the page through shared memory with coalesced 16-byte loads (one DRAM round trip per page instead of one per
     * row); pages longer than the staging buffer are read in place */
    __shared__ __align__(16) uint8_t s_page[4][OG_RAW_STAGE];
    const uint8_t *g16 = (const uint8_t *)((uintptr_t)vals & ~(uintptr_t)15);
    const uint32_t mis = (uint32_t)(vals - g16), need = mis + 8 * rows;
    const bool staged = need <= OG_RAW_STAGE;
    if (staged) {
        uint4 *dst = (uint4 *)s_page[threadIdx.x >> 5];
        for (uint32_t o = lane; o * 16 < need; o += 32) dst[o] = __ldg((const uint4 *)g16 + o);
        __syncwarp();
    }
    const uint8_t *srow = s_page[threadIdx.x >> 5] + mis;
    const uint32_t b_first = bucket_of(t0 + (int64_t)r_lo * dt, q.start, q.interval), b_last = bucket_of(t0 + (int64_t)r_hi * dt, q.start, q.interval);
    const uint32_t nwin = b_last - b_first + 1;
    for (uint32_t w = lane; w < nwin; w += 32) {
        const int64_t W0 = q.start + (int64_t)(b_first + w) * q.interval, W1 = W0 + q.interval;
        /* rows with W0 <= t0 + r*dt < W1, clipped to [r_lo, r_hi] */
        uint32_t ra = r_lo, rb = r_hi + 1;
        if (W0 > t0) { uint64_t k = ((uint64_t)(W0 - t0) + dtu - 1) / dtu; if (k > ra) ra = k > rb ? rb : (uint32_t)k; }
        { uint64_t k = ((uint64_t)(W1 - t0) + dtu - 1) / dtu; if (k < rb) rb = (uint32_t)k; }
        if (ra >= rb) continue; /* a window without rows (cadence coarser than the interval) */
        /* same accumulators and seeding rules as k_fused_fast (first value seeds min/max/first; strict compares) */
        double sum = 0.0; uint64_t mn = 0, mx = 0, fi = 0, lastv = 0; uint32_t r_mn = ra, r_mx = ra;
        for (uint32_t r = ra; r < rb; r++) {
            uint64_t v;
            if (staged) { /* unaligned 8 bytes from three aligned words */
                const uint8_t *pr = srow + 8 * (size_t)r;
                const uint32_t *w32 = (const uint32_t *)((uintptr_t)pr & ~(uintptr_t)3);
                const uint32_t sh = ((uint32_t)(uintptr_t)pr & 3) * 8;
                v = ((uint64_t)__funnelshift_r(w32[1], w32[2], sh) << 32) | __funnelshift_r(w32[0], w32[1], sh);
            }
            else v = ld_le64(vals + 8 * (size_t)r);
            if (r == ra) { mn = mx = fi = v; }
            if (FM & FM_SUM) sum = sum + u2d(v);
            if (FM & FM_MIN) { if (u2d(mn) > u2d(v)) { mn = v; r_mn = r; } }
            if (FM & FM_MAX) { if (u2d(mx) < u2d(v)) { mx = v; r_mx = r; } }
            lastv = v;
        }
        const uint32_t cnt = rb - ra;
        const int kind = w == 0 ? 0 : w == nwin - 1 ? 1 : 2;
        const size_t idx = kind == 2 ? cell_idx(ch, series, b_first + w) : e + kind;
#pragma unroll 1
        for (uint32_t c = 0; c < q.n_calls; c++) {
            Part pp; pp.ok = 1; pp.v = 0; pp.t = 0;
            switch (q.calls[c].func) {
            case OG_AGG_COUNT: pp.v = cnt; break;
            case OG_AGG_SUM: pp.v = d2u(sum); break;
            case OG_AGG_MIN: pp.v = mn; if (TIMES) pp.t = t0 + (int64_t)r_mn * dt; break;
            case OG_AGG_MAX: pp.v = mx; if (TIMES) pp.t = t0 + (int64_t)r_mx * dt; break;
            case OG_AGG_FIRST: pp.v = fi; pp.t = t0 + (int64_t)ra * dt; break;
            default: pp.v = lastv; pp.t = t0 + (int64_t)(rb - 1) * dt; break;
            }
            const Tri &dst = kind == 2 ? ch.cells[c] : ch.edges[c];
            store_part(dst, idx, pp);
        }
    }
    if (lane == 0) { ch.edge_bucket[e] = b_first; ch.edge_bucket[e + 1] = nwin > 1 ? b_last : OG_NO_BUCKET; }
}

/* 64 bits of the stream at bit position p: rows (p>>5), +1, +2 of the lane's window column */
__device__ __forceinline__ uint64_t fetch64(uint32_t col, uint32_t p) {
    uint32_t a = col + ((p << 2) & ((OG_IL_NW - 1) << 7)); /* ((p >> 5) % NW) * 128 */
    uint32_t a0 = lds32o<0>(a), a1 = lds32o<128>(a), a2 = lds32o<256>(a);
    uint32_t hi = __funnelshift_l(a1, a0, p), lo = __funnelshift_l(a2, a1, p); /* shift amount taken mod 32 */
    return ((uint64_t)hi << 32) | lo;
}

template <int FM, bool TIMES>
__global__ void __launch_bounds__(OG_FAST_THREADS, OG_FAST_MINB) k_fused_fast(DirP d, QueryP q, ChunkP ch, const uint8_t *cls, IlP il, uint32_t grp_begin, uint32_t grp_end) {
    constexpr uint32_t NW = OG_IL_NW;
    constexpr uint32_t FULL = 0xffffffffu;
    __shared__ __align__(128) uint32_t s_win[(OG_FAST_THREADS / 32) * OG_IL_ROWS * 32];

    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const uint32_t grp = grp_begin + blockIdx.x * (OG_FAST_THREADS / 32) + wid;
    const uint32_t seg = grp < grp_end ? il.lane_seg[(size_t)grp * 32 + lane] : OG_IL_NONE;
    /* a lane group that straddles two chunks runs in both, with complementary lanes */
    bool active = seg != OG_IL_NONE && seg >= ch.seg_begin && seg < ch.seg_end && cls[seg] == SEG_FAST;
    if (!__any_sync(FULL, active)) return;
    const size_t e = 2 * (size_t)(seg - ch.seg_begin);

    uint32_t rows = 0, series = 0, r_lo = 0, r_hi = 0;
    int64_t t0 = 0, dt = 1; uint64_t dtu = 1;
    if (active) {
        rows = d.seg_rows[seg]; series = d.seg_series[seg];
        /* const-delta time page: t(r) = t0 + r*dt */
        const size_t ti_idx = (size_t)d.n_columns * d.n_segments + seg;
        const uint8_t *tp = d.data + d.page_off[ti_idx];
        t0 = (int64_t)ld_be64(tp + 6);
        ld_uvarint(tp + 14, d.page_len[ti_idx] - 14, &dtu);
        dt = (int64_t)dtu;
        /* rows inside [tmin, tmax] (FilterByTime) */
        r_lo = 0; r_hi = rows - 1;
        if (t0 < q.tmin) { uint64_t k = ((uint64_t)(q.tmin - t0) + dtu - 1) / dtu; r_lo = k > rows ? rows : (uint32_t)k; }
        { int64_t t_last = t0 + (int64_t)(rows - 1) * dt; if (t_last > q.tmax) { if (q.tmax < t0) r_lo = rows; else r_hi = (uint32_t)((uint64_t)(q.tmax - t0) / dtu); } }
        if (r_lo > r_hi || r_lo >= rows) { ch.edge_bucket[e] = OG_NO_BUCKET; ch.edge_bucket[e + 1] = OG_NO_BUCKET; active = false; }
    }

    /* ---- the lane's column of the warp window, and its interleaved source ---- */
    const uint32_t col = opaque(smem_u32(s_win) + wid * (OG_IL_ROWS * 128) + lane * 4);
    const uint32_t n_words = active ? il.grp_words[grp] : 0;
    const uint32_t *src = il.words + (active ? il.grp_off[grp] : 0) + lane; /* word `issued` is at src[0] */
    uint32_t issued = 0;
    auto refill = [&](uint32_t lim) { /* copy whole batches while they fit below lim (n_words is a multiple of 8) */
        while (issued + OG_IL_BATCH <= lim) {
            const uint32_t r = issued & (NW - 1), dst = col + r * 128;
            cp_async4o<0 * 128, 0 * 128>(dst, src); cp_async4o<1 * 128, 1 * 128>(dst, src);
            cp_async4o<2 * 128, 2 * 128>(dst, src); cp_async4o<3 * 128, 3 * 128>(dst, src);
#if OG_IL_BATCH == 8
            cp_async4o<4 * 128, 4 * 128>(dst, src); cp_async4o<5 * 128, 5 * 128>(dst, src);
            cp_async4o<6 * 128, 6 * 128>(dst, src); cp_async4o<7 * 128, 7 * 128>(dst, src);
#endif
            if (r == 0) { cp_async4o<NW * 128, 0>(dst, src); cp_async4o<(NW + 1) * 128, 128>(dst, src); } /* mirror rows */
            src += OG_IL_BATCH * 32; issued += OG_IL_BATCH;
        }
    };
    refill(min(NW, n_words));
    cp_async_commit();

    /* ---- window bookkeeping: bucket of row r_lo, first row of the next bucket (rb), Bresenham advance of rb ---- */
    uint32_t cur_b = 0, rb = 0xffffffffu, step_q = 0; uint64_t rem = 0, step_r = 0;
    if (active) {
        const int64_t t_lo = t0 + (int64_t)r_lo * dt;
        cur_b = bucket_of(t_lo, q.start, q.interval);
        const uint64_t ivl = (uint64_t)q.interval;
        const uint64_t sq64 = ivl / dtu;
        step_q = sq64 > 0xffffffffull ? 0xffffffffu : (uint32_t)sq64;
        step_r = ivl - sq64 * dtu;
        /* rb = ceil((W - t0)/dt), W = start + (cur_b+1)*interval > t_lo >= t0 */
        uint64_t D = (uint64_t)(q.start + (int64_t)(cur_b + 1) * q.interval - t0) + dtu - 1;
        uint64_t qq = D / dtu; rem = D - qq * dtu; rb = qq > 0xffffffffull ? 0xffffffffu : (uint32_t)qq;
    }

    /* ---- per-window partials ---- */
    double sum = 0.0, mn = 0.0, mx = 0.0; uint64_t fi = 0, lastv = 0; /* min/max live as doubles: aligned register pairs for DSETP */
    uint32_t n_mn = 0, n_mx = 0;  /* countdown value at the extreme row (row = stop - countdown) */
    uint32_t w_row0 = r_lo;       /* first row of the open window */
    bool head_done = false; uint32_t head_b = OG_NO_BUCKET;
    auto part_of = [&](int func, uint32_t stop, uint32_t cnt) -> Part {
        Part pp; pp.ok = cnt != 0; pp.v = 0; pp.t = 0;
        switch (func) {
        case OG_AGG_COUNT: pp.v = cnt; break;
        case OG_AGG_SUM: if (FM & FM_SUM) pp.v = d2u(sum); break;
        case OG_AGG_MIN: if (FM & FM_MIN) { pp.v = d2u(mn); if (TIMES) pp.t = t0 + (int64_t)(stop - n_mn) * dt; } break;
        case OG_AGG_MAX: if (FM & FM_MAX) { pp.v = d2u(mx); if (TIMES) pp.t = t0 + (int64_t)(stop - n_mx) * dt; } break;
        case OG_AGG_FIRST: if (FM & FM_FIRST) { pp.v = fi; pp.t = t0 + (int64_t)w_row0 * dt; } break;
        default: if (FM & FM_LAST) { pp.v = lastv; pp.t = t0 + (int64_t)(stop - 1) * dt; } break;
        }
        return pp;
    };
    auto flush = [&](uint32_t stop, bool final) { /* the window [w_row0, stop): head edge, tail edge or an interior cell */
        const uint32_t cnt = stop - w_row0;
        const int kind = !head_done ? 0 : final ? 1 : 2;
        const size_t idx = kind == 2 ? cell_idx(ch, series, cur_b) : e + kind;
#pragma unroll 1
        for (uint32_t c = 0; c < q.n_calls; c++) { /* rolled on purpose: this path runs once per window, keep it small */
            const Tri &dst = kind == 2 ? ch.cells[c] : ch.edges[c];
            store_part(dst, idx, part_of(q.calls[c].func, stop, cnt));
        }
        if (kind == 0) { head_done = true; head_b = cur_b; }
    };

    /* ---- decode state ---- */
    uint32_t p = 0;          /* bit position of the next unread bit of the lane's stream */
    uint32_t m = 64, tr = 0, sr = 0, kfast = 0, fast_ctrl = 5, bad = 0; uint64_t MASK = 0;
    bool done = !active;

    cp_async_wait<0>();
    uint64_t val = 0;
    if (active) { val = fetch64(col, 0); p = 64; } /* first value: 64 raw bits */

    /* ---- row events: skip rows before r_lo, window boundaries, end at r_hi ---- */
    bool skipping = r_lo > 0;
    uint32_t stop = skipping ? r_lo : (rb < r_hi + 1 ? rb : r_hi + 1); /* row index of the next event */
    uint32_t n_ev = stop; /* rows until the next event (row 0 is current) */
    if (!skipping) { fi = val; mn = mx = u2d(val); n_mn = n_mx = n_ev; }

    /* Refill schedule.  Service s (every K records) copies 8-word batches up to (p_s>>5) + NW (so at least up to
     * (p_s>>5) + NW - 7) and then waits for the PREVIOUS service's group.  Until service s+1 the lane reads at most
     * 77 + 64 bits past p_{s+1} <= p_{s-1} + 2*K*77, i.e. words <= (p_{s-1}>>5) + (2*8*77 + 141)/32 + 3 = +46 < NW - 7.
     *
     * Lanes that are finished (or never were active) keep executing the record decode on whatever their window holds —
     * it has no side effects — so the hot loop carries no per-lane "done" branch: they have n_ev = 2^32-1 (no event
     * for 2^32 records), copy nothing (words_lim = 0) and are ignored by the corrupt-page check. */
    constexpr uint32_t K = OG_IL_K;
    static_assert((2 * K * 77 + 13) / 32 + 4 <= OG_IL_NW - OG_IL_BATCH + 1, "window too small for the refill period");
    uint32_t words_lim = done ? 0u : n_words;
    if (done) n_ev = 0xffffffffu;
    auto on_event = [&]() {
        bool fin = false;
        if (skipping) { skipping = false; sum = 0.0; }
        else {
            flush(stop, stop > r_hi);
            fin = stop > r_hi;
            sum = 0.0;
            if (!fin) while (stop >= rb) { /* advance to the window that contains row `stop` (a loop: dt may exceed the interval) */
                cur_b++;
                rem += step_r; uint32_t adv = step_q;
                if (rem >= dtu) { rem -= dtu; adv++; }
                rb = (rb > 0xffffffffu - adv) ? 0xffffffffu : rb + adv;
            }
        }
        if (fin) { done = true; n_ev = 0xffffffffu; words_lim = 0; }
        else {
            w_row0 = stop;
            uint32_t nstop = rb < r_hi + 1 ? rb : r_hi + 1;
            n_ev = nstop - stop; stop = nstop;
            fi = val; mn = mx = u2d(val); n_mn = n_mx = n_ev; /* the first value of a window seeds min/max/first (column_util.go:190-278) */
        }
    };
    auto record = [&]() {
        /* accumulate the current row */
        if (FM & FM_SUM) sum = sum + u2d(val);
        if (FM & FM_MIN) { if (mn > u2d(val)) { mn = u2d(val); if (TIMES) n_mn = n_ev; } }
        if (FM & FM_MAX) { if (mx < u2d(val)) { mx = u2d(val); if (TIMES) n_mx = n_ev; } }
        if (FM & FM_LAST) lastv = val;
        n_ev--;
        /* next record (batch_float.go:352-508) */
        uint64_t x = fetch64(col, p);
        uint32_t ctrl = (uint32_t)(x >> 62);
        if (ctrl == fast_ctrl) {   /* '10' with a window that starts at or below bit 61: reuse it in place */
            val ^= (x >> sr) & MASK;
            p += kfast;
        } else if (ctrl < 2) {     /* '0': same value */
            p += 1;
        } else {
            if (ctrl == 3) {       /* '11': 5 bits leading, 6 bits meaningful */
                uint32_t lm = (uint32_t)(x >> 51) & 0x7ff;
                uint32_t lead = lm >> 6; m = lm & 0x3f;
                if (m == 0) { m = 64; tr = 0; }
                else { if (lead + m > 64) { if (!done) bad = 1; lead = 0; m = 64; } tr = 64 - lead - m; }
                p += 13;
                fast_ctrl = lead >= 2 ? 2u : 5u;
                sr = lead - 2; kfast = 2 + m;
                MASK = (m == 64 ? ~0ull : ((1ull << m) - 1)) << tr;
            } else p += 2;
            uint64_t y = fetch64(col, p);
            uint64_t sig = m == 64 ? y : (y >> (64 - m));
            p += m;
            val ^= sig << tr;
        }
    };
    for (;;) {
        if (__all_sync(FULL, done)) break;
        if (!done && (p >> 5) > n_words) { bad = 1; done = true; n_ev = 0xffffffffu; words_lim = 0; } /* ran past the stream: corrupt page */
        refill(min((p >> 5) + NW, words_lim));
        cp_async_commit();
        cp_async_wait<1>();
        /* When no lane reaches a window boundary within the next K records (lanes of a regular shard are in lockstep, so
         * this is 14 rounds out of 15 at 60 rows per window) the records run without the per-record event test. */
        if (__reduce_min_sync(FULL, n_ev) >= K) {
#pragma unroll 2
            for (uint32_t k = 0; k < K; k++) record();
        } else {
#pragma unroll 1
            for (uint32_t k = 0; k < K; k++) {
                if (n_ev == 0) on_event(); /* current row == stop */
                record();
            }
        }
    }
    if (active) {
        if (bad) report_err(ch.err, D_CORRUPT, seg);
        ch.edge_bucket[e] = head_b;
        ch.edge_bucket[e + 1] = (head_b == OG_NO_BUCKET || cur_b == head_b) ? OG_NO_BUCKET : cur_b;
    }
    /* copies still in flight must land before this CTA's shared memory can be reused */
    cp_async_wait<0>();
}

} // namespace ogpu
