/*
 * fused_fast.cuh — K5 specialised for the headline shape: float64 Gorilla pages (tag 3, no nulls) with const-delta
 * time pages.  One thread per segment (a warp = 32 consecutive segments), built for instruction count and for the
 * memory system:
 *
 *   staging   each lane owns a 256-byte ring (4 slots x 64 B, + 8 mirror bytes) in shared memory that holds the next
 *             bytes of ITS page.  The warp refills rings cooperatively: whenever some lanes ("owners") have a free
 *             slot, every group of 8 lanes copies one owner's next 64-byte chunk with one 8-byte cp.async each
 *             (LDGSTS: fully coalesced 64 B segments, no register staging, asynchronous), and signals that owner's
 *             per-slot mbarrier (cp.async.mbarrier.arrive).  Owners wait on their own mbarrier just before they enter
 *             a chunk — two chunks after it was requested — so DRAM latency is off the critical path.
 *             The 264-byte ring stride skews lanes by 8 bytes: LDS from lanes that run in lockstep are 2-way at worst.
 *   decode    stateless bit addressing: the 64 bits at bit position p come from three LDS.32 + three PRMT (byte swap)
 *             + two funnel shifts (the mirror bytes make the three words never wrap); the '10' (window reuse) record —
 *             >95% of records on noisy-mantissa data — is then one shift + mask + xor and p += 2+m.
 *   reduce    window boundaries are row countdowns derived from the const-delta time page (no time decode and no
 *             division in the loop); partials stay in registers and are flushed to the same edge/cell arrays the
 *             general kernel uses, so k_fix_edges / k_merge_groups are shared and float sums keep the reference's
 *             left-to-right order.
 *
 * A TMA variant of the staging (per-lane cp.async.bulk, UBLKCP) was measured first: the uniform-datapath waterfall
 * (9 warp instructions per 128-byte copy) plus divergent per-lane service made it 3x more instructions per value;
 * see profiles/r01_fast_kernel_history.md.
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
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
/* 8-byte asynchronous global->shared copy (LDGSTS) and its completion hook on an mbarrier */
__device__ __forceinline__ void cp_async8(uint32_t dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory"); }
template <int OFF> __device__ __forceinline__ uint32_t lds32o(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF)); return v; }
__device__ __forceinline__ uint64_t lds64(uint32_t a) { uint64_t v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a)); return v; }

enum { FM_COUNT = 1, FM_SUM = 2, FM_MIN = 4, FM_MAX = 8, FM_FIRST = 16, FM_LAST = 32 };
enum { SEG_GENERAL = 0, SEG_FAST = 1 };  /* per-segment classes computed by k_classify */

#define OG_FAST_THREADS 128
#define OG_FAST_CH 64u         /* bytes per staged chunk */
#define OG_FAST_NS 4u          /* ring slots per lane */
#define OG_FAST_RING 256u      /* ring bytes per lane */
#define OG_FAST_STRIDE 264u    /* ring + 8 mirror bytes; 8-byte skew between lanes */

/* per-segment eligibility for the fast kernel (one thread per segment, header bytes only; run once per query plan) */
__global__ void k_classify(DirP d, QueryP q, uint8_t *cls) {
    uint32_t seg = blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= d.n_segments) return;
    uint8_t c = SEG_GENERAL;
    uint32_t rows = d.seg_rows[seg];
    if (!(d.seg_tmax[seg] < q.tmin || d.seg_tmin[seg] > q.tmax || rows < 2) && q.col_type[0] == OG_TYPE_FLOAT) {
        size_t pi = (size_t)q.col_index[0] * d.n_segments + seg, ti = (size_t)d.n_columns * d.n_segments + seg;
        const uint8_t *p = d.data + d.page_off[pi], *t = d.data + d.page_off[ti];
        uint32_t len = d.page_len[pi], tlen = d.page_len[ti];
        /* value page: [31][u32 rows][0x30][0x10][8 B first]...; time page: [32][u32 rows][0x10][t0][uvarint dt][uvarint n-1] */
        if (len >= 16 && tlen >= 16 && __ldg(p) == 31 && (__ldg(p + 5) >> 4) == 3 && __ldg(t) == 32 && (__ldg(t + 5) >> 4) == 1) {
            TimeDesc td;
            if (parse_time_page(t, tlen, td) == D_OK && td.kind == 0 && td.delta > 0 && td.delta < (1ull << 40) && ld_be32(p + 1) == rows) c = SEG_FAST;
        }
    }
    cls[seg] = c;
}

/* 64 bits of the stream at bit position p, from the lane's ring (ring + mirror: the three words never wrap) */
__device__ __forceinline__ uint64_t fetch64(uint32_t ring, uint32_t p) {
    uint32_t a = ring + ((p >> 3) & (OG_FAST_RING - 4));
    uint32_t a0 = lds32o<0>(a), a1 = lds32o<4>(a), a2 = lds32o<8>(a);
    a0 = __byte_perm(a0, 0, 0x0123); a1 = __byte_perm(a1, 0, 0x0123); a2 = __byte_perm(a2, 0, 0x0123);
    uint32_t hi = __funnelshift_l(a1, a0, p), lo = __funnelshift_l(a2, a1, p); /* shift amount taken mod 32 */
    return ((uint64_t)hi << 32) | lo;
}

template <int FM, bool TIMES>
__global__ void __launch_bounds__(OG_FAST_THREADS) k_fused_fast(DirP d, QueryP q, ChunkP ch, const uint8_t *cls) {
    constexpr uint32_t CH = OG_FAST_CH, NS = OG_FAST_NS, STRIDE = OG_FAST_STRIDE;
    constexpr uint32_t AHEAD = 24; /* bytes past the read position that a record decode may touch (13 + 64 bits, word-granular loads) */
    constexpr uint32_t FULL = 0xffffffffu;
    __shared__ __align__(128) uint8_t s_ring[OG_FAST_THREADS * STRIDE];
    __shared__ __align__(8) uint64_t s_bar[OG_FAST_THREADS * NS];
    __shared__ __align__(8) uint64_t s_src[OG_FAST_THREADS]; /* 8-byte aligned stream base of every lane */

    const uint32_t lane = threadIdx.x & 31, wbase = threadIdx.x & ~31u;
    const uint32_t seg = ch.seg_begin + blockIdx.x * blockDim.x + threadIdx.x;
    bool active = seg < ch.seg_end && cls[seg < ch.seg_end ? seg : ch.seg_begin] == SEG_FAST;
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

    /* ---- stream: page = [31][rows u32][0x30][0x10] | first value 8 B BE | records... ---- */
    uint32_t skip = 0, n_chunks = 0;
    {
        const uint8_t *g8 = d.data;
        if (active) {
            const size_t pi = (size_t)q.col_index[0] * d.n_segments + seg;
            const uint8_t *stream = d.data + d.page_off[pi] + 7;
            const uint32_t stream_len = d.page_len[pi] - 7;
            g8 = (const uint8_t *)((uintptr_t)stream & ~(uintptr_t)7);
            skip = (uint32_t)(stream - g8);
            n_chunks = (skip + stream_len + AHEAD + CH - 1) / CH;
        }
        s_src[threadIdx.x] = (uint64_t)(uintptr_t)g8;
    }
    const uint32_t ring = opaque(smem_u32(s_ring) + threadIdx.x * STRIDE);
    const uint32_t bar0 = opaque(smem_u32(s_bar) + threadIdx.x * NS * 8);
    /* every chunk is written by 8 cp.async (one per serving lane) + the owner's own arrival (mirror copy or plain) */
#pragma unroll
    for (uint32_t i = 0; i < NS; i++) mbar_init(bar0 + 8 * i, 9);
    __syncwarp();

    /* serving side, fixed mapping: in step j this lane copies piece (lane&7) of the chunk of owner 4j + (lane>>3) */
    const uint32_t sv_dst0 = opaque(smem_u32(s_ring) + (wbase + (lane >> 3)) * STRIDE + 8 * (lane & 7));
    const uint32_t sv_bar0 = opaque(smem_u32(s_bar) + (wbase + (lane >> 3)) * NS * 8);
    const uint32_t sv_src0 = opaque(smem_u32(s_src) + (wbase + (lane >> 3)) * 8);
    const uint32_t my_src = opaque(smem_u32(s_src) + threadIdx.x * 8);
    uint32_t staged = 0; /* chunks requested so far by this lane */
    /* one cooperative round: every lane with want!=0 gets chunk `staged` copied into slot staged%NS */
    auto coop_round = [&](bool want) {
        const uint32_t req = want ? staged : FULL;
        const uint32_t wm = __ballot_sync(FULL, want);
#pragma unroll
        for (uint32_t j = 0; j < 8; j++) {
            if (((wm >> (4 * j)) & 0xfu) == 0) continue; /* warp-uniform */
            uint32_t c = __shfl_sync(FULL, req, 4 * j + (lane >> 3));
            if (c != FULL) {
                uint64_t base = lds64(sv_src0 + 32 * j);
                uint32_t sl = c % NS;
                cp_async8(sv_dst0 + j * (4 * STRIDE) + sl * CH, (const uint8_t *)(uintptr_t)base + (size_t)c * CH + 8 * (lane & 7));
                cp_async_arrive(sv_bar0 + j * (4 * NS * 8) + sl * 8);
            }
        }
        if (want) {
            uint32_t sl = staged % NS;
            if (sl == 0) { /* mirror: the first 8 bytes of slot 0 again after the ring end */
                cp_async8(ring + OG_FAST_RING, (const uint8_t *)(uintptr_t)lds64(my_src) + (size_t)staged * CH);
                cp_async_arrive(bar0);
            } else mbar_arrive(bar0 + 8 * sl);
            staged++;
        }
    };
#pragma unroll
    for (uint32_t i = 0; i < NS; i++) coop_round(active && staged < n_chunks);

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
    double sum = 0.0; uint64_t mn = 0, mx = 0, fi = 0, lastv = 0;
    uint32_t n_mn = 0, n_mx = 0;  /* countdown value at the extreme row (row = stop - countdown) */
    uint32_t w_row0 = r_lo;       /* first row of the open window */
    bool head_done = false; uint32_t head_b = OG_NO_BUCKET;
    auto part_of = [&](int func, uint32_t stop, uint32_t cnt) -> Part {
        Part pp; pp.ok = cnt != 0; pp.v = 0; pp.t = 0;
        switch (func) {
        case OG_AGG_COUNT: pp.v = cnt; break;
        case OG_AGG_SUM: if (FM & FM_SUM) pp.v = d2u(sum); break;
        case OG_AGG_MIN: if (FM & FM_MIN) { pp.v = mn; if (TIMES) pp.t = t0 + (int64_t)(stop - n_mn) * dt; } break;
        case OG_AGG_MAX: if (FM & FM_MAX) { pp.v = mx; if (TIMES) pp.t = t0 + (int64_t)(stop - n_mx) * dt; } break;
        case OG_AGG_FIRST: if (FM & FM_FIRST) { pp.v = fi; pp.t = t0 + (int64_t)w_row0 * dt; } break;
        default: if (FM & FM_LAST) { pp.v = lastv; pp.t = t0 + (int64_t)(stop - 1) * dt; } break;
        }
        return pp;
    };
    auto flush = [&](uint32_t stop, bool final) { /* the window [w_row0, stop) */
        uint32_t cnt = stop - w_row0;
        if (!head_done) {
#pragma unroll
            for (uint32_t c = 0; c < OG_MAX_CALLS; c++) if (c < q.n_calls) store_part(ch.edges[c], e, part_of(q.calls[c].func, stop, cnt));
            head_done = true; head_b = cur_b;
        } else if (final) {
#pragma unroll
            for (uint32_t c = 0; c < OG_MAX_CALLS; c++) if (c < q.n_calls) store_part(ch.edges[c], e + 1, part_of(q.calls[c].func, stop, cnt));
        } else {
            size_t ci = (size_t)(series - ch.series_begin) * q.n_buckets + cur_b;
#pragma unroll
            for (uint32_t c = 0; c < OG_MAX_CALLS; c++) if (c < q.n_calls) store_part(ch.cells[c], ci, part_of(q.calls[c].func, stop, cnt));
        }
    };

    /* ---- decode state ---- */
    uint32_t p = skip * 8;   /* bit position of the next unread bit (relative to the 8-byte aligned stream base) */
    uint32_t landed = 0;     /* chunks [0, landed) have been waited for */
    uint32_t m = 64, tr = 0, sr = 0, kfast = 0; uint64_t MASK = 0; bool fastok = false, bad = false;
    bool done = !active;
    /* Service runs on a fixed schedule — every K iterations, for the whole warp — because per-lane triggers would fire
     * in almost every iteration (32 unsynchronised lanes, one event per ~10 records each).  A record consumes at most
     * 77 bits, so between two services a lane touches at most [b, b + 10*K + AHEAD): that much must have landed. */
    constexpr uint32_t K = 8, SPAN = 10 * K + AHEAD;
    auto lane_wait_ahead = [&]() {
        uint32_t b = p >> 3;
        if (b > skip + (n_chunks * CH)) { bad = true; done = true; return; } /* ran past the page: corrupt stream */
        uint32_t target = (b + SPAN + CH - 1) / CH;
        if (target > staged) target = staged;
        while (landed < target) { mbar_wait(bar0 + 8 * (landed % NS), (landed / NS) & 1); landed++; }
    };
    auto want_now = [&]() -> bool { return !done && staged < n_chunks && (p >> 3) >= (staged + 1 - NS) * CH; };

    uint64_t val = 0;
    if (active) { lane_wait_ahead(); val = fetch64(ring, p); p += 64; } /* first value: 64 raw bits */

    /* ---- row events: skip rows before r_lo, window boundaries, end at r_hi ---- */
    bool skipping = r_lo > 0;
    uint32_t stop = skipping ? r_lo : (rb < r_hi + 1 ? rb : r_hi + 1); /* row index of the next event */
    uint32_t n_ev = stop; /* rows until the next event (row 0 is current) */
    if (!skipping) { mn = mx = fi = val; n_mn = n_mx = n_ev; }

    for (uint32_t it = 0;; it++) {
        /* ---- warp service: staging rounds + landing waits (warp-uniform schedule) ---- */
        if ((it & (K - 1)) == 0) {
            for (;;) {
                bool w = want_now();
                if (!__any_sync(FULL, w)) break;
                coop_round(w);
            }
            if (!done) lane_wait_ahead();
            if (__all_sync(FULL, done)) break;
        }
        if (done) continue;
        if (n_ev == 0) { /* current row == stop */
            if (skipping) { skipping = false; sum = 0.0; }
            else {
                flush(stop, stop > r_hi);
                if (stop > r_hi) { done = true; continue; }
                sum = 0.0;
                while (stop >= rb) { /* advance to the window that contains row `stop` (a loop: dt may exceed the interval) */
                    cur_b++;
                    rem += step_r; uint32_t adv = step_q;
                    if (rem >= dtu) { rem -= dtu; adv++; }
                    rb = (rb > 0xffffffffu - adv) ? 0xffffffffu : rb + adv;
                }
            }
            w_row0 = stop;
            uint32_t nstop = rb < r_hi + 1 ? rb : r_hi + 1;
            n_ev = nstop - stop; stop = nstop;
            mn = mx = fi = val; n_mn = n_mx = n_ev; /* the first value of a window seeds min/max/first (column_util.go:190-278) */
        }
        /* accumulate the current row */
        if (FM & FM_SUM) sum = sum + u2d(val);
        if (FM & FM_MIN) { if (u2d(mn) > u2d(val)) { mn = val; if (TIMES) n_mn = n_ev; } }
        if (FM & FM_MAX) { if (u2d(mx) < u2d(val)) { mx = val; if (TIMES) n_mx = n_ev; } }
        if (FM & FM_LAST) lastv = val;
        n_ev--;
        /* next record (batch_float.go:352-508); service guarantees AHEAD readable bytes at p */
        uint64_t x = fetch64(ring, p);
        uint32_t ctrl = (uint32_t)(x >> 62);
        if (ctrl == 2 && fastok) { /* '10': reuse the window */
            val ^= (x >> sr) & MASK;
            p += kfast;
        } else if (ctrl < 2) {     /* '0': same value */
            p += 1;
        } else {
            if (ctrl == 3) {       /* '11': 5 bits leading, 6 bits meaningful */
                uint32_t lm = (uint32_t)(x >> 51) & 0x7ff;
                uint32_t lead = lm >> 6; m = lm & 0x3f;
                if (m == 0) { m = 64; tr = 0; } else { if (lead + m > 64) { bad = true; done = true; continue; } tr = 64 - lead - m; }
                p += 13;
                fastok = lead >= 2;
                sr = lead - 2; kfast = 2 + m;
                MASK = (m == 64 ? ~0ull : ((1ull << m) - 1)) << tr;
            } else p += 2;
            uint64_t y = fetch64(ring, p);
            uint64_t sig = m == 64 ? y : (y >> (64 - m));
            p += m;
            val ^= sig << tr;
        }
    }
    if (active) {
        if (bad) report_err(ch.err, D_CORRUPT, seg);
        ch.edge_bucket[e] = head_b;
        ch.edge_bucket[e + 1] = (head_b == OG_NO_BUCKET || cur_b == head_b) ? OG_NO_BUCKET : cur_b;
    }
    /* copies still in flight must land before this CTA's shared memory can be reused */
    asm volatile("cp.async.wait_all;" ::: "memory");
}

} // namespace ogpu
