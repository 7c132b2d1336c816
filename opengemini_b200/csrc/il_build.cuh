/*
 * il_build.cuh — builds the lane-interleaved, length-binned stream copy that k_fused_il reads (once per shard and column).
 *
 *   k_il_scan        per segment: static class (Gorilla stream / raw page / not eligible), stream length in words, the
 *                    const-delta time page's (t0, dt), and the sort key (domain, words).  A domain is the set of segments that
 *                    may share a lane group: on regular shards (every series has J segments) segment index j of a block of
 *                    OG_IL_SUPER consecutive series — they cover the same time range, so their windows coincide; otherwise
 *                    the whole shard.
 *   radix sort       (cub::DeviceRadixSort, stable) orders the eligible segments by (domain, words): 32 consecutive entries of
 *                    one domain make a lane group of similar stream lengths.
 *   k_il_assign      sorted position -> (group, lane) slot; writes the per-lane metadata the kernel needs (segment, rows,
 *                    series, t0, dt) as coalesced arrays.
 *   k_il_group_rows  rows of a group = longest lane + pad, rounded up to the bulk-copy batch.
 *   k_il_repack      word w of lane l -> il[grp_off + w*32 + l], big-endian stream words stored in native order; raw pages
 *                    (float.go:96-99: Gorilla output above 90 % of raw) become [v0][v1^v0][v2^v1]... so that the Gorilla loop
 *                    decodes them as 64-bit XOR records without control bits.
 */
#pragma once
#include "fused_il.cuh"

namespace ogpu {

#define OG_IL_SUPER 4096u      /* series per binning domain (chunks of series are multiples of it when possible) */
#define OG_IL_WORD_BITS 24u    /* sort key = domain << 24 | words */

struct IlScanOut {
    uint8_t *ok;            /* [n_segments] SEG_* */
    uint32_t *seg_words;    /* [n_segments] stream words incl. pad (0 = not eligible) */
    int64_t *seg_t0;        /* [n_segments] */
    uint64_t *seg_dt;
    uint64_t *keys;         /* [n_segments] sort key, ~0 = not eligible */
    uint32_t *vals;         /* [n_segments] = segment id */
    uint32_t *dom_cnt;      /* [n_domains] eligible segments per domain */
};

__global__ void k_il_scan(DirP d, int col, int col_type, uint32_t J, IlScanOut o) {
    uint32_t seg = blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= d.n_segments) return;
    uint8_t c = SEG_GENERAL; uint32_t nw = 0; int64_t t0 = 0; uint64_t dt = 0;
    const uint32_t rows = d.seg_rows[seg];
    if (rows >= 2 && rows < (1u << 22) && col_type == OG_TYPE_FLOAT) {
        size_t pi = (size_t)col * d.n_segments + seg, ti = (size_t)d.n_columns * d.n_segments + seg;
        const uint8_t *p = d.data + d.page_off[pi], *t = d.data + d.page_off[ti];
        uint32_t len = d.page_len[pi], tlen = d.page_len[ti];
        /* time page: [32][u32 rows][0x10][t0][uvarint dt][uvarint n-1] */
        TimeDesc td;
        if (len >= 16 && tlen >= 16 && __ldg(p) == 31 && __ldg(t) == 32 && (__ldg(t + 5) >> 4) == 1 && ld_be32(p + 1) == rows &&
            parse_time_page(t, tlen, td) == D_OK && td.kind == 0 && td.delta > 0 && td.delta < (1ull << 40)) {
            const int tag = __ldg(p + 5) >> 4;
            if (tag == 3 && __ldg(p + 6) == 0x10) { /* value page: [31][u32 rows][0x30][0x10][8 B first]... */
                c = SEG_FAST; nw = (len - OG_IL_HDR + 3) / 4 + OG_IL_PAD_WORDS;
            } else if (tag == 0 && len == OG_IL_RAW_HDR + 8 * (size_t)rows) { /* raw page: [31][u32 rows][0x00][rows x 8 B LE] */
                c = SEG_RAWX; nw = 2 * rows + OG_IL_PAD_WORDS;
            }
            if (nw >= (1u << OG_IL_WORD_BITS)) { c = SEG_GENERAL; nw = 0; }
            t0 = td.t0; dt = td.delta;
        }
    }
    o.ok[seg] = c; o.seg_words[seg] = nw; o.seg_t0[seg] = t0; o.seg_dt[seg] = dt; o.vals[seg] = seg;
    if (c == SEG_GENERAL) { o.keys[seg] = ~0ull; return; }
    const uint32_t series = d.seg_series[seg];
    const uint32_t dom = J ? (series / OG_IL_SUPER) * J + (seg - d.series_seg_begin[series]) : 0u;
    o.keys[seg] = ((uint64_t)dom << OG_IL_WORD_BITS) | nw;
    atomicAdd(&o.dom_cnt[dom], 1u);
}

struct IlAssign {
    const uint64_t *keys; const uint32_t *segs;    /* sorted */
    const uint32_t *elem_first, *grp_first;        /* [n_domains] first sorted position / first group of each domain */
    const uint32_t *seg_words; const int64_t *seg_t0; const uint64_t *seg_dt; const uint8_t *ok;
    uint32_t *lane_seg, *lane_rows, *lane_series, *grp_col; int64_t *lane_t0; uint64_t *lane_dt;
    uint32_t n_elig, J, cols_per_super;
};
__global__ void k_il_assign(DirP d, IlAssign a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_elig) return;
    const uint32_t seg = a.segs[i], dom = (uint32_t)(a.keys[i] >> OG_IL_WORD_BITS);
    const uint32_t rank = i - a.elem_first[dom], g = a.grp_first[dom] + rank / 32;
    const size_t slot = (size_t)g * 32 + (rank & 31);
    a.lane_seg[slot] = seg;
    a.lane_rows[slot] = d.seg_rows[seg] | (a.ok[seg] == SEG_RAWX ? OG_IL_RAWFLAG : 0u);
    a.lane_series[slot] = d.seg_series[seg];
    a.lane_t0[slot] = a.seg_t0[seg]; a.lane_dt[slot] = a.seg_dt[seg];
    if ((rank & 31) == 0) a.grp_col[g] = (a.J ? dom / a.J : 0u) * a.cols_per_super + rank / 32;
}

/* rows of every lane group = max over its lanes, rounded up to the bulk-copy batch (one warp per group) */
__global__ void k_il_group_rows(uint32_t n_groups, const uint32_t *lane_seg, const uint32_t *seg_words, uint32_t *grp_rows) {
    uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g >= n_groups) return;
    uint32_t seg = lane_seg[(size_t)g * 32 + lane];
    uint32_t w = seg != OG_IL_NONE ? seg_words[seg] : 0;
#pragma unroll
    for (int o = 16; o; o >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, o));
    if (lane == 0) grp_rows[g] = (w + OG_IL_B - 1) / OG_IL_B * OG_IL_B;
}

/* the repack (one warp per group; every store is one full 128-byte row) */
__global__ void k_il_repack(DirP d, int col, const uint8_t *ok, const uint32_t *lane_seg, const uint64_t *grp_off, const uint32_t *grp_rows,
                            uint32_t n_groups, uint32_t *il) {
    uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g >= n_groups) return;
    const uint32_t nw = grp_rows[g];
    if (nw == 0) return;
    const uint32_t seg = lane_seg[(size_t)g * 32 + lane];
    const bool live = seg != OG_IL_NONE;
    const bool raw = live && ok[seg] == SEG_RAWX;
    const uint32_t *base = nullptr; uint32_t sh = 0, own_words = 0;
    if (live) {
        size_t pi = (size_t)col * d.n_segments + seg;
        const uint8_t *s = d.data + d.page_off[pi] + (raw ? OG_IL_RAW_HDR : OG_IL_HDR);
        base = (const uint32_t *)((uintptr_t)s & ~(uintptr_t)3);
        sh = (uint32_t)((uintptr_t)s & 3);
        own_words = raw ? 2 * d.seg_rows[seg] : (d.page_len[pi] - OG_IL_HDR + 3) / 4 + OG_IL_PAD_WORDS; /* bytes past the page are the next page or the shard's tail padding */
    }
    uint32_t *out = il + grp_off[g] + lane;
    uint32_t a = live ? __ldg(base) : 0;
    if (!raw) {
        /* bytes s[4w..4w+3] big-endian: from aligned words a=base[w], b=base[w+1] (little-endian loads) */
        const uint32_t sel = sh == 0 ? 0x0123u : sh == 1 ? 0x1234u : sh == 2 ? 0x2345u : 0x3456u;
        for (uint32_t w = 0; w < nw; w++) {
            uint32_t v = 0;
            if (w < own_words) {
                uint32_t b = __ldg(base + w + 1);
                v = __byte_perm(a, b, sel);
                a = b;
            }
            out[(size_t)w * 32] = v;
        }
    } else {
        /* little-endian doubles: value i = bytes s[8i..8i+7]; emit hi word then lo word of v_i ^ v_{i-1} */
        const uint32_t sel = sh == 0 ? 0x3210u : sh == 1 ? 0x4321u : sh == 2 ? 0x5432u : 0x6543u;
        uint32_t plo = 0, phi = 0;
        for (uint32_t w = 0; w < nw; w += 2) {
            uint32_t hi = 0, lo = 0;
            if (w < own_words) {
                uint32_t b = __ldg(base + w + 1), c = __ldg(base + w + 2);
                const uint32_t vlo = __byte_perm(a, b, sel), vhi = __byte_perm(b, c, sel);
                a = c;
                hi = vhi ^ phi; lo = vlo ^ plo; phi = vhi; plo = vlo;
            }
            out[(size_t)w * 32] = hi;
            if (w + 1 < nw) out[(size_t)(w + 1) * 32] = lo;
        }
    }
}

} // namespace ogpu
