/*
 * snappy_load.cuh — Snappy pages are transcoded to raw pages when a shard is opened.
 *
 * The reference picks Snappy for float segments with few decimals or a NaN (lib/compress/float.go:77-85,206-208,248-254:
 * block = [0x20][snappy block of the raw LE doubles], compress.go:123-144) and for time segments whose deltas fit neither
 * const-delta nor Simple8b (lib/encoding/timestamp.go:132-148,274-297: [0x30][u32 srcLen][u32 compLen][snappy block of the raw
 * LE int64 times]).  A Snappy block is a byte-serial LZ stream with no parallelism inside a page, and its output is the raw
 * column anyway — so the loader decodes each such page ONCE (one thread per page) into the page's raw form
 *     float : [header as found][0x00][n x 8 B LE]                        (floatCompressedNull, float.go:133-137)
 *     time  : [32][u32 rows][0x40][u32 8*rows][rows x u64 BE zigzag]     (unpackUncompressedData, timestamp.go:299-308)
 * appended behind the shard's data, and points the directory at it.  Every later reader (fused Gorilla kernel via the XOR-delta
 * repack, general kernel, tile decoders, og_decode_segment) sees standard raw pages.
 */
#pragma once
#include "agg_kernels.cuh"

namespace ogpu {

/* per (column slot c in 0..n_columns, segment): transcoded size, 0 = not a Snappy page.  One thread per segment. */
__global__ void k_snappy_scan(DirP d, const int32_t *col_types, uint32_t *tr_size, unsigned long long *tot /*[0] pages [1] old bytes [2] new bytes*/) {
    uint32_t seg = blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= d.n_segments) return;
    const size_t ti = (size_t)d.n_columns * d.n_segments + seg;
    const uint8_t *tp = d.data + d.page_off[ti];
    const uint32_t tlen = d.page_len[ti];
    uint32_t rows = 0;
    tr_size[ti] = 0;
    if (tlen >= 6 && __ldg(tp) == 32) {
        rows = ld_be32(tp + 1);
        if ((__ldg(tp + 5) >> 4) == 3 && rows < (1u << 27)) { tr_size[ti] = 10 + 8 * rows; atomicAdd(&tot[0], 1ull); atomicAdd(&tot[1], (unsigned long long)tlen); atomicAdd(&tot[2], 10ull + 8ull * rows); }
    } else if (tlen >= 1 && __ldg(tp) == 18) rows = 1;
    for (uint32_t c = 0; c < d.n_columns; c++) {
        const size_t pi = (size_t)c * d.n_segments + seg;
        tr_size[pi] = 0;
        const uint32_t len = d.page_len[pi];
        if (len == 0 || col_types[c] != OG_TYPE_FLOAT || rows == 0) continue;
        PageHdr h;
        if (parse_field_header(d.data + d.page_off[pi], len, OG_TYPE_FLOAT, rows, h) != D_OK) continue; /* k_validate reports it */
        if (h.one_row || h.nil_count >= h.rows || h.block_len < 1 || (__ldg(h.block) >> 4) != 2) continue;
        const uint32_t hdr = (uint32_t)(h.block - (d.data + d.page_off[pi]));
        const uint32_t sz = hdr + 1 + 8 * (h.rows - h.nil_count);
        tr_size[pi] = sz; atomicAdd(&tot[0], 1ull); atomicAdd(&tot[1], (unsigned long long)len); atomicAdd(&tot[2], (unsigned long long)sz);
    }
}

/* one thread per page with tr_size != 0: decode into new_data + tr_off[page], then repoint the directory */
__global__ void k_snappy_transcode(DirP d, const uint32_t *tr_size, const uint64_t *tr_off, uint8_t *new_data, uint64_t new_base,
                                   uint64_t *page_off, uint32_t *page_len, int *err) {
    const size_t n_pages = (size_t)(d.n_columns + 1) * d.n_segments;
    const size_t pi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= n_pages) return;
    const uint32_t sz = tr_size[pi];
    if (sz == 0) return;
    const uint32_t seg = (uint32_t)(pi % d.n_segments);
    const bool is_time = pi / d.n_segments == d.n_columns;
    const uint8_t *src = d.data + d.page_off[pi];
    const uint32_t len = d.page_len[pi];
    uint8_t *dst = new_data + new_base + tr_off[pi];
    int rc = D_OK; uint32_t got = 0;
    if (is_time) {
        const uint32_t rows = ld_be32(src + 1);
        if (len < 14) rc = D_CORRUPT;
        else {
            const uint32_t srcl = ld_be32(src + 6), compl_ = ld_be32(src + 10);
            if (srcl != 8 * rows || compl_ > len - 14) rc = D_CORRUPT;
            else rc = snappy_decode_dev(src + 14, compl_, dst + 10, 8 * rows, &got);
            if (rc == D_OK && got != 8 * rows) rc = D_CORRUPT;
        }
        if (rc == D_OK) {
            for (int i = 0; i < 5; i++) dst[i] = __ldg(src + i);
            dst[5] = 0x40;
            const uint32_t bl = 8 * rows;
            dst[6] = (uint8_t)(bl >> 24); dst[7] = (uint8_t)(bl >> 16); dst[8] = (uint8_t)(bl >> 8); dst[9] = (uint8_t)bl;
            for (uint32_t i = 0; i < rows; i++) { /* raw LE int64 -> zigzag big-endian, in place */
                uint8_t *p = dst + 10 + 8 * (size_t)i;
                uint64_t v = 0;
                for (int k = 7; k >= 0; k--) v = (v << 8) | p[k];
                const uint64_t z = (v << 1) ^ (uint64_t)((int64_t)v >> 63);
                for (int k = 0; k < 8; k++) p[k] = (uint8_t)(z >> (56 - 8 * k));
            }
        }
    } else {
        PageHdr h;
        rc = parse_field_header(src, len, OG_TYPE_FLOAT, d.seg_rows ? ld_be32(d.data + d.page_off[(size_t)d.n_columns * d.n_segments + seg] + 1) : 0, h);
        if (rc == D_OK) {
            const uint32_t hdr = (uint32_t)(h.block - src), n = h.rows - h.nil_count;
            for (uint32_t i = 0; i < hdr; i++) dst[i] = __ldg(src + i);
            dst[hdr] = 0x00;
            rc = snappy_decode_dev(h.block + 1, h.block_len - 1, dst + hdr + 1, 8 * n, &got);
            if (rc == D_OK && got != 8 * n) rc = D_CORRUPT;
        }
    }
    if (rc != D_OK) { report_err(err, rc, seg); return; }
    page_off[pi] = new_base + tr_off[pi];
    page_len[pi] = sz;
}

} // namespace ogpu
