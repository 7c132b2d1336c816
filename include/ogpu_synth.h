/*
 * ogpu_synth.h — definition of the synthetic column distributions as pure functions of
 * (seed, column, series, row).  Shared verbatim by the device generator (og_shard_synth), the CPU oracle's
 * shard builder and the tests, so that "identical synthetic TSM shards" is true by construction.
 *
 * Counter-based (SplitMix64 finaliser over a mixed key) so any row can be produced independently.
 *
 *   OG_SYNTH_F_HI     v = 100 + U[0,1) with a full 52-bit mantissa tail  -> Gorilla ~6 B/value (HBM-bound case)
 *   OG_SYNTH_F_LO     integer-valued walk: v = 1000 + W(seg) + sum of steps in {-2..2} inside the segment
 *                     -> intOnly stays true (lib/compress/float.go:244,206) -> Gorilla ~1-2 B/value
 *   OG_SYNTH_INT_WALK int64 walk, steps in [-1000,1000], restarted per segment -> Simple8b (int.go:123-134)
 *   OG_SYNTH_BOOL     Bernoulli(0.5) -> bit-packed (bool.go:40-61)
 */
#ifndef OGPU_SYNTH_H
#define OGPU_SYNTH_H

#include <stdint.h>

#if defined(__CUDACC__)
#define OG_HD __host__ __device__ __forceinline__
#else
#define OG_HD static inline
#endif

OG_HD uint64_t og_mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

/* one 64-bit draw per (seed, column, series, row, lane) */
OG_HD uint64_t og_synth_draw(uint64_t seed, uint32_t column, uint32_t series, uint64_t row, uint32_t lane) {
    uint64_t k = og_mix64(seed ^ (0xD6E8FEB86659FD93ull * (uint64_t)(column + 1)));
    k = og_mix64(k ^ (0xA0761D6478BD642Full * (uint64_t)(series + 1)) ^ ((uint64_t)lane << 56));
    return og_mix64(k + row);
}

OG_HD int og_synth_is_null(uint64_t seed, uint32_t column, uint32_t series, uint64_t row, uint32_t null_permille) {
    if (null_permille == 0) return 0;
    return (og_synth_draw(seed, column, series, row, 1) % 1000ull) < (uint64_t)null_permille;
}

OG_HD double og_synth_f_hi(uint64_t seed, uint32_t column, uint32_t series, uint64_t row) {
    uint64_t h = og_synth_draw(seed, column, series, row, 0);
    return 100.0 + (double)(h >> 11) * (1.0 / 9007199254740992.0); /* 2^-53 */
}

/* walk restart offset of a segment */
OG_HD int64_t og_synth_seg_base(uint64_t seed, uint32_t column, uint32_t series, uint64_t seg) {
    return (int64_t)(og_synth_draw(seed, column, series, seg, 2) % 1001ull) - 500;
}
OG_HD int64_t og_synth_f_lo_step(uint64_t seed, uint32_t column, uint32_t series, uint64_t row) {
    return (int64_t)(og_synth_draw(seed, column, series, row, 0) % 5ull) - 2;
}
OG_HD int64_t og_synth_int_step(uint64_t seed, uint32_t column, uint32_t series, uint64_t row) {
    return (int64_t)(og_synth_draw(seed, column, series, row, 0) % 2001ull) - 1000;
}
OG_HD int og_synth_bool(uint64_t seed, uint32_t column, uint32_t series, uint64_t row) {
    return (int)(og_synth_draw(seed, column, series, row, 0) & 1ull);
}

/*
 * Sequential generator of one segment's rows [row0, row0+n): state carried in *acc.
 *   F_LO / INT_WALK: value(row0) = 1000|0 + seg_base; value(r) = value(r-1) + step(r) for r > row0.
 * Call og_synth_walk_first for the first row of the segment, og_synth_walk_next for the rest.
 */
OG_HD int64_t og_synth_walk_first(uint64_t seed, uint32_t column, uint32_t series, uint64_t seg, int is_float_lo) {
    return (is_float_lo ? 1000 : 0) + og_synth_seg_base(seed, column, series, seg);
}

#endif /* OGPU_SYNTH_H */
