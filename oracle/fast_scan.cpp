/*
 * fast_scan.cpp — CPU BASELINE leg (test infrastructure, see og_oracle.h; used by bench.py's reference arm and
 * `cpu_baseline` only, never by the product).
 *
 * The checker in codec.cpp reads Gorilla streams one bit at a time, which is fine for a checker and unfair as "the
 * reference's CPU path".  This file restates the path the way the reference actually runs it for the headline query shape
 * (one float64 column, no WHERE, one tagset, count/sum/min/max pushed down):
 *   - tsm1.FloatArrayDecodeAll's batch decode with a 64-bit cached bit reader (batch_float.go:278-514, brCachedVal /
 *     brValidBits :308-347): whole segment -> []float64;
 *   - Time.constDeltaDecoding in closed form (timestamp.go:190-225);
 *   - FilterByTime (reader.go:754-771), getIntervalIndex (aggregate_cursor.go:343-356), the per-window reduce loops
 *     (series_agg_func.gen.go:24-162) with prevBuf/currBuf stitching (series_agg_reducer.gen.go:228-266);
 *   - AggTagSetCursor's update of the dense interval record in series order (reccord_functions.go:586-760).
 * Workers stride the series like group cursors (file_cursor.go:190-195) and their partial records are merged at the end.
 * Results are identical to scan_aggregate's (bitwise with one worker): tests/test_oracle_fast_scan.py.
 */
#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>

#include "og_oracle.h"

namespace ogo {

namespace {

inline uint64_t be64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return __builtin_bswap64(v); }
inline uint32_t be32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return __builtin_bswap32(v); }

/* 64-bit cached reader over a byte stream padded by the caller (>= 16 readable bytes past `end`) */
struct BitCache {
    const uint8_t *p; uint64_t cache; int valid; /* `valid` most significant bits of cache are unread */
    inline void init(const uint8_t *s) { p = s; cache = 0; valid = 0; }
    inline void refill() { /* top up to >= 57 valid bits */
        const uint64_t w = be64(p);
        cache |= valid ? (w >> valid) : w;
        const int take = (64 - valid) >> 3; /* whole bytes that fit */
        p += take; valid += take * 8;
    }
    inline uint64_t take(int k) { /* 1 <= k <= 56 */
        if (valid < k) refill();
        const uint64_t v = cache >> (64 - k);
        cache <<= k; valid -= k;
        return v;
    }
    inline uint64_t take64(int k) { /* 1 <= k <= 64 */
        if (k <= 56) return take(k);
        const uint64_t hi = take(k - 32);
        return (hi << 32) | take(32);
    }
};

/* FloatArrayDecodeAll: [0x10][8 B BE first][records...] terminated by the NaN sentinel; returns the value count */
inline int gorilla_batch(const uint8_t *b, size_t len, double *out, int cap) {
    if (len < 9) return 0;
    uint64_t val = be64(b + 1);
    if (val == 0x7FF8000000000001ull) return 0;
    int n = 0;
    memcpy(&out[n++], &val, 8);
    BitCache br; br.init(b + 9);
    const uint8_t *end = b + len;
    unsigned trailing = 0, meaningful = 64;
    while (n < cap) {
        if (br.p > end + 8) return -1; /* ran past the stream without meeting the sentinel */
        if (br.take(1)) {
            if (br.take(1)) {
                const unsigned lm = (unsigned)br.take(11);
                const unsigned leading = (lm >> 6) & 0x1f;
                meaningful = lm & 0x3f;
                if (meaningful > 0) trailing = 64 - leading - meaningful;
                else { trailing = 0; meaningful = 64; }
            }
            val ^= br.take64((int)meaningful) << (trailing & 0x3f);
            if (val == 0x7FF8000000000001ull) break;
        }
        memcpy(&out[n++], &val, 8);
    }
    return n;
}

struct Cell { double sum = 0.0; int64_t cnt = 0; double mn = 0.0, mx = 0.0; bool ok = false; };

} // namespace

int fast_scan_aggregate(const og_shard_desc &sh, const og_query_desc &q, int threads, uint32_t s_begin, uint32_t s_end, ScanResult &out) {
    if (q.n_filter || q.group_mode != OG_GROUP_ALL || q.n_calls == 0 || q.n_calls > 8 || q.interval <= 0) return E_UNSUPPORTED;
    const int col = q.calls[0].column;
    if (col < 0 || (uint32_t)col >= sh.n_columns || sh.columns[col].type != OG_TYPE_FLOAT) return E_UNSUPPORTED;
    const bool multi = q.n_calls > 1;
    for (uint32_t c = 0; c < q.n_calls; c++) {
        const int f = q.calls[c].func;
        if (q.calls[c].column != col) return E_UNSUPPORTED;
        if (f != OG_AGG_COUNT && f != OG_AGG_SUM && !(multi && (f == OG_AGG_MIN || f == OG_AGG_MAX))) return E_UNSUPPORTED; /* selectors that carry a time stay with the checker */
    }
    if (s_end > sh.n_series) s_end = sh.n_series;
    /* dense geometry exactly like scan_aggregate: file range intersected with the query range */
    int64_t fmin = INT64_MAX, fmax = INT64_MIN;
    for (uint32_t g = 0; g < sh.n_segments; g++) { fmin = std::min(fmin, sh.seg_tmin[g]); fmax = std::max(fmax, sh.seg_tmax[g]); }
    int64_t gmin = std::max(q.tmin, fmin), gmax = std::min(q.tmax, fmax);
    if (gmin > gmax) gmin = gmax = q.tmin;
    int64_t s0, e0, s1, e1;
    WindowOpt w; w.interval = q.interval; w.offset = q.offset; w.start_time = q.tmin; w.end_time = q.tmax;
    window(w, gmin, &s0, &e0);
    window(w, gmax + 1, &s1, &e1);
    const int64_t interval = e0 - s0;
    const uint32_t nb = (uint32_t)((uint64_t)(e1 - s0) / (uint64_t)interval);
    if (threads < 1) threads = 1;
    std::vector<std::vector<Cell>> part((size_t)threads, std::vector<Cell>(nb));
    std::vector<uint64_t> rows_t((size_t)threads, 0), segs_t((size_t)threads, 0), bytes_t((size_t)threads, 0);
    std::vector<int> rc_t((size_t)threads, E_OK);
    const uint8_t *data = sh.data;
    auto worker = [&](int tid) {
        std::vector<double> vals(70000);
        std::vector<Cell> &dense = part[(size_t)tid];
        for (uint32_t sr = s_begin + (uint32_t)tid; sr < s_end; sr += (uint32_t)threads) {
            Cell open; uint32_t open_b = 0xffffffffu; /* the series' window still open across records (prevBuf) */
            auto flush = [&]() { /* AggTagSetCursor: one series partial into the interval record (update*Sum/Count/Column{Min,Max}Impl) */
                if (open_b == 0xffffffffu || !open.ok) return;
                Cell &d = dense[open_b];
                d.sum = open.sum + d.sum; d.cnt = open.cnt + d.cnt;
                if (!(d.ok && d.mn <= open.mn)) d.mn = open.mn;
                if (!(d.ok && d.mx >= open.mx)) d.mx = open.mx;
                d.ok = true;
            };
            for (uint32_t g = sh.series_seg_begin[sr]; g < sh.series_seg_begin[sr + 1]; g++) {
                if (sh.seg_tmax[g] < q.tmin || sh.seg_tmin[g] > q.tmax) continue;
                const uint8_t *tp = data + sh.time_page_off[g]; const uint32_t tl = sh.time_page_len[g];
                const uint8_t *vp = data + sh.columns[col].page_off[g]; const uint32_t vl = sh.columns[col].page_len[g];
                if (tl < 16 || tp[0] != 32 || (tp[5] >> 4) != 1 || vl < 6 || vp[0] != 31) { rc_t[(size_t)tid] = E_UNSUPPORTED; return; }
                const uint32_t rows = be32(tp + 1);
                const int64_t t0 = (int64_t)be64(tp + 6);
                uint64_t dt = 0; { unsigned s = 0; for (uint32_t i = 14; i < tl; i++) { uint8_t c = tp[i]; dt |= (uint64_t)(c & 0x7f) << s; if (c < 0x80) break; s += 7; } }
                if (rows > vals.size() || be32(vp + 1) != rows || dt == 0) { rc_t[(size_t)tid] = E_UNSUPPORTED; return; }
                int n;
                const int tag = vp[5] >> 4;
                if (tag == 3) n = gorilla_batch(vp + 6, vl - 6, vals.data(), (int)rows);
                else if (tag == 0 && vl == 6 + 8ull * rows) { memcpy(vals.data(), vp + 6, 8ull * rows); n = (int)rows; }
                else { rc_t[(size_t)tid] = E_UNSUPPORTED; return; }
                if (n != (int)rows) { rc_t[(size_t)tid] = E_CORRUPT; return; }
                rows_t[(size_t)tid] += rows; segs_t[(size_t)tid]++; bytes_t[(size_t)tid] += tl + vl;
                /* FilterByTime */
                uint32_t r_lo = 0, r_hi = rows;
                if (t0 < q.tmin) r_lo = (uint32_t)std::min<uint64_t>(rows, ((uint64_t)(q.tmin - t0) + dt - 1) / dt);
                if (t0 + (int64_t)((rows - 1) * dt) > q.tmax) r_hi = q.tmax < t0 ? 0 : (uint32_t)((uint64_t)(q.tmax - t0) / dt) + 1;
                uint32_t r = r_lo;
                while (r < r_hi) {
                    const int64_t t = t0 + (int64_t)(r * dt);
                    const uint32_t b = (uint32_t)((uint64_t)(t - s0) / (uint64_t)interval);
                    const int64_t wend = s0 + (int64_t)(b + 1) * interval;
                    uint32_t re = (uint32_t)std::min<uint64_t>(r_hi, ((uint64_t)(wend - t0) + dt - 1) / dt); /* first row of the next window */
                    /* the window's rows of this record, left to right from 0.0 (floatSumReduce); first value seeds min/max */
                    double s = 0.0, mn = vals[r], mx = vals[r];
                    for (uint32_t i = r; i < re; i++) { const double v = vals[i]; s = s + v; if (mn > v) mn = v; if (mx < v) mx = v; }
                    const int64_t c = re - r;
                    if (b == open_b) { /* prevBuf (+) currBuf */
                        open.sum = open.sum + s; open.cnt += c;
                        if (mn < open.mn) open.mn = mn;
                        if (mx > open.mx) open.mx = mx;
                    } else {
                        flush();
                        open_b = b; open.sum = s; open.cnt = c; open.mn = mn; open.mx = mx; open.ok = true;
                    }
                    r = re;
                }
            }
            flush();
        }
    };
    if (threads == 1) worker(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++) th.emplace_back(worker, t);
        for (auto &t : th) t.join();
    }
    for (int t = 0; t < threads; t++) if (rc_t[(size_t)t] != E_OK) return rc_t[(size_t)t];
    out.n_groups = 1; out.n_buckets = nb; out.start = s0; out.interval = interval;
    out.values.assign(q.n_calls, std::vector<uint64_t>(nb, 0)); out.valid.assign(q.n_calls, std::vector<uint8_t>(nb, 0));
    out.times.assign(q.n_calls, std::vector<int64_t>(nb, 0));
    out.rows_decoded = out.segments = out.page_bytes = 0;
    for (int t = 0; t < threads; t++) { out.rows_decoded += rows_t[(size_t)t]; out.segments += segs_t[(size_t)t]; out.page_bytes += bytes_t[(size_t)t]; }
    for (uint32_t b = 0; b < nb; b++) {
        Cell d = part[0][b];
        for (int t = 1; t < threads; t++) { /* the workers' partial records, in worker order */
            const Cell &p = part[(size_t)t][b];
            if (!p.ok) continue;
            d.sum = p.sum + d.sum; d.cnt += p.cnt;
            if (!(d.ok && d.mn <= p.mn)) d.mn = p.mn;
            if (!(d.ok && d.mx >= p.mx)) d.mx = p.mx;
            d.ok = true;
        }
        for (uint32_t c = 0; c < q.n_calls; c++) {
            out.valid[c][b] = d.ok; out.times[c][b] = s0 + (int64_t)b * interval;
            if (!d.ok) continue;
            uint64_t u = 0;
            switch (q.calls[c].func) {
            case OG_AGG_COUNT: u = (uint64_t)d.cnt; break;
            case OG_AGG_SUM: memcpy(&u, &d.sum, 8); break;
            case OG_AGG_MIN: memcpy(&u, &d.mn, 8); break;
            default: memcpy(&u, &d.mx, 8); break;
            }
            out.values[c][b] = u;
        }
    }
    return E_OK;
}

} // namespace ogo
