/*
 * capi.cpp — CPU ORACLE (test infrastructure, see og_oracle.h): flat C entry points so tests/ and bench.py's
 * cpu_baseline leg can drive the oracle through ctypes.  Not part of the product ABI (that is include/ogpu.h).
 */
#include <cstring>

#include "og_oracle.h"

using namespace ogo;

#define OGO_API extern "C" __attribute__((visibility("default")))

static long copy_out(const Bytes &b, uint8_t *out, size_t cap) {
    if (b.size() > cap) return E_INVAL;
    if (!b.empty()) memcpy(out, b.data(), b.size());
    return (long)b.size();
}

OGO_API long ogo_float_encode(const double *v, size_t n, uint8_t *out, size_t cap) {
    Bytes b; int rc = float_block_encode(v, n, b); if (rc != E_OK) return rc; return copy_out(b, out, cap);
}
OGO_API long ogo_float_decode(const uint8_t *in, size_t len, double *out, size_t cap) {
    std::vector<double> v; int rc = float_block_decode(in, len, v); if (rc != E_OK) return rc;
    if (v.size() > cap) return E_INVAL; if (!v.empty()) memcpy(out, v.data(), v.size() * 8); return (long)v.size();
}
OGO_API long ogo_gorilla_encode(const double *v, size_t n, uint8_t *out, size_t cap) {
    Bytes b; int rc = gorilla_encode(v, n, b); if (rc != E_OK) return rc; return copy_out(b, out, cap);
}
OGO_API long ogo_int_encode(const int64_t *v, size_t n, uint8_t *out, size_t cap) {
    Bytes b; int rc = int_block_encode(v, n, b); if (rc != E_OK) return rc; return copy_out(b, out, cap);
}
OGO_API long ogo_int_decode(const uint8_t *in, size_t len, int64_t *out, size_t cap) {
    std::vector<int64_t> v; int rc = int_block_decode(in, len, v); if (rc != E_OK) return rc;
    if (v.size() > cap) return E_INVAL; if (!v.empty()) memcpy(out, v.data(), v.size() * 8); return (long)v.size();
}
OGO_API long ogo_time_encode(const int64_t *v, size_t n, uint8_t *out, size_t cap) {
    Bytes b; int rc = time_block_encode(v, n, b); if (rc != E_OK) return rc; return copy_out(b, out, cap);
}
OGO_API long ogo_time_decode(const uint8_t *in, size_t len, int64_t *out, size_t cap) {
    std::vector<int64_t> v; int rc = time_block_decode(in, len, v); if (rc != E_OK) return rc;
    if (v.size() > cap) return E_INVAL; if (!v.empty()) memcpy(out, v.data(), v.size() * 8); return (long)v.size();
}
OGO_API long ogo_bool_encode(const uint8_t *v, size_t n, uint8_t *out, size_t cap) {
    Bytes b; int rc = bool_block_encode(v, n, b); if (rc != E_OK) return rc; return copy_out(b, out, cap);
}
OGO_API long ogo_bool_decode(const uint8_t *in, size_t len, uint8_t *out, size_t cap) {
    std::vector<uint8_t> v; int rc = bool_block_decode(in, len, v); if (rc != E_OK) return rc;
    if (v.size() > cap) return E_INVAL; if (!v.empty()) memcpy(out, v.data(), v.size()); return (long)v.size();
}
OGO_API long ogo_snappy_roundtrip(const uint8_t *in, size_t len, uint8_t *out, size_t cap) {
    Bytes c, d; snappy_encode(in, len, c); int rc = snappy_decode(c.data(), c.size(), d); if (rc != E_OK) return rc; return copy_out(d, out, cap);
}
OGO_API long ogo_s8b_encode(uint64_t *v, size_t n) { return s8b_encode_all(v, n); }

/* rows: per-row cells (8 B, bool 1 B) with a slot for nulls; valid: per-row 0/1 or NULL */
OGO_API long ogo_field_page_encode(int type, const void *cells, const uint8_t *valid, size_t rows, uint8_t *out, size_t cap) {
    ColVal cv;
    for (size_t i = 0; i < rows; i++) {
        bool ok = !valid || valid[i];
        if (!ok) { cv.append_null(type, false); continue; }
        if (type == OG_TYPE_BOOL) cv.append_boolean(((const uint8_t *)cells)[i] != 0);
        else { int64_t x; memcpy(&x, (const uint8_t *)cells + 8 * i, 8); cv.append_integer(x); }
    }
    Bytes b; int rc = encode_field_page(cv, type, b); if (rc != E_OK) return rc; return copy_out(b, out, cap);
}
/* values_out: dense non-null values; valid_out: per-row 0/1 */
OGO_API long ogo_field_page_decode(int type, const uint8_t *page, size_t len, void *values_out, size_t cap_values,
                                   uint8_t *valid_out, size_t cap_rows, int *nil_count) {
    ColVal cv; int rc = decode_field_page(page, len, type, cv); if (rc != E_OK) return rc;
    size_t nv = cv.n_values(type);
    if (nv > cap_values || (size_t)cv.len > cap_rows) return E_INVAL;
    if (!cv.val.empty()) memcpy(values_out, cv.val.data(), cv.val.size());
    for (int i = 0; i < cv.len; i++) valid_out[i] = cv.is_nil(i) ? 0 : 1;
    if (nil_count) *nil_count = cv.nil_count;
    return cv.len;
}
OGO_API long ogo_time_page_encode(const int64_t *t, size_t n, uint8_t *out, size_t cap) {
    Bytes b; int rc = encode_time_page(t, n, b); if (rc != E_OK) return rc; return copy_out(b, out, cap);
}
OGO_API long ogo_time_page_decode(const uint8_t *page, size_t len, int64_t *out, size_t cap) {
    ColVal cv; int rc = decode_time_page(page, len, cv); if (rc != E_OK) return rc;
    if ((size_t)cv.len > cap) return E_INVAL; memcpy(out, cv.val.data(), cv.val.size()); return cv.len;
}

OGO_API void ogo_window(int64_t interval, int64_t offset, int64_t tmin, int64_t tmax, int64_t t, int64_t *s, int64_t *e) {
    WindowOpt w; w.interval = interval; w.offset = offset; w.start_time = tmin; w.end_time = tmax; window(w, t, s, e);
}

/* ---- synthetic shard ---- */
OGO_API void *ogo_synth_build_mt(const og_synth_desc *d, int threads, int *status) {
    HostShard *h = new HostShard; int rc = build_synth_shard(*d, *h, threads); if (status) *status = rc;
    if (rc != E_OK) { delete h; return nullptr; }
    return h;
}
OGO_API void *ogo_synth_build(const og_synth_desc *d, int *status) {
    HostShard *h = new HostShard; int rc = build_synth_shard(*d, *h, 1); if (status) *status = rc;
    if (rc != E_OK) { delete h; return nullptr; }
    return h;
}
struct ShardHolder { og_shard_desc d; };
OGO_API void ogo_shard_desc(void *h, og_shard_desc *out) { *out = ((HostShard *)h)->desc(); }
OGO_API void ogo_shard_free(void *h) { delete (HostShard *)h; }

/* ---- scan ---- */
OGO_API void *ogo_scan(const og_shard_desc *sh, const og_query_desc *q, int threads, uint32_t s0, uint32_t s1, int *status) {
    ScanResult *r = new ScanResult; int rc = scan_aggregate(*sh, *q, threads, s0, s1, *r); if (status) *status = rc;
    if (rc != E_OK) { delete r; return nullptr; }
    return r;
}
OGO_API void *ogo_fast_scan(const og_shard_desc *sh, const og_query_desc *q, int threads, uint32_t s0, uint32_t s1, int *status) {
    ScanResult *r = new ScanResult; int rc = fast_scan_aggregate(*sh, *q, threads, s0, s1, *r); if (status) *status = rc;
    if (rc != E_OK) { delete r; return nullptr; }
    return r;
}
OGO_API void ogo_scan_dims(void *h, uint32_t *n_groups, uint32_t *n_buckets, int64_t *start, int64_t *interval,
                           uint64_t *rows, uint64_t *segs, uint64_t *bytes) {
    ScanResult *r = (ScanResult *)h;
    *n_groups = r->n_groups; *n_buckets = r->n_buckets; *start = r->start; *interval = r->interval;
    *rows = r->rows_decoded; *segs = r->segments; *bytes = r->page_bytes;
}
OGO_API void ogo_scan_col(void *h, uint32_t k, const uint64_t **values, const uint8_t **valid, const int64_t **times) {
    ScanResult *r = (ScanResult *)h;
    *values = r->values[k].data(); *valid = r->valid[k].data(); *times = r->times[k].data();
}
OGO_API void ogo_scan_free(void *h) { delete (ScanResult *)h; }
