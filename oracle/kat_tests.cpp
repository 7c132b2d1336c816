/*
 * kat_tests.cpp — CPU ORACLE (test infrastructure): the reference's own known-answer tests for this path,
 * restated against the oracle so that the oracle is PINNED before anything is compared with it.
 *
 *   engine/iterators_test.go:748-2043  TestAggregateCursor_{Multi,Single}_{Count,Sum,Min,Max,First,Last}
 *                                      (the aux-column sub-cases "select min(int),float" need auxProcessors, which the
 *                                      GPU path does not push down; they are not restated)
 *   engine/iterators_test.go:2956-3050 TestIntervalRecordBuild{Asc}, TestTransIntervalRecord2Rec
 *   lib/encoding/timestamp_test.go:25-60 literal 32-timestamp vector (round trip + codec choice)
 *   lib/encoding/encoding_test.go / lib/compress/float_test.go round-trip shapes
 *
 * Exit code 0 = all pass.  Run by tests/test_oracle_kat.py.
 */
#include <cmath>
#include <cstdio>
#include <cstring>
#include <deque>
#include <functional>

#include "og_oracle.h"

using namespace ogo;

static int g_fail = 0, g_checks = 0;
#define CHECK(cond, ...) do { g_checks++; if (!(cond)) { g_fail++; printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } } while (0)

/* ---------- helpers mirroring the Go test harness (iterators_test.go:614-746) ---------- */
struct VecCursor { std::deque<const Record *> buf; };
static const Record *vec_next(void *ctx) {
    VecCursor *c = (VecCursor *)ctx;
    if (c->buf.empty()) return nullptr;
    const Record *r = c->buf.front(); c->buf.pop_front(); return r;
}

struct Col { int type; std::vector<double> f; std::vector<int64_t> i; std::vector<int> b; };
static Col I(std::vector<int64_t> v) { Col c; c.type = OG_TYPE_INT; c.i = v; return c; }
static Col F(std::vector<double> v) { Col c; c.type = OG_TYPE_FLOAT; c.f = v; return c; }
static Col B(std::vector<int> v) { Col c; c.type = OG_TYPE_BOOL; c.b = v; return c; }

static Record mk(const std::vector<Field> &schema, const std::vector<Col> &cols, const std::vector<int64_t> &times,
                 const std::vector<std::vector<int64_t>> &meta = {}) {
    Record r(schema);
    for (size_t k = 0; k < cols.size(); k++) {
        for (double x : cols[k].f) r.cols[k].append_float(x);
        for (int64_t x : cols[k].i) r.cols[k].append_integer(x);
        for (int x : cols[k].b) r.cols[k].append_boolean(x != 0);
    }
    for (int64_t t : times) r.append_time(t);
    for (size_t k = 0; k < meta.size(); k++) r.meta_times[k] = meta[k];
    return r;
}

static bool rec_equal(const Record &a, const Record &b, bool check_meta) { /* isRecEqual :504-536 */
    if (a.cols.size() != b.cols.size()) return false;
    for (size_t i = 0; i < a.cols.size(); i++) {
        const ColVal &x = a.cols[i], &y = b.cols[i];
        if (x.len != y.len || x.nil_count != y.nil_count || x.val != y.val) return false;
        for (int j = 0; j < x.len; j++) if (x.is_nil(j) != y.is_nil(j)) return false;
        if (check_meta && a.meta_times[i] != b.meta_times[i]) return false;
    }
    return true;
}

static void dump(const Record &r) {
    for (size_t i = 0; i < r.cols.size(); i++) {
        printf("   col %zu (%s, type %d) len=%d nil=%d:", i, r.schema[i].name.c_str(), r.schema[i].type, r.cols[i].len, r.cols[i].nil_count);
        size_t n = r.cols[i].n_values(r.schema[i].type);
        for (size_t k = 0; k < n; k++) {
            if (r.schema[i].type == OG_TYPE_FLOAT) printf(" %.17g", r.cols[i].floats()[k]);
            else if (r.schema[i].type == OG_TYPE_BOOL) printf(" %d", r.cols[i].val[k]);
            else printf(" %lld", (long long)r.cols[i].integers()[k]);
        }
        printf("\n");
    }
}

static void run_case(const char *name, const std::vector<Field> &in_schema, const std::vector<Field> &out_schema,
                     const std::vector<Record> &src, const std::vector<Record> &dst, const std::vector<ExprOpt> &exprs,
                     int64_t interval, int chunk, bool check_meta = false) {
    WindowOpt w; w.interval = interval; w.start_time = INT64_MIN + 2; w.end_time = INT64_MAX - 1;
    AggCursor *c = agg_cursor_new(in_schema, out_schema, exprs, w, chunk);
    CHECK(c != nullptr, "%s: cursor build", name);
    if (!c) return;
    VecCursor vc;
    for (const Record &r : src) vc.buf.push_back(&r);
    agg_cursor_set_input(c, vec_next, &vc);
    std::vector<Record> outs;
    while (const Record *o = agg_cursor_next(c)) outs.push_back(*o);
    CHECK(outs.size() == dst.size(), "%s: record count %zu != expected %zu", name, outs.size(), dst.size());
    for (size_t i = 0; i < outs.size() && i < dst.size(); i++) {
        bool ok = rec_equal(outs[i], dst[i], check_meta);
        CHECK(ok, "%s: record %zu differs", name, i);
        if (!ok) { printf("  got:\n"); dump(outs[i]); printf("  want:\n"); dump(dst[i]); }
    }
    agg_cursor_free(c);
}

static const std::vector<Field> S_IFT = {{"int", OG_TYPE_INT}, {"float", OG_TYPE_FLOAT}, {"time", OG_TYPE_INT}};
static const std::vector<Field> S_IT = {{"int", OG_TYPE_INT}, {"time", OG_TYPE_INT}};

static std::vector<Record> src_if() {
    return {mk(S_IFT, {I({1, 2, 3}), F({1.1, 2.2, 3.3})}, {1, 2, 3}), mk(S_IFT, {I({4, 5, 6}), F({4.4, 5.5, 6.6})}, {4, 5, 6}),
            mk(S_IFT, {I({7, 8, 9}), F({7.7, 8.8, 9.9})}, {7, 8, 9})};
}
static std::vector<Record> src_i() {
    return {mk(S_IT, {I({1, 2, 3})}, {1, 2, 3}), mk(S_IT, {I({4, 5, 6})}, {4, 5, 6}), mk(S_IT, {I({7, 8, 9})}, {7, 8, 9})};
}

static void test_count() {
    /* TestAggregateCursor_Multi_Count :748-876 (int/float/boolean columns; string counting is not on the GPU path) */
    std::vector<Field> in = {{"int", OG_TYPE_INT}, {"float", OG_TYPE_FLOAT}, {"boolean", OG_TYPE_BOOL}, {"time", OG_TYPE_INT}};
    std::vector<Field> out = {{"int", OG_TYPE_INT}, {"float", OG_TYPE_INT}, {"boolean", OG_TYPE_INT}, {"time", OG_TYPE_INT}};
    auto src = [&]() {
        return std::vector<Record>{mk(in, {I({1, 2, 3}), F({1.1, 2.2, 3.3}), B({1, 1, 1})}, {1, 2, 3}),
                                   mk(in, {I({4, 5, 6}), F({4.4, 5.5, 6.6}), B({0, 0, 0})}, {4, 5, 6}),
                                   mk(in, {I({7, 8, 9}), F({7.7, 8.8, 9.9}), B({1, 1, 1})}, {7, 8, 9})};
    };
    std::vector<ExprOpt> ex = {{OG_AGG_COUNT, "int", "int"}, {OG_AGG_COUNT, "float", "float"}, {OG_AGG_COUNT, "boolean", "boolean"}};
    run_case("multi_count/1", in, out, src(), {mk(out, {I({9}), I({9}), I({9})}, {7})}, ex, 0, 3);
    { /* :812-826 one empty input record */
        auto s = src(); s[1] = Record(in);
        run_case("multi_count/empty-record", in, out, s, {mk(out, {I({6}), I({6}), I({6})}, {7})}, ex, 0, 3);
    }
    run_case("multi_count/time(2)", in, out, src(),
             {mk(out, {I({1, 2, 2}), I({1, 2, 2}), I({1, 2, 2})}, {1, 2, 4}), mk(out, {I({2, 2}), I({2, 2}), I({2, 2})}, {7, 8})}, ex, 2, 3);
    /* TestAggregateCursor_Single_Count :878-949 */
    std::vector<ExprOpt> e1 = {{OG_AGG_COUNT, "int", "int"}};
    run_case("single_count/1", S_IT, S_IT, src_i(), {mk(S_IT, {I({9})}, {1})}, e1, 0, 3);
    run_case("single_count/time(2)", S_IT, S_IT, src_i(), {mk(S_IT, {I({1, 2, 2})}, {1, 2, 4}), mk(S_IT, {I({2, 2})}, {6, 8})}, e1, 2, 3);
}

static void test_sum() { /* :951-1107 */
    std::vector<ExprOpt> ex = {{OG_AGG_SUM, "int", "int"}, {OG_AGG_SUM, "float", "float"}};
    run_case("multi_sum/1", S_IFT, S_IFT, src_if(), {mk(S_IFT, {I({45}), F({49.5})}, {7})}, ex, 0, 3);
    run_case("multi_sum/time(2)", S_IFT, S_IFT, src_if(),
             {mk(S_IFT, {I({1, 5, 9}), F({1.1, 5.5, 9.9})}, {1, 2, 4}), mk(S_IFT, {I({13, 17}), F({14.3, 18.700000000000003})}, {7, 8})}, ex, 2, 3);
    std::vector<ExprOpt> e1 = {{OG_AGG_SUM, "int", "int"}};
    run_case("single_sum/1", S_IT, S_IT, src_i(), {mk(S_IT, {I({45})}, {1})}, e1, 0, 3);
    run_case("single_sum/time(2)", S_IT, S_IT, src_i(), {mk(S_IT, {I({1, 5, 9})}, {1, 2, 4}), mk(S_IT, {I({13, 17})}, {6, 8})}, e1, 2, 3);
}

static void test_min_max() { /* :1109-1582 */
    std::vector<ExprOpt> mn = {{OG_AGG_MIN, "int", "int"}, {OG_AGG_MIN, "float", "float"}};
    run_case("multi_min/1", S_IFT, S_IFT, src_if(), {mk(S_IFT, {I({1}), F({1.1})}, {7})}, mn, 0, 3);
    run_case("multi_min/time(2)", S_IFT, S_IFT, src_if(),
             {mk(S_IFT, {I({1, 2, 4}), F({1.1, 2.2, 4.4})}, {1, 2, 4}), mk(S_IFT, {I({6, 8}), F({6.6, 8.8})}, {7, 8})}, mn, 2, 3);
    std::vector<ExprOpt> mn1 = {{OG_AGG_MIN, "int", "int"}};
    run_case("single_min/1", S_IT, S_IT, src_i(), {mk(S_IT, {I({1})}, {1})}, mn1, 0, 3);
    run_case("single_min/time(2)", S_IT, S_IT, src_i(), {mk(S_IT, {I({1, 2, 4})}, {1, 2, 4}), mk(S_IT, {I({6, 8})}, {6, 8})}, mn1, 2, 3);
    std::vector<ExprOpt> mx = {{OG_AGG_MAX, "int", "int"}, {OG_AGG_MAX, "float", "float"}};
    run_case("multi_max/1", S_IFT, S_IFT, src_if(), {mk(S_IFT, {I({9}), F({9.9})}, {7})}, mx, 0, 3);
    run_case("multi_max/time(2)", S_IFT, S_IFT, src_if(),
             {mk(S_IFT, {I({1, 3, 5}), F({1.1, 3.3, 5.5})}, {1, 2, 4}), mk(S_IFT, {I({7, 9}), F({7.7, 9.9})}, {7, 8})}, mx, 2, 3);
    std::vector<ExprOpt> mx1 = {{OG_AGG_MAX, "int", "int"}};
    run_case("single_max/1", S_IT, S_IT, src_i(), {mk(S_IT, {I({9})}, {9})}, mx1, 0, 3);
    run_case("single_max/time(2)", S_IT, S_IT, src_i(), {mk(S_IT, {I({1, 3, 5})}, {1, 3, 5}), mk(S_IT, {I({7, 9})}, {7, 9})}, mx1, 2, 3);
}

static void test_first_last() { /* :1584-2042 */
    std::vector<Field> sc = {{"int", OG_TYPE_INT}, {"float", OG_TYPE_FLOAT}, {"boolean", OG_TYPE_BOOL}, {"time", OG_TYPE_INT}};
    auto src = [&]() {
        return std::vector<Record>{mk(sc, {I({1, 2, 3}), F({1.1, 2.2, 3.3}), B({1, 1, 1})}, {1, 2, 3}),
                                   mk(sc, {I({4, 5, 6}), F({4.4, 5.5, 6.6}), B({0, 0, 0})}, {4, 5, 6}),
                                   mk(sc, {I({7, 8, 9}), F({7.7, 8.8, 9.9}), B({1, 1, 1})}, {7, 8, 9})};
    };
    std::vector<ExprOpt> fi = {{OG_AGG_FIRST, "int", "int"}, {OG_AGG_FIRST, "float", "float"}, {OG_AGG_FIRST, "boolean", "boolean"}};
    run_case("multi_first/1", sc, sc, src(), {mk(sc, {I({1}), F({1.1}), B({1})}, {7}, {{1}, {1}, {1}})}, fi, 0, 3, true);
    run_case("multi_first/time(2)", sc, sc, src(),
             {mk(sc, {I({1, 2, 4}), F({1.1, 2.2, 4.4}), B({1, 1, 0})}, {1, 2, 4}, {{1, 2, 4}, {1, 2, 4}, {1, 2, 4}}),
              mk(sc, {I({6, 8}), F({6.6, 8.8}), B({0, 1})}, {7, 8}, {{6, 8}, {6, 8}, {6, 8}})}, fi, 2, 3, true);
    std::vector<ExprOpt> f1 = {{OG_AGG_FIRST, "int", "int"}};
    run_case("single_first/1", S_IT, S_IT, src_i(), {mk(S_IT, {I({1})}, {1})}, f1, 0, 3);
    run_case("single_first/time(2)", S_IT, S_IT, src_i(), {mk(S_IT, {I({1, 2, 4})}, {1, 2, 4}), mk(S_IT, {I({6, 8})}, {6, 8})}, f1, 2, 3);
    std::vector<ExprOpt> la = {{OG_AGG_LAST, "int", "int"}, {OG_AGG_LAST, "float", "float"}, {OG_AGG_LAST, "boolean", "boolean"}};
    run_case("multi_last/1", sc, sc, src(), {mk(sc, {I({9}), F({9.9}), B({1})}, {7}, {{9}, {9}, {9}})}, la, 0, 3, true);
    run_case("multi_last/time(2)", sc, sc, src(),
             {mk(sc, {I({1, 3, 5}), F({1.1, 3.3, 5.5}), B({1, 1, 0})}, {1, 2, 4}, {{1, 3, 5}, {1, 3, 5}, {1, 3, 5}}),
              mk(sc, {I({7, 9}), F({7.7, 9.9}), B({1, 1})}, {7, 8}, {{7, 9}, {7, 9}, {7, 9}})}, la, 2, 3, true);
    std::vector<ExprOpt> l1 = {{OG_AGG_LAST, "int", "int"}};
    run_case("single_last/1", S_IT, S_IT, src_i(), {mk(S_IT, {I({9})}, {9})}, l1, 0, 3);
    run_case("single_last/time(2)", S_IT, S_IT, src_i(), {mk(S_IT, {I({1, 3, 5})}, {1, 3, 5}), mk(S_IT, {I({7, 9})}, {7, 9})}, l1, 2, 3);
}

static void test_interval_record() { /* TestIntervalRecordBuildAsc :2956, TestTransIntervalRecord2Rec :3006 */
    const int64_t S = 1000000000;
    IntervalRecord ir;
    ir.schema = {{"int", OG_TYPE_INT}, {"float", OG_TYPE_FLOAT}, {"boolean", OG_TYPE_BOOL}, {"time", OG_TYPE_INT}};
    ir.build(0, S, S, true);
    CHECK(ir.n_rows == 1 && ir.times[0] == 0, "BuildEmptyIntervalRec(0,1s,1s): rows=%u first=%lld", ir.n_rows, (long long)ir.times[0]);
    ir.build(0, 6 * S, S, true);
    CHECK(ir.n_rows == 6 && ir.times[5] == 5 * S, "BuildEmptyIntervalRec(0,6s,1s)");
    /* single-call first: UpdateIntervalRecRow semantics through update_from */
    ir.exprs = {{OG_AGG_FIRST, "int", "int"}};
    ir.schema = {{"int", OG_TYPE_INT}, {"time", OG_TYPE_INT}};
    ir.build(0, 6 * S, S, true);
    Record r1 = mk(S_IT, {I({0, 1, 2})}, {S, 3 * S, 5 * S});
    ir.update_from(r1);
    CHECK(ir.valid[0][1] && ir.values[0][1] == 0 && ir.valid[0][3] && ir.values[0][3] == 1 && !ir.valid[0][0], "interval first update");
    Record r2 = mk(S_IT, {I({7})}, {S}); /* same time -> larger value wins (reccord_functions.go:63-74) */
    ir.update_from(r2);
    CHECK(ir.values[0][1] == 7, "first tie on time: larger value wins, got %lld", (long long)ir.values[0][1]);
    /* max: tie on value -> earlier time kept (:482-494) */
    IntervalRecord mx; mx.exprs = {{OG_AGG_MAX, "int", "int"}}; mx.schema = {{"int", OG_TYPE_INT}, {"time", OG_TYPE_INT}};
    mx.build(0, 10, 10, true);
    mx.update_from(mk(S_IT, {I({5})}, {7}));
    mx.update_from(mk(S_IT, {I({5})}, {3}));
    mx.update_from(mk(S_IT, {I({5})}, {8}));
    CHECK(mx.values[0][0] == 5 && mx.times[0] == 3, "max tie -> earlier time, got t=%lld", (long long)mx.times[0]);
    /* sum across series in arrival order */
    IntervalRecord sm; sm.exprs = {{OG_AGG_SUM, "float", "float"}}; sm.schema = {{"float", OG_TYPE_FLOAT}, {"time", OG_TYPE_INT}};
    sm.build(0, 10, 10, true);
    std::vector<Field> sf = {{"float", OG_TYPE_FLOAT}, {"time", OG_TYPE_INT}};
    sm.update_from(mk(sf, {F({0.1})}, {1})); sm.update_from(mk(sf, {F({0.2})}, {2})); sm.update_from(mk(sf, {F({0.3})}, {3}));
    double got; memcpy(&got, &sm.values[0][0], 8);
    CHECK(got == (0.1 + 0.2) + 0.3 && sm.times[0] == 0, "sum order / time stays window start");
}

static void test_window() { /* select.go:579 semantics, incl. negative times and offsets */
    WindowOpt w; w.interval = 60; w.offset = 0;
    int64_t s, e;
    window(w, 125, &s, &e); CHECK(s == 120 && e == 180, "window(125)");
    window(w, -1, &s, &e); CHECK(s == -60 && e == 0, "window(-1) = [%lld,%lld)", (long long)s, (long long)e);
    window(w, 120, &s, &e); CHECK(s == 120 && e == 180, "window(120)");
    w.offset = 7;
    window(w, 125, &s, &e); CHECK(s == 67 && e == 127, "window(125, offset 7) = [%lld,%lld)", (long long)s, (long long)e);
    window(w, 127, &s, &e); CHECK(s == 127 && e == 187, "window(127, offset 7)");
    w.interval = 0; w.start_time = 5; w.end_time = 9;
    window(w, 7, &s, &e); CHECK(s == 5 && e == 10, "no interval");
}

/* ---------- codec round trips (lib/encoding/encoding_test.go, timestamp_test.go, lib/compress/float_test.go) ---------- */
static uint64_t rng_state = 0x1234567;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

static void rt_float(const char *name, const std::vector<double> &v, int want_tag) {
    Bytes b = {0xAA, 0xBB}; /* prefix preservation: encode appends after existing bytes (encoding_test.go) */
    int rc = float_block_encode(v.data(), v.size(), b);
    CHECK(rc == E_OK, "%s: encode rc=%d", name, rc);
    CHECK(b[0] == 0xAA && b[1] == 0xBB, "%s: prefix", name);
    if (want_tag >= 0 && !v.empty()) CHECK((b[2] >> 4) == want_tag, "%s: tag %d, want %d", name, b[2] >> 4, want_tag);
    std::vector<double> o;
    rc = float_block_decode(b.data() + 2, b.size() - 2, o);
    CHECK(rc == E_OK && o.size() == v.size(), "%s: decode rc=%d n=%zu/%zu", name, rc, o.size(), v.size());
    if (o.size() == v.size()) CHECK(memcmp(o.data(), v.data(), v.size() * 8) == 0, "%s: values differ", name);
}

static void test_codecs() {
    { /* timestamp_test.go:28-37 literal vector */
        std::vector<int64_t> t = {1675065600000000000, 1675065610000000000, 1675065620000000000, 1675065630000000000};
        Bytes b; CHECK(time_block_encode(t.data(), t.size(), b) == E_OK && (b[0] >> 4) == 1, "time const-delta tag");
        CHECK(b.size() == 1 + 8 + 5 + 1, "const-delta size %zu", b.size());
        std::vector<int64_t> o; CHECK(time_block_decode(b.data(), b.size(), o) == E_OK && o == t, "time const rt");
    }
    { /* irregular cadence, common scale 1e6 -> simple8b with scale */
        std::vector<int64_t> t; int64_t cur = 1700000000000000000;
        for (int i = 0; i < 1000; i++) { t.push_back(cur); cur += (int64_t)(1 + rnd() % 50) * 1000000; }
        Bytes b; CHECK(time_block_encode(t.data(), t.size(), b) == E_OK && (b[0] >> 4) == 2, "time s8b tag");
        CHECK(get_u64be(b.data() + 1) == 1000000, "time scale %llu", (unsigned long long)get_u64be(b.data() + 1));
        std::vector<int64_t> o; CHECK(time_block_decode(b.data(), b.size(), o) == E_OK && o == t, "time s8b rt");
    }
    { /* fewer than 3 -> raw */
        std::vector<int64_t> t = {5, 9};
        Bytes b; CHECK(time_block_encode(t.data(), 2, b) == E_OK && (b[0] >> 4) == 4, "time raw tag");
        std::vector<int64_t> o; CHECK(time_block_decode(b.data(), b.size(), o) == E_OK && o == t, "time raw rt");
    }
    { /* huge delta -> snappy or raw, value round trip only */
        std::vector<int64_t> t = {0, 1, (int64_t)1 << 61, ((int64_t)1 << 61) + 5, ((int64_t)1 << 62)};
        Bytes b; CHECK(time_block_encode(t.data(), t.size(), b) == E_OK, "time big enc");
        std::vector<int64_t> o; CHECK(time_block_decode(b.data(), b.size(), o) == E_OK && o == t, "time big rt (tag %d)", b[0] >> 4);
    }
    /* ints */
    auto rt_int = [&](const char *name, const std::vector<int64_t> &v, int want) {
        Bytes b; int rc = int_block_encode(v.data(), v.size(), b);
        CHECK(rc == E_OK, "%s enc rc=%d", name, rc); if (rc != E_OK) return;
        if (!v.empty()) CHECK((b[0] >> 4) == want, "%s tag %d want %d", name, b[0] >> 4, want);
        std::vector<int64_t> o; rc = int_block_decode(b.data(), b.size(), o);
        CHECK(rc == E_OK && o == v, "%s rt rc=%d", name, rc);
    };
    rt_int("int const", {10, 20, 30, 40, 50}, 1);
    rt_int("int const neg", {50, 40, 30, 20}, 1);
    rt_int("int raw<3", {7, -9}, 4);
    { std::vector<int64_t> v; int64_t c = 0; for (int i = 0; i < 1000; i++) { c += (int64_t)(rnd() % 2001) - 1000; v.push_back(c); } rt_int("int walk", v, 2); }
    { std::vector<int64_t> v; for (int i = 0; i < 500; i++) v.push_back(i % 2 ? INT64_C(1) << 40 : -(INT64_C(1) << 40)); rt_int("int big deltas", v, 2); }
    { std::vector<int64_t> v(300, 5); v[100] = 6; rt_int("int mostly same (s8b sel 0/1 quirk)", v, 2); }
    { /* all deltas == zigzag 1 -> selector 0/1 runs (canPack bits==0 quirk :455-462) */
        std::vector<int64_t> v; int64_t c = 0; for (int i = 0; i < 400; i++) { v.push_back(c); c -= 1; } rt_int("int delta -1", v, 1); }
    { std::vector<int64_t> v; int64_t c = 0; for (int i = 0; i < 400; i++) { v.push_back(c); c -= 1; } v[399] += 3; rt_int("int delta -1 then change", v, 2); }
    { std::vector<int64_t> v = {0, INT64_MAX, INT64_MIN, 5}; Bytes b; CHECK(int_block_encode(v.data(), v.size(), b) == E_UNSUPPORTED, "int zstd path is reported unsupported"); }
    /* bools */
    for (int n : {1, 7, 8, 9, 1000}) {
        std::vector<uint8_t> v; for (int i = 0; i < n; i++) v.push_back(rnd() & 1);
        Bytes b; CHECK(bool_block_encode(v.data(), v.size(), b) == E_OK && b.size() == (size_t)5 + (n + 7) / 8, "bool size n=%d", n);
        std::vector<uint8_t> o; CHECK(bool_block_decode(b.data(), b.size(), o) == E_OK && o == v, "bool rt n=%d", n);
    }
    /* floats: float_test.go shapes */
    rt_float("float <=4 raw", {1.5, 2.5, 3.5}, 0);
    rt_float("float same", std::vector<double>(100, 3.25), 4);
    rt_float("float same zero", std::vector<double>(100, 0.0), 4);
    { std::vector<double> v; for (int i = 0; i < 1000; i++) v.push_back((double)(i / 200)); rt_float("float rle", v, 5); }
    { std::vector<double> v; for (int i = 0; i < 1000; i++) v.push_back(i / 300 == 1 ? 0.0 : 7.0); rt_float("float rle zero-run", v, 5); }
    { std::vector<double> v; for (int i = 0; i < 1000; i++) v.push_back((double)(int64_t)(rnd() % 100000)); rt_float("float ints -> gorilla", v, 3); }
    { std::vector<double> v; for (int i = 0; i < 1000; i++) v.push_back((double)(rnd() % 100000) / 100.0); rt_float("float 2 decimals -> snappy", v, 2); }
    { std::vector<double> v; for (int i = 0; i < 1000; i++) v.push_back(100.0 + (double)(rnd() >> 11) / 9007199254740992.0); rt_float("float G-hi -> gorilla", v, 3); }
    { std::vector<double> v; for (int i = 0; i < 1000; i++) { uint64_t u = rnd(); double d; memcpy(&d, &u, 8); if (std::isnan(d) || std::isinf(d)) d = 1; v.push_back(d); } rt_float("float random bits -> raw (>90%)", v, 0); }
    { std::vector<double> v; for (int i = 0; i < 100; i++) v.push_back(i % 10 == 3 ? NAN : (double)i * 1.37); rt_float("float NaN -> snappy", v, 2); }
    { std::vector<double> v; for (int i = 0; i < 100; i++) v.push_back(i % 2 ? INFINITY : (double)i); rt_float("float +Inf", v, -1); }
    { std::vector<double> v; for (int i = 0; i < 100; i++) v.push_back(i == 50 ? INFINITY : i == 60 ? -INFINITY : (double)i);
      Bytes b; CHECK(float_block_encode(v.data(), v.size(), b) == E_NAN, "+Inf and -Inf: encoder error (batch_float.go:245)"); }
    { /* clz >= 32 wrap quirk (batch_float.go:88-91): tiny mantissa differences */
        std::vector<double> v; uint64_t base = 0x4059000000000000ull;
        for (int i = 0; i < 200; i++) { uint64_t u = base + (uint64_t)(rnd() % 1000); double d; memcpy(&d, &u, 8); v.push_back(d); }
        rt_float("float clz>=32", v, -1);
    }
    { /* gorilla empty + error shapes */
        Bytes g; CHECK(gorilla_encode(nullptr, 0, g) == E_OK && g.size() == 9, "gorilla empty = 9 bytes");
        std::vector<double> o; CHECK(gorilla_decode(g.data(), g.size(), o) == E_OK && o.empty(), "gorilla empty decode");
        uint8_t bad[] = {0x70, 1, 2, 3};
        CHECK(float_block_decode(bad, sizeof bad, o) == E_CORRUPT, "bad float tag (float_test.go:166-175)");
        uint8_t mlf[] = {0x60, 1, 2, 3};
        CHECK(float_block_decode(mlf, sizeof mlf, o) == E_UNSUPPORTED, "mlf unsupported");
    }
    { /* snappy codec self-consistency on compressible and incompressible input */
        for (int shape = 0; shape < 3; shape++) {
            Bytes in; for (int i = 0; i < 70000; i++) in.push_back(shape == 0 ? (uint8_t)(i % 7) : shape == 1 ? (uint8_t)rnd() : (uint8_t)((i / 100) & 0xff));
            Bytes c, d; snappy_encode(in.data(), in.size(), c);
            CHECK(snappy_decode(c.data(), c.size(), d) == E_OK && d == in, "snappy rt shape %d (%zu -> %zu)", shape, in.size(), c.size());
        }
    }
    { /* page framing: full / empty / one-row / partial-null (reader_test.go:308-341) */
        for (int type : {OG_TYPE_FLOAT, OG_TYPE_INT, OG_TYPE_BOOL}) {
            for (int shape = 0; shape < 4; shape++) {
                ColVal cv; int rows = shape == 2 ? 1 : 1000;
                for (int i = 0; i < rows; i++) {
                    bool nil = shape == 1 ? true : shape == 3 ? (rnd() % 20 == 0) : false;
                    if (nil) cv.append_null(type, false);
                    else if (type == OG_TYPE_FLOAT) cv.append_float(100.0 + (double)(rnd() >> 11) / 9007199254740992.0);
                    else if (type == OG_TYPE_INT) cv.append_integer((int64_t)(rnd() % 1000));
                    else cv.append_boolean(rnd() & 1);
                }
                Bytes p; int rc = encode_field_page(cv, type, p);
                CHECK(rc == E_OK, "page enc type %d shape %d", type, shape);
                uint8_t want = shape == 0 ? (type == OG_TYPE_FLOAT ? 31 : type == OG_TYPE_INT ? 32 : 33)
                             : shape == 1 ? (type == OG_TYPE_FLOAT ? 41 : type == OG_TYPE_INT ? 42 : 43)
                             : shape == 2 ? (type == OG_TYPE_FLOAT ? 17 : type == OG_TYPE_INT ? 18 : 19) : (uint8_t)type;
                CHECK(p[0] == want, "page header byte %d want %d (type %d shape %d)", p[0], want, type, shape);
                ColVal o; rc = decode_field_page(p.data(), p.size(), type, o);
                CHECK(rc == E_OK && o.len == cv.len && o.nil_count == cv.nil_count && o.val == cv.val, "page rt type %d shape %d rc=%d", type, shape, rc);
                for (int i = 0; i < cv.len && rc == E_OK; i++) if (o.is_nil(i) != cv.is_nil(i)) { CHECK(false, "page bitmap row %d", i); break; }
                ColVal wrong; if (shape == 3) CHECK(decode_field_page(p.data(), p.size(), type == OG_TYPE_INT ? OG_TYPE_FLOAT : OG_TYPE_INT, wrong) == OG_E_TYPE, "type mismatch error");
            }
        }
    }
}

int main() {
    test_window();
    test_count();
    test_sum();
    test_min_max();
    test_first_last();
    test_interval_record();
    test_codecs();
    printf("%d checks, %d failures\n", g_checks, g_fail);
    return g_fail ? 1 : 0;
}
