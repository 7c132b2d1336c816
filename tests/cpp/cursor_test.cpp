/* GPU parity test of the C++ KeyCursor mirror (opengemini_b200/host) against the CPU oracle.
 *
 * Shaped after the reference's own cursor tests (engine/iterators_test.go:748-2043: build records → cursor.SetOps/SinkPlan →
 * loop Next() until nil → compare with expected records), with the expected side computed by oracle/ (test infrastructure).
 * Run by tests/test_gpu_cursor_cpp.py on a GPU box; exit code 0 = all checks passed.
 */
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>

#include "../../opengemini_b200/host/gpu_cursor.h"
#include "../../oracle/og_oracle.h"

static int g_fail = 0, g_checks = 0;
#define CHECK(cond, ...) do { g_checks++; if (!(cond)) { if (g_fail++ < 25) { std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); } } } while (0)

using namespace ogpu;

struct Case {
    const char *name;
    std::vector<CallOption> ops;
    QueryPlan plan;
};

static int func_id(const std::string &s) {
    return s == "count" ? OG_AGG_COUNT : s == "sum" ? OG_AGG_SUM : s == "min" ? OG_AGG_MIN : s == "max" ? OG_AGG_MAX : s == "first" ? OG_AGG_FIRST : OG_AGG_LAST;
}

static void run_case(const Case &cs, ogo::HostShard &hs, std::shared_ptr<GpuShard> shard) {
    og_shard_desc sd = hs.desc();
    /* oracle side */
    std::vector<og_call> calls;
    for (auto &op : cs.ops) calls.push_back(og_call{func_id(op.Call), op.Ref == "f0" ? 0 : 1});
    og_query_desc qd{};
    qd.interval = cs.plan.Interval; qd.offset = cs.plan.Offset; qd.tmin = cs.plan.StartTime; qd.tmax = cs.plan.EndTime;
    qd.ascending = 1; qd.n_calls = (uint32_t)calls.size(); qd.calls = calls.data();
    qd.group_mode = cs.plan.GroupBy == QueryPlan::GroupAll ? OG_GROUP_ALL : cs.plan.GroupBy == QueryPlan::GroupBySeries ? OG_GROUP_PER_SERIES : OG_GROUP_MAP;
    qd.n_groups = cs.plan.NumGroups; qd.series_group = cs.plan.SeriesGroup.empty() ? nullptr : cs.plan.SeriesGroup.data();
    qd.chunk_size = cs.plan.ChunkSize;
    ogo::ScanResult exp;
    int rc = ogo::scan_aggregate(sd, qd, 4, 0, sd.n_series, exp);
    CHECK(rc == 0, "%s: oracle scan rc=%d", cs.name, rc);

    /* cursor side, the way ChunkReader.nextRecord drains a KeyCursor (iterator_plan.go:707-717) */
    GpuAggCursor cur(shard);
    cur.SetOps(cs.ops);
    Error e = cur.SinkPlan(cs.plan);
    CHECK(!e, "%s: SinkPlan: %s", cs.name, e.Error_());
    CHECK(cur.GetSchema().size() == cs.ops.size() + 1 && cur.GetSchema().back().Name == std::string(TimeField), "%s: schema", cs.name);
    std::map<uint64_t, size_t> seen; /* (group,bucket) -> rows */
    size_t n_rows = 0, n_recs = 0;
    int64_t last_t = INT64_MIN; uint32_t last_g = 0;
    for (;;) {
        const Record *rec = nullptr; const SeriesInfo *info = nullptr;
        e = cur.Next(&rec, &info);
        CHECK(!e, "%s: Next: %s", cs.name, e.Error_());
        if (e || !rec) break;
        n_recs++;
        CHECK(rec->RowNums() > 0 && rec->RowNums() <= (cs.plan.ChunkSize > 0 ? cs.plan.ChunkSize : 1 << 30), "%s: rows %d", cs.name, rec->RowNums());
        uint32_t g = info->group;
        if (g != last_g) last_t = INT64_MIN;
        last_g = g;
        for (int r = 0; r < rec->RowNums(); r++) {
            int64_t t = rec->Times()[r];
            CHECK(t > last_t || cs.plan.Interval == 0, "%s: time order", cs.name);
            last_t = t;
            /* which bucket: rows carry the window start (multi-call / sum / count) or the selector's time */
            int64_t b = cs.plan.Interval ? (t - exp.start) / exp.interval : 0;
            size_t idx = (size_t)g * exp.n_buckets + (size_t)b;
            CHECK(b >= 0 && (uint32_t)b < exp.n_buckets && g < exp.n_groups, "%s: bucket range", cs.name);
            if (b < 0 || (uint32_t)b >= exp.n_buckets || g >= exp.n_groups) continue;
            seen[idx]++;
            n_rows++;
            bool any = false;
            for (size_t c = 0; c < cs.ops.size(); c++) {
                const ColVal &cv = rec->ColVals[c];
                bool nil = cv.IsNil(r);
                CHECK(nil == !exp.valid[c][idx], "%s: col %zu validity g=%u b=%lld", cs.name, c, g, (long long)b);
                if (nil || !exp.valid[c][idx]) continue;
                any = true;
                uint64_t got; std::memcpy(&got, cv.Val + 8 * (size_t)r, 8);
                if (cur.GetSchema()[c].Type == Field_Type_Float && func_id(cs.ops[c].Call) == OG_AGG_SUM) {
                    double a, x; std::memcpy(&a, &got, 8); std::memcpy(&x, &exp.values[c][idx], 8);
                    CHECK(std::fabs(a - x) <= 1e-9 * std::fabs(x), "%s: float sum %g vs %g", cs.name, a, x); /* north_star tolerance */
                } else {
                    CHECK(got == exp.values[c][idx], "%s: col %zu value g=%u b=%lld", cs.name, c, g, (long long)b);
                }
            }
            CHECK(any, "%s: all-nil row survived TransIntervalRec2Rec", cs.name);
        }
    }
    /* every non-empty oracle cell was emitted exactly once */
    size_t want = 0;
    for (size_t idx = 0; idx < (size_t)exp.n_groups * exp.n_buckets; idx++) {
        bool any = false;
        for (size_t c = 0; c < cs.ops.size(); c++) any |= exp.valid[c][idx] != 0;
        if (any) { want++; CHECK(seen.count(idx) && seen[idx] == 1, "%s: cell %zu emitted %zu times", cs.name, idx, seen.count(idx) ? seen[idx] : 0); }
    }
    CHECK(want == n_rows, "%s: %zu rows, oracle has %zu non-empty cells", cs.name, n_rows, want);
    /* end of stream is sticky: (nil, nil, nil) again */
    const Record *rec = nullptr; const SeriesInfo *info = nullptr;
    e = cur.Next(&rec, &info);
    CHECK(!e && rec == nullptr, "%s: EOF not sticky", cs.name);
    CHECK(!cur.Close() && !cur.Close(), "%s: Close", cs.name);
    e = cur.Next(&rec, &info);
    CHECK(e.code == OG_E_STATE, "%s: Next after Close must fail, got %d", cs.name, e.code);
    std::printf("ok   %-28s records=%zu rows=%zu\n", cs.name, n_recs, n_rows);
}

int main() {
    int st = og_init(0);
    if (st != OG_OK) { std::printf("og_init failed: %s %s\n", og_strerror(st), og_last_error()); return 2; }
    /* SURVEY §8d config 1 shape: 100 series x 10^4 rows, 1000-row segments, 1 s cadence */
    og_synth_column cols[2] = {{OG_TYPE_FLOAT, OG_SYNTH_F_HI, 0}, {OG_TYPE_INT, OG_SYNTH_INT_WALK, 50}};
    og_synth_desc sdsc{};
    sdsc.n_series = 100; sdsc.rows_per_series = 10000; sdsc.rows_per_segment = 1000;
    sdsc.t0 = 1700000000LL * 1000000000LL; sdsc.dt = 1000000000LL; sdsc.seed = 7; sdsc.n_columns = 2; sdsc.columns = cols;
    ogo::HostShard hs;
    if (ogo::build_synth_shard(sdsc, hs, 4) != 0) { std::printf("oracle shard build failed\n"); return 2; }
    og_shard_desc sd = hs.desc();
    std::shared_ptr<GpuShard> shard;
    Error e = GpuShard::Open(sd, &shard);
    if (e) { std::printf("open: %s\n", e.Error_()); return 2; }
    Schemas sch = {{Field_Type_Float, "f0"}, {Field_Type_Int, "f1"}, {Field_Type_Int, TimeField}};
    shard->set_schema(sch);

    const int64_t MIN = 60LL * 1000000000LL;
    std::vector<Case> cases;
    { Case c{"mean_max_by_1m", {{"sum", "f0"}, {"count", "f0"}, {"max", "f0"}}, {}}; c.plan.Interval = MIN; cases.push_back(c); }
    { Case c{"single_max_selector_time", {{"max", "f0"}}, {}}; c.plan.Interval = MIN; cases.push_back(c); }
    { Case c{"int_nulls_chunk7", {{"count", "f1"}, {"sum", "f1"}, {"min", "f1"}, {"last", "f1"}}, {}}; c.plan.Interval = 5 * MIN; c.plan.ChunkSize = 7; cases.push_back(c); }
    { Case c{"per_series_first_last", {{"first", "f0"}, {"last", "f0"}}, {}}; c.plan.Interval = 10 * MIN; c.plan.GroupBy = QueryPlan::GroupBySeries; c.plan.NumGroups = 100; cases.push_back(c); }
    { Case c{"tagset_map_range", {{"sum", "f1"}, {"max", "f0"}}, {}}; c.plan.Interval = MIN; c.plan.GroupBy = QueryPlan::GroupByTagSet; c.plan.NumGroups = 3;
      for (uint32_t s = 0; s < 100; s++) c.plan.SeriesGroup.push_back(s % 3);
      c.plan.StartTime = sdsc.t0 + 1234 * sdsc.dt + 5; c.plan.EndTime = sdsc.t0 + 8765 * sdsc.dt; cases.push_back(c); }
    { Case c{"no_interval", {{"count", "f0"}, {"sum", "f0"}}, {}}; c.plan.Interval = 0; cases.push_back(c); }
    for (auto &c : cases) run_case(c, hs, shard);

    /* error behaviour: unknown call / unknown field are errors, not panics; Next before SinkPlan is OG_E_STATE */
    {
        GpuAggCursor cur(shard);
        const Record *rec; const SeriesInfo *info;
        CHECK(cur.Next(&rec, &info).code == OG_E_STATE, "Next before SinkPlan");
        cur.SetOps({{"percentile", "f0"}});
        CHECK(cur.SinkPlan(QueryPlan{}).code == OG_E_UNSUPPORTED, "unknown call");
        cur.SetOps({{"sum", "nope"}});
        CHECK(cur.SinkPlan(QueryPlan{}).code == OG_E_INVAL, "unknown field");
    }
    std::printf("%d checks, %d failures\n", g_checks, g_fail);
    return g_fail ? 1 : 0;
}
