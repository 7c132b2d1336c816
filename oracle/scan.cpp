/*
 * scan.cpp — CPU ORACLE (test infrastructure, see og_oracle.h): the reference-structured pull loop over a shard and
 * the host-side synthetic shard builder.
 *
 * Control structure follows the reference: per series, pull one <=1000-row segment at a time
 * (Location.readData engine/immutable/location.go:261-315) -> decode every needed column into a Record
 * (tssp_file.go:369 readSegmentRecord) -> FilterByTime (reader.go:754) -> FilterByField via the RPN bitmaps
 * (reader.go:895, lib/binaryfilterfunc/functions.go:632) -> aggregateCursor (aggregate_cursor.go:267) ->
 * AggTagSetCursor.RecordInit (agg_tagset_cursor.go:1069) into the dense interval record.
 */
#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>

#include "../include/ogpu_synth.h"
#include "og_oracle.h"

namespace ogo {

namespace {

inline double as_f(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }

/* one term of the WHERE RPN on a decoded column: "valid && !(complement test)" (eval_generator.gen.go:52-144, SURVEY App.B.12) */
bool term_pass(const og_filter_item &it, int type, uint64_t raw) {
    if (type == OG_TYPE_FLOAT || (type == OG_TYPE_INT && it.const_is_float)) {
        double v = type == OG_TYPE_FLOAT ? as_f(raw) : (double)(int64_t)raw; /* Int64ToFloat64Slice functions.go:439 */
        double c = it.const_is_float ? it.fval : (double)it.ival;
        switch (it.op) {
        case OG_OP_LT: return !(v >= c);
        case OG_OP_LTE: return !(v > c);
        case OG_OP_GT: return !(v <= c);
        case OG_OP_GTE: return !(v < c);
        case OG_OP_EQ: return !(v != c);
        default: return !(v == c);
        }
    }
    int64_t v = type == OG_TYPE_BOOL ? (int64_t)(raw != 0) : (int64_t)raw, c = it.ival;
    switch (it.op) {
    case OG_OP_LT: return !(v >= c);
    case OG_OP_LTE: return !(v > c);
    case OG_OP_GT: return !(v <= c);
    case OG_OP_GTE: return !(v < c);
    case OG_OP_EQ: return !(v != c);
    default: return !(v == c);
    }
}

struct SeriesFeeder { /* plays the role of fileCursor/Location for one series */
    const og_shard_desc *sh;
    const og_query_desc *q;
    std::vector<int> cols;          /* shard field columns materialised, in in_schema order */
    std::vector<Field> in_schema;
    uint32_t seg, seg_end;
    Record ring[4]; /* record pool: the two most recently returned records stay untouched (the aggregate cursor holds the
                       current record while it peeks the next one; CircularRecordPool rings, engine/iterators.go:61-70) */
    int held[2] = {-1, -1};
    int pick(int other) const { for (int i = 0; i < 4; i++) if (i != held[0] && i != held[1] && i != other) return i; return 0; }
    const Record *give(int slot) { held[0] = held[1]; held[1] = slot; return &ring[slot]; }
    uint64_t rows_decoded = 0, segments = 0, page_bytes = 0;
    int error = 0;

    const Record *next() {
        while (seg < seg_end) {
            uint32_t s = seg++;
            if (sh->seg_tmax[s] < q->tmin || sh->seg_tmin[s] > q->tmax) continue; /* tr.Overlaps location.go:276-280 */
            int r_slot = pick(-1);
            Record &r = ring[r_slot];
            r.reset();
            ColVal &tc = r.cols.back();
            int rc = decode_time_page(sh->data + sh->time_page_off[s], sh->time_page_len[s], tc);
            if (rc != E_OK) { error = rc; return nullptr; }
            page_bytes += sh->time_page_len[s];
            int rows = tc.len;
            for (size_t k = 0; k < cols.size(); k++) {
                const og_column_desc &cd = sh->columns[cols[k]];
                ColVal &c = r.cols[k];
                if (cd.page_len[s] == 0) { /* column absent in this chunk: all rows null */
                    c.init(); c.len = rows; c.nil_count = rows; c.bitmap.assign((size_t)(rows + 7) / 8, 0);
                    continue;
                }
                rc = decode_field_page(sh->data + cd.page_off[s], cd.page_len[s], cd.type, c);
                if (rc != E_OK) { error = rc; return nullptr; }
                if (c.len != rows) { error = E_CORRUPT; return nullptr; }
                page_bytes += cd.page_len[s];
            }
            rows_decoded += (uint64_t)rows; segments++;
            /* FilterByTime reader.go:754-771 and FilterByField :895-974, as one row mask + gather (genRecByRowNumbers :809-860) */
            const int64_t *t = tc.integers();
            bool all_in = q->tmin <= t[0] && t[rows - 1] <= q->tmax;
            /* KickNilRow (lib/record/record.go:1188-1240, called at file_cursor.go:293,334,349 after the filters): rows whose
               field columns are all null are dropped; fast path when any column has no nulls */
            bool may_kick = true;
            for (size_t k = 0; k + 1 < r.cols.size(); k++) if (r.cols[k].len != 0 && r.cols[k].nil_count == 0) may_kick = false;
            if (all_in && q->n_filter == 0 && !may_kick) return give(r_slot);
            std::vector<uint8_t> keep((size_t)rows, 1);
            if (!all_in) for (int i = 0; i < rows; i++) keep[i] = t[i] >= q->tmin && t[i] <= q->tmax;
            if (q->n_filter) {
                std::vector<std::vector<uint8_t>> stack;
                for (uint32_t fi = 0; fi < q->n_filter; fi++) {
                    const og_filter_item &it = q->filter[fi];
                    if (it.kind == OG_F_TERM) {
                        int k = -1;
                        for (size_t x = 0; x < cols.size(); x++) if (cols[x] == it.column) k = (int)x;
                        const ColVal &c = r.cols[k];
                        int type = sh->columns[it.column].type;
                        std::vector<uint8_t> m((size_t)rows, 0);
                        int vi = 0;
                        for (int i = 0; i < rows; i++) {
                            if (c.is_nil(i)) continue;
                            uint64_t raw;
                            if (type == OG_TYPE_BOOL) raw = c.val[vi]; else memcpy(&raw, c.val.data() + 8 * (size_t)vi, 8);
                            vi++;
                            m[i] = term_pass(it, type, raw);
                        }
                        stack.push_back(std::move(m));
                    } else {
                        std::vector<uint8_t> b = std::move(stack.back()); stack.pop_back();
                        std::vector<uint8_t> &a = stack.back();
                        for (int i = 0; i < rows; i++) a[i] = it.kind == OG_F_AND ? (a[i] & b[i]) : (a[i] | b[i]);
                    }
                }
                for (int i = 0; i < rows; i++) keep[i] &= stack.back()[i];
            }
            if (may_kick)
                for (int i = 0; i < rows; i++) {
                    if (!keep[i]) continue;
                    bool all_nil = true;
                    for (size_t k = 0; k + 1 < r.cols.size(); k++) all_nil = all_nil && r.cols[k].is_nil(i);
                    if (all_nil) keep[i] = 0;
                }
            /* gather surviving rows */
            int o_slot = pick(r_slot);
            Record &o = ring[o_slot];
            o.reset();
            for (size_t k = 0; k < r.cols.size(); k++) {
                const ColVal &c = r.cols[k];
                ColVal &d = o.cols[k];
                int type = k + 1 == r.cols.size() ? OG_TYPE_INT : in_schema[k].type;
                int vi = 0;
                for (int i = 0; i < rows; i++) {
                    bool nil = c.is_nil(i);
                    if (keep[i]) {
                        if (nil) d.append_null(type, false);
                        else if (type == OG_TYPE_BOOL) d.append_boolean(c.val[vi] != 0);
                        else { int64_t x; memcpy(&x, c.val.data() + 8 * (size_t)vi, 8); d.append_integer(x); }
                    }
                    if (!nil) vi++;
                }
            }
            if (o.row_nums() == 0) continue;
            return give(o_slot);
        }
        return nullptr;
    }
};

const Record *feeder_next(void *ctx) { return ((SeriesFeeder *)ctx)->next(); }

/* TransIntervalRec2Rec record.go:1340-1358: rows with >=1 non-null field, as a Record of partial rows */
void interval_to_record(const IntervalRecord &ir, Record &rec) {
    size_t ncol = ir.schema.size() - 1;
    rec = Record(ir.schema);
    for (uint32_t i = 0; i < ir.n_rows; i++) {
        bool any = false;
        for (size_t k = 0; k < ncol; k++) any |= ir.valid[k][i] != 0;
        if (!any) continue;
        for (size_t k = 0; k < ncol; k++) {
            int type = ir.schema[k].type;
            if (!ir.valid[k][i]) rec.cols[k].append_null(type, false);
            else if (type == OG_TYPE_BOOL) rec.cols[k].append_boolean(ir.values[k][i] != 0);
            else rec.cols[k].append_integer((int64_t)ir.values[k][i]);
            rec.meta_times[k].push_back(ir.col_times[k][i]);
        }
        rec.append_time(ir.times[i]);
    }
}

} // namespace

int scan_aggregate(const og_shard_desc &sh, const og_query_desc &q_in, int threads, uint32_t series_begin,
                   uint32_t series_end, ScanResult &out) {
    og_query_desc q = q_in; /* influxql.MinTime/MaxTime (ast.go:92,102) bound the range */
    q.tmin = std::max(q_in.tmin, (int64_t)(INT64_MIN + 2)); q.tmax = std::min(q_in.tmax, (int64_t)(INT64_MAX - 1));
    /* descending scans: the dense interval record holds the same aggregates; only the emission order of rows differs (it is not
     * part of ScanResult).  The reference's reversed-record tie-breaks are NOT restated — parity unpinned for them. */
    if (series_end > sh.n_series) series_end = sh.n_series;
    /* schemas: the input record holds every column a call or the filter touches (+ time) */
    std::vector<int> cols;
    auto add_col = [&](int c) { if (std::find(cols.begin(), cols.end(), c) == cols.end()) cols.push_back(c); };
    for (uint32_t i = 0; i < q.n_calls; i++) { if (q.calls[i].column < 0 || (uint32_t)q.calls[i].column >= sh.n_columns) return E_INVAL; add_col(q.calls[i].column); }
    for (uint32_t i = 0; i < q.n_filter; i++) if (q.filter[i].kind == OG_F_TERM) { if ((uint32_t)q.filter[i].column >= sh.n_columns) return E_INVAL; add_col(q.filter[i].column); }
    std::sort(cols.begin(), cols.end());
    std::vector<Field> in_schema, out_schema;
    for (int c : cols) in_schema.push_back({sh.columns[c].name ? sh.columns[c].name : ("c" + std::to_string(c)), sh.columns[c].type});
    in_schema.push_back({"time", OG_TYPE_INT});
    std::vector<ExprOpt> exprs;
    for (uint32_t i = 0; i < q.n_calls; i++) {
        int c = q.calls[i].column;
        std::string in_name = sh.columns[c].name ? sh.columns[c].name : ("c" + std::to_string(c));
        std::string out_name = "o" + std::to_string(i);
        int ot = q.calls[i].func == OG_AGG_COUNT ? OG_TYPE_INT : sh.columns[c].type;
        out_schema.push_back({out_name, ot});
        exprs.push_back({q.calls[i].func, in_name, out_name});
    }
    out_schema.push_back({"time", OG_TYPE_INT});

    WindowOpt w; w.interval = q.interval; w.offset = q.offset; w.start_time = q.tmin; w.end_time = q.tmax;
    /* TimeWindowsInit agg_tagset_cursor.go:1012-1027 with FileInfo.{Min,Max}Time = query range (updateQueryTime :448-463) */
    /* the file range is intersected with the query range before it reaches TimeWindowsInit (fileLoopCursor.updateQueryTime) */
    int64_t sh_min = INT64_MAX, sh_max = INT64_MIN;
    for (uint32_t g = 0; g < sh.n_segments; g++) { sh_min = std::min(sh_min, sh.seg_tmin[g]); sh_max = std::max(sh_max, sh.seg_tmax[g]); }
    int64_t gmin = std::max(q.tmin, sh_min), gmax = std::min(q.tmax, sh_max);
    if (gmin > gmax) gmin = gmax = q.tmin; /* no overlap: one empty window */
    int64_t min_s, min_e, max_s, max_e;
    if (q.interval == 0) { min_s = gmin; min_e = gmax + 1; max_s = min_s; max_e = min_e; }
    else { window(w, gmin, &min_s, &min_e); window(w, gmax + 1, &max_s, &max_e); }
    int64_t interval_time = min_e - min_s;
    bool has_interval = q.interval != 0;

    uint32_t n_groups = q.group_mode == OG_GROUP_ALL ? 1 : q.group_mode == OG_GROUP_PER_SERIES ? sh.n_series : q.n_groups;
    auto group_of = [&](uint32_t s) -> uint32_t { return q.group_mode == OG_GROUP_ALL ? 0 : q.group_mode == OG_GROUP_PER_SERIES ? s : q.series_group[s]; };

    if (threads < 1) threads = 1;
    struct Worker { std::vector<IntervalRecord> groups; std::vector<uint8_t> touched; uint64_t rows = 0, segs = 0, bytes = 0; int error = 0; };
    std::vector<Worker> workers((size_t)threads);
    auto make_ir = [&]() { IntervalRecord ir; ir.schema = out_schema; ir.exprs = exprs; ir.multi = exprs.size() > 1; ir.build(min_s, max_e, interval_time, has_interval); return ir; };

    auto work = [&](int tid) {
        Worker &wk = workers[(size_t)tid];
        wk.groups.resize(n_groups); wk.touched.assign(n_groups, 0);
        for (uint32_t s = series_begin + (uint32_t)tid; s < series_end; s += (uint32_t)threads) { /* start, step striding file_cursor.go:190-195 */
            SeriesFeeder f;
            f.sh = &sh; f.q = &q; f.cols = cols; f.in_schema = in_schema;
            f.seg = sh.series_seg_begin[s]; f.seg_end = sh.series_seg_begin[s + 1];
            for (auto &r : f.ring) r = Record(in_schema);
            AggCursor *ac = agg_cursor_new(in_schema, out_schema, exprs, w, q.chunk_size > 0 ? q.chunk_size : 1024);
            if (!ac) { wk.error = E_INVAL; return; }
            agg_cursor_set_input(ac, feeder_next, &f);
            uint32_t g = group_of(s);
            if (!wk.touched[g]) { wk.groups[g] = make_ir(); wk.touched[g] = 1; }
            while (const Record *r = agg_cursor_next(ac)) wk.groups[g].update_from(*r);
            agg_cursor_free(ac);
            wk.rows += f.rows_decoded; wk.segs += f.segments; wk.bytes += f.page_bytes;
            if (f.error) { wk.error = f.error; return; }
        }
    };
    if (threads == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++) th.emplace_back(work, t);
        for (auto &t : th) t.join();
    }
    for (auto &wk : workers) if (wk.error) return wk.error;

    /* cross-worker merge: what the executor's re-aggregation does with per-cursor partials (agg_transform.go:248-304) */
    IntervalRecord proto = make_ir();
    out.n_groups = n_groups; out.n_buckets = proto.n_rows; out.start = proto.start; out.interval = has_interval ? interval_time : 0;
    size_t ncol = exprs.size();
    out.values.assign(ncol, std::vector<uint64_t>((size_t)n_groups * proto.n_rows, 0));
    out.valid.assign(ncol, std::vector<uint8_t>((size_t)n_groups * proto.n_rows, 0));
    out.times.assign(ncol, std::vector<int64_t>((size_t)n_groups * proto.n_rows, 0));
    for (uint32_t g = 0; g < n_groups; g++) {
        IntervalRecord *final_ir = nullptr;
        IntervalRecord merged;
        int holders = 0;
        for (auto &wk : workers) if (wk.touched[g]) holders++;
        if (holders == 1) { for (auto &wk : workers) if (wk.touched[g]) final_ir = &wk.groups[g]; }
        else if (holders > 1) {
            merged = make_ir();
            for (auto &wk : workers) if (wk.touched[g]) { Record r; interval_to_record(wk.groups[g], r); merged.update_from(r); }
            final_ir = &merged;
        }
        for (size_t k = 0; k < ncol; k++) {
            for (uint32_t b = 0; b < proto.n_rows; b++) {
                size_t o = (size_t)g * proto.n_rows + b;
                if (final_ir) {
                    out.values[k][o] = final_ir->values[k][b]; out.valid[k][o] = final_ir->valid[k][b];
                    bool sel = exprs[k].func >= OG_AGG_MIN;
                    out.times[k][o] = !sel ? proto.times[b] : (final_ir->multi ? (exprs[k].func >= OG_AGG_FIRST ? final_ir->col_times[k][b] : proto.times[b]) : final_ir->times[b]);
                } else out.times[k][o] = proto.times[b];
            }
        }
    }
    for (auto &wk : workers) { out.rows_decoded += wk.rows; out.segments += wk.segs; out.page_bytes += wk.bytes; }
    return E_OK;
}

/* ===================== synthetic shard (host) ===================== */
og_shard_desc HostShard::desc() {
    col_descs.resize(col_types.size());
    for (size_t c = 0; c < col_types.size(); c++) {
        col_descs[c].name = col_names[c].c_str(); col_descs[c].type = col_types[c];
        col_descs[c].page_off = page_off[c].data(); col_descs[c].page_len = page_len[c].data();
    }
    og_shard_desc d; memset(&d, 0, sizeof d);
    d.data = data.data(); d.data_len = data.size();
    d.n_series = (uint32_t)sids.size(); d.sids = sids.data(); d.series_seg_begin = series_seg_begin.data();
    d.n_segments = (uint32_t)seg_tmin.size(); d.seg_tmin = seg_tmin.data(); d.seg_tmax = seg_tmax.data();
    d.n_columns = (uint32_t)col_types.size(); d.columns = col_descs.data();
    d.time_page_off = page_off.back().data(); d.time_page_len = page_len.back().data();
    return d;
}

int build_synth_shard(const og_synth_desc &d, HostShard &out, int threads) {
    uint32_t rps = d.rows_per_segment ? d.rows_per_segment : 1000;
    uint32_t segs_per_series = (d.rows_per_series + rps - 1) / rps;
    uint32_t nseg = d.n_series * segs_per_series;
    out.data.clear(); out.sids.resize(d.n_series); out.series_seg_begin.resize(d.n_series + 1);
    out.seg_tmin.resize(nseg); out.seg_tmax.resize(nseg);
    out.page_off.assign(d.n_columns + 1, std::vector<uint64_t>(nseg));
    out.page_len.assign(d.n_columns + 1, std::vector<uint32_t>(nseg));
    out.col_types.resize(d.n_columns); out.col_names.resize(d.n_columns);
    for (uint32_t c = 0; c < d.n_columns; c++) { out.col_types[c] = d.columns[c].type; out.col_names[c] = "f" + std::to_string(c); }
    /* layout mirrors a TSSP chunk: per series, each field column's pages back to back, then the time pages
       (chunkdata_builder_ts.go:36-82); the 4-byte per-column CRC slots are kept so offsets look like a real file.
       Series ranges are encoded by worker threads into private buffers and concatenated in series order. */
    if (threads < 1) threads = 1;
    if ((uint32_t)threads > d.n_series) threads = (int)d.n_series;
    std::vector<Bytes> bufs((size_t)threads);
    std::vector<int> rcs((size_t)threads, E_OK);
    auto range_of = [&](int t, uint32_t *a, uint32_t *b) { *a = (uint32_t)((uint64_t)d.n_series * t / threads); *b = (uint32_t)((uint64_t)d.n_series * (t + 1) / threads); };
    auto build_range = [&](int tid) -> int {
    uint32_t s_a, s_b; range_of(tid, &s_a, &s_b);
    Bytes &data = bufs[(size_t)tid];
    std::vector<int64_t> times(rps);
    for (uint32_t s = s_a; s < s_b; s++) {
        out.sids[s] = (uint64_t)d.series_base + s + 1;
        out.series_seg_begin[s] = s * segs_per_series;
        for (uint32_t c = 0; c <= d.n_columns; c++) {
            data.insert(data.end(), 4, 0); /* crc32 placeholder (not verified by the attached read path) */
            for (uint32_t g = 0; g < segs_per_series; g++) {
                uint32_t seg = s * segs_per_series + g;
                uint64_t row0 = (uint64_t)g * rps;
                uint32_t n = (uint32_t)std::min<uint64_t>(rps, d.rows_per_series - row0);
                uint64_t off = data.size();
                int rc;
                if (c == d.n_columns) {
                    for (uint32_t i = 0; i < n; i++) times[i] = d.t0 + (int64_t)(row0 + i) * d.dt;
                    out.seg_tmin[seg] = times[0]; out.seg_tmax[seg] = times[n - 1];
                    rc = encode_time_page(times.data(), n, data);
                } else {
                    const og_synth_column &sc = d.columns[c];
                    ColVal cv;
                    int64_t walk = 0;
                    const uint32_t ps = d.series_base + s; /* position of the series in the synthetic population */
                    for (uint32_t i = 0; i < n; i++) {
                        uint64_t row = row0 + i;
                        bool nil = og_synth_is_null(d.seed, c, ps, row, sc.null_permille);
                        switch (sc.dist) {
                        case OG_SYNTH_F_HI: if (nil) cv.append_null(OG_TYPE_FLOAT, false); else cv.append_float(og_synth_f_hi(d.seed, c, ps, row)); break;
                        case OG_SYNTH_F_LO:
                            walk = i == 0 ? og_synth_walk_first(d.seed, c, ps, g, 1) : walk + og_synth_f_lo_step(d.seed, c, ps, row);
                            if (nil) cv.append_null(OG_TYPE_FLOAT, false); else cv.append_float((double)walk); break;
                        case OG_SYNTH_INT_WALK:
                            walk = i == 0 ? og_synth_walk_first(d.seed, c, ps, g, 0) : walk + og_synth_int_step(d.seed, c, ps, row);
                            if (nil) cv.append_null(OG_TYPE_INT, false); else cv.append_integer(walk); break;
                        default:
                            if (nil) cv.append_null(OG_TYPE_BOOL, false); else cv.append_boolean(og_synth_bool(d.seed, c, ps, row) != 0); break;
                        }
                    }
                    rc = encode_field_page(cv, sc.type, data);
                }
                if (rc != E_OK) return rc;
                out.page_off[c][seg] = off;
                out.page_len[c][seg] = (uint32_t)(data.size() - off);
            }
        }
    }
    return E_OK;
    };
    if (threads == 1) rcs[0] = build_range(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < threads; t++) th.emplace_back([&, t]() { rcs[(size_t)t] = build_range(t); });
        for (auto &t : th) t.join();
    }
    for (int rc : rcs) if (rc != E_OK) return rc;
    uint64_t total = 0;
    std::vector<uint64_t> base((size_t)threads);
    for (int t = 0; t < threads; t++) { base[(size_t)t] = total; total += bufs[(size_t)t].size(); }
    out.data.resize(total);
    for (int t = 0; t < threads; t++) {
        if (!bufs[(size_t)t].empty()) memcpy(out.data.data() + base[(size_t)t], bufs[(size_t)t].data(), bufs[(size_t)t].size());
        uint32_t s_a, s_b; range_of(t, &s_a, &s_b);
        for (uint32_t c = 0; c <= d.n_columns; c++)
            for (uint32_t seg = s_a * segs_per_series; seg < s_b * segs_per_series; seg++) out.page_off[c][seg] += base[(size_t)t];
        Bytes().swap(bufs[(size_t)t]);
    }
    out.series_seg_begin[d.n_series] = nseg;
    out.data.reserve(out.data.size() + 32); /* word-granular readers (fast_scan.cpp) may look a few bytes past the last page; desc().data_len excludes the slack */
    return E_OK;
}

} // namespace ogo
