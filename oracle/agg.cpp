/*
 * agg.cpp — CPU ORACLE (test infrastructure, see og_oracle.h): restatement of the per-series windowed aggregation
 * (engine/aggregate_cursor.go, series_call_processor.go, series_agg_reducer.gen.go, series_agg_func.gen.go,
 * lib/record/column_util.go) and of the tagset-level dense interval record
 * (engine/agg_tagset_cursor.go:959-1120, lib/record/reccord_functions.go, lib/record/record.go:1298-1365).
 */
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "og_oracle.h"

namespace ogo {

/* ===================== ProcessorOptions.Window (lib/util/lifted/influx/query/select.go:579-655, Location == nil) ===================== */
static const int64_t MIN_TIME = INT64_MIN + 2; /* influxql.MinTime ast.go:92 */
static const int64_t MAX_TIME = INT64_MAX - 1; /* influxql.MaxTime ast.go:102 */

void window(const WindowOpt &o, int64_t t, int64_t *start, int64_t *end) {
    if (o.interval == 0) { *start = o.start_time; *end = o.end_time + 1; return; } /* :580-582 */
    t -= o.offset;
    int64_t dt = t % o.interval;
    if (dt < 0) dt += o.interval;                                                  /* :594-599 */
    int64_t s;
    if ((int64_t)((uint64_t)MIN_TIME + (uint64_t)dt) >= t) s = MIN_TIME; else s = t - dt; /* :602-606 */
    s += o.offset;
    int64_t d2 = o.interval - dt, e;
    if (MAX_TIME - d2 <= t) e = MAX_TIME; else e = t + d2;                         /* :622-626 */
    e += o.offset;                                                                 /* :654 */
    *start = s; *end = e;
}

/* ===================== reducers ===================== */
namespace {

struct ColBuf { /* floatColBuf / integerColBuf / booleanColBuf series_agg_reducer.gen.go:70-180 */
    int index = 0; int64_t time = 0; uint64_t value = 0; bool is_nil = true;
    void set(int i, int64_t t, uint64_t v) { index = i; time = t; value = v; is_nil = false; }
    void reset() { is_nil = true; }
    void assign(const ColBuf &s) { index = s.index; time = s.time; value = s.value; }
};

inline double as_f(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
inline uint64_t f_bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

struct ReduceOut { int index; uint64_t value; bool is_nil; };

/* generic restatement of lib/record/column_util.go:23-278 over 8-byte or 1-byte cells */
struct ValView {
    const ColVal *cv; int type;
    size_t n() const { return cv->n_values(type); }
    uint64_t raw(size_t i) const {
        if (type == OG_TYPE_BOOL) return cv->val[i];
        uint64_t u; memcpy(&u, cv->val.data() + 8 * i, 8); return u;
    }
    /* a > b / a < b under the column's type; bools: false < true (column_boolean.go:57-143) */
    bool gt(uint64_t a, uint64_t b) const {
        if (type == OG_TYPE_FLOAT) return as_f(a) > as_f(b);
        if (type == OG_TYPE_INT) return (int64_t)a > (int64_t)b;
        return a != b && a;
    }
    bool lt(uint64_t a, uint64_t b) const {
        if (type == OG_TYPE_FLOAT) return as_f(a) < as_f(b);
        if (type == OG_TYPE_INT) return (int64_t)a < (int64_t)b;
        return a != b && !a;
    }
};

inline bool bit_set(const ColVal &cv, int i) { int idx = cv.bitmap_offset + i; return (cv.bitmap[idx >> 3] & (1 << (idx & 7))) != 0; }

ReduceOut min_max_reduce(const ValView &vv, int start, int end, bool want_max) { /* minValue :190-233 / maxValue :235-278 */
    const ColVal &cv = *vv.cv;
    if (vv.n() == 0) return {0, 0, true};
    int row = -1; uint64_t best = 0;
    if (cv.nil_count == 0) {
        best = vv.raw(start); row = start;
        for (int i = start; i < end; i++) {
            uint64_t x = vv.raw(i);
            if (want_max ? vv.lt(best, x) : vv.gt(best, x)) { best = x; row = i; }
        }
        return {row, best, false};
    }
    int skip = cv.valid_count(0, start), vidx = skip;
    for (int i = start; i < end && (size_t)vidx < vv.n(); i++) {
        if (!bit_set(cv, i)) continue;
        uint64_t x = vv.raw(vidx);
        if (vidx == skip) { best = x; row = i; }
        else if (want_max ? vv.lt(best, x) : vv.gt(best, x)) { best = x; row = i; }
        vidx++;
    }
    if (row == -1) return {0, 0, true};
    return {row, best, false};
}

ReduceOut first_reduce(const ValView &vv, int start, int end) { /* firstValue :23-52 */
    const ColVal &cv = *vv.cv;
    if (vv.n() == 0) return {0, 0, true};
    if (cv.nil_count == 0) return {start, vv.raw(start), false};
    int vidx = cv.valid_count(0, start);
    for (int i = start; i < end && (size_t)vidx < vv.n(); i++) {
        if (!bit_set(cv, i)) continue;
        return {i, vv.raw(vidx), false};
    }
    return {0, 0, true};
}

ReduceOut last_reduce(const ValView &vv, int start, int end) { /* lastValue :54-85 */
    const ColVal &cv = *vv.cv;
    if (vv.n() == 0) return {0, 0, true};
    if (cv.nil_count == 0) return {end - 1, vv.raw(end - 1), false};
    int row = -1;
    for (int i = end - 1; i >= start; i--) { if (bit_set(cv, i)) { row = i; break; } }
    if (row < start) return {0, 0, true};
    return {row, vv.raw(cv.valid_count(0, row)), false};
}

ReduceOut count_reduce(const ValView &vv, int start, int end) { /* *CountReduce series_agg_func.gen.go:24-42 */
    int64_t c = vv.cv->valid_count(start, end);
    return {start, (uint64_t)c, c == 0};
}

ReduceOut sum_reduce(const ValView &vv, int start, int end) { /* floatSumReduce :48-60 / integerSumReduce :66-78 */
    const ColVal &cv = *vv.cv;
    if ((int)cv.val.size() + cv.nil_count == 0) return {start, 0, true};
    int s, e;
    cv.value_index_range(start, end, &s, &e);
    int agg = 0;
    if (vv.type == OG_TYPE_FLOAT) {
        double sum = 0;
        for (int i = s; i < e; i++) { sum += as_f(vv.raw(i)); agg++; }
        return {s, f_bits(sum), agg == 0}; /* quirk kept: the returned index is the VALUE index (:53,59) */
    }
    int64_t sum = 0;
    for (int i = s; i < e; i++) { sum = (int64_t)((uint64_t)sum + vv.raw(i)); agg++; }
    return {s, (uint64_t)sum, agg == 0};
}

struct Reducer {
    int func, in_type, out_type, in_ord, out_ord;
    bool time_col; /* first/last use the *TimeCol* reducers (series_call_processor.go:211-283) */
    ColBuf prev, curr;

    ReduceOut fn(const ColVal &cv, int start, int end) const {
        ValView vv{&cv, in_type};
        switch (func) {
        case OG_AGG_COUNT: return count_reduce(vv, start, end);
        case OG_AGG_SUM: return sum_reduce(vv, start, end);
        case OG_AGG_MIN: return min_max_reduce(vv, start, end, false);
        case OG_AGG_MAX: return min_max_reduce(vv, start, end, true);
        case OG_AGG_FIRST: return first_reduce(vv, start, end);
        default: return last_reduce(vv, start, end);
        }
    }
    void fv() { /* *Merge series_agg_func.gen.go:44-46,62-64,92-98,140-146,188,233 */
        ValView vv{nullptr, out_type == OG_TYPE_INT && in_type != OG_TYPE_INT && func == OG_AGG_COUNT ? OG_TYPE_INT : out_type};
        switch (func) {
        case OG_AGG_COUNT: prev.value = (uint64_t)((int64_t)prev.value + (int64_t)curr.value); break;
        case OG_AGG_SUM:
            if (out_type == OG_TYPE_FLOAT) prev.value = f_bits(as_f(prev.value) + as_f(curr.value));
            else prev.value = prev.value + curr.value;
            break;
        case OG_AGG_MIN: if (vv.lt(curr.value, prev.value)) prev.assign(curr); break;
        case OG_AGG_MAX: if (vv.gt(curr.value, prev.value)) prev.assign(curr); break;
        case OG_AGG_FIRST: break;
        default: prev.assign(curr); break;
        }
    }
    void append_value(ColVal &c, uint64_t v) const {
        if (out_type == OG_TYPE_BOOL) c.append_boolean(v != 0);
        else if (out_type == OG_TYPE_FLOAT) c.append_float(as_f(v));
        else c.append_integer((int64_t)v);
    }
};

struct Params { bool multi_call = false, same_window = false; const std::vector<uint16_t> *interval_index = nullptr; };

/* floatColFloatReducer.Aggregate series_agg_reducer.gen.go:206-300 and the *TimeCol* variant :1032-1132 */
void aggregate(Reducer &r, const Record &in, Record &out, const Params &p) {
    const std::vector<uint16_t> &ii = *p.interval_index;
    int first_index = 0, last_index = (int)ii.size() - 1;
    const ColVal &icol = in.cols[r.in_ord];
    ColVal &ocol = out.cols[r.out_ord];
    auto emit_time = [&](int64_t t, bool null_row) {
        if (!p.multi_call) out.append_time(t);
        else if (r.time_col) out.meta_times[r.out_ord].push_back(null_row ? 0 : t);
    };
    for (int i = 0; i <= last_index; i++) {
        int start = ii[i];
        int end = i < last_index ? (int)ii[i + 1] : in.row_nums();
        ReduceOut ro = r.fn(icol, start, end);
        int index = ro.index;
        if (!r.time_col && icol.nil_count == icol.len) index = start; /* :225-227 */
        if (!ro.is_nil) {
            if (i == first_index && !r.prev.is_nil) {
                r.curr.set(index + 1, in.time(index), ro.value);
                r.fv();
                if (first_index == last_index && p.same_window) {
                    r.prev.index = 0;
                } else {
                    r.append_value(ocol, r.prev.value);
                    emit_time(r.prev.time, false);
                    r.prev.reset();
                }
                r.curr.reset();
                continue;
            } else if (i == last_index && p.same_window) {
                r.prev.set(0, in.time(index), ro.value);
                break;
            }
            r.append_value(ocol, ro.value);
            emit_time(in.time(index), false);
        } else {
            if ((i == first_index && !r.prev.is_nil) && (first_index < last_index || !p.same_window)) {
                r.append_value(ocol, r.prev.value);
                emit_time(r.prev.time, false);
                r.prev.reset();
                continue;
            } else if (i == last_index && p.same_window) {
                break;
            }
            ocol.append_null(r.out_type, false);
            emit_time(in.time(index), true);
        }
    }
}

} // namespace

/* ===================== aggregateCursor (engine/aggregate_cursor.go) ===================== */
struct AggCursor {
    std::vector<Field> in_schema, out_schema;
    WindowOpt w;
    int max_record_size = 1024;
    bool multi_call = false, init_col_meta = false, in_next_win = false;
    int time_ordinal = 0;
    std::vector<Reducer> reducers;
    NextFn fn = nullptr; void *ctx = nullptr;
    const Record *buf_record = nullptr;
    Record pool[2]; int pool_idx = 0; /* aggCursorRecordNum = 2 (engine/iterators.go:61-70) */
    std::vector<uint16_t> interval_index;
};

AggCursor *agg_cursor_new(const std::vector<Field> &in_schema, const std::vector<Field> &out_schema,
                          const std::vector<ExprOpt> &exprs, const WindowOpt &w, int chunk_size) {
    AggCursor *c = new AggCursor;
    c->in_schema = in_schema; c->out_schema = out_schema; c->w = w;
    c->max_record_size = chunk_size > 0 ? chunk_size : 1024; /* NewAggregateCursor :65-80 (no limit/offset) */
    c->time_ordinal = (int)out_schema.size() - 1;
    Record tmp_in(in_schema), tmp_out(out_schema);
    for (const ExprOpt &e : exprs) { /* newProcessor series_call_processor.go:26-85 */
        Reducer r;
        r.func = e.func;
        r.in_ord = tmp_in.field_index(e.in_name);
        r.out_ord = tmp_out.field_index(e.out_name);
        if (r.in_ord < 0 || r.out_ord < 0) { delete c; return nullptr; } /* "schemas are not aligned" panic */
        r.in_type = in_schema[r.in_ord].type;
        r.out_type = e.func == OG_AGG_COUNT ? OG_TYPE_INT : r.in_type;
        if (e.func == OG_AGG_SUM && r.in_type == OG_TYPE_BOOL) { delete c; return nullptr; } /* unsupported sum iterator type */
        r.time_col = (e.func == OG_AGG_FIRST || e.func == OG_AGG_LAST);
        if (r.time_col) c->init_col_meta = true;
        c->reducers.push_back(r);
    }
    c->multi_call = exprs.size() > 1;
    c->pool[0] = Record(out_schema); c->pool[1] = Record(out_schema);
    return c;
}
void agg_cursor_free(AggCursor *c) { delete c; }
void agg_cursor_set_input(AggCursor *c, NextFn fn, void *ctx) { c->fn = fn; c->ctx = ctx; }

static const Record *next_record(AggCursor *c) { /* nextRecord :257-264 */
    if (c->buf_record) { const Record *r = c->buf_record; c->buf_record = nullptr; return r; }
    return c->fn(c->ctx);
}

static void in_next_window(AggCursor *c, const Record *cur) { /* inNextWindow :314-341 */
    const Record *next = next_record(c); /* peekRecord */
    c->buf_record = next;
    if (!next || cur->row_nums() == 0) { c->in_next_win = false; return; }
    if (next->row_nums() == 0) { c->in_next_win = true; return; }
    if (c->w.interval == 0) { c->in_next_win = true; return; }
    int64_t last = cur->time(cur->row_nums() - 1), s, e;
    window(c->w, next->time(0), &s, &e);
    c->in_next_win = (s <= last && last < e);
}

static void get_interval_index(AggCursor *c, const Record *rec) { /* getIntervalIndex :343-356 */
    if (c->w.interval == 0) { c->interval_index.push_back(0); return; }
    int64_t s = 0, e = 0;
    const int64_t *t = rec->times();
    for (int i = 0; i < rec->row_nums(); i++) {
        if (i == 0 || t[i] >= e || t[i] < s) {
            c->interval_index.push_back((uint16_t)i);
            window(c->w, t[i], &s, &e);
        }
    }
}

const Record *agg_cursor_next(AggCursor *c) { /* Next :267-304 (NextAggData :90-142 is identical for one file) */
    Record *nr = &c->pool[c->pool_idx];
    c->pool_idx ^= 1;
    nr->reset();
    for (;;) {
        const Record *in = next_record(c);
        if (!in) return nr->row_nums() > 0 ? nr : nullptr;
        if (in->row_nums() == 0) continue;
        if (nr->row_nums() >= c->max_record_size) { c->buf_record = in; return nr; }
        in_next_window(c, in);
        /* reduce :306-312 */
        get_interval_index(c, in);
        Params p; p.multi_call = c->multi_call; p.same_window = c->in_next_win; p.interval_index = &c->interval_index;
        for (Reducer &r : c->reducers) aggregate(r, *in, *nr, p);
        if (c->multi_call) { /* deriveIntervalIndex :358-375 */
            int add = (int)c->interval_index.size() - (c->in_next_win ? 1 : 0);
            for (int i = 0; i < add; i++) nr->cols[c->time_ordinal].append_integer(in->time(c->interval_index[i]));
        }
        c->interval_index.clear();
        c->in_next_win = false;
    }
}

/* ===================== dense interval record (AggTagSetCursor) ===================== */
void IntervalRecord::build(int64_t min, int64_t max, int64_t iv, bool has_interval) { /* BuildEmptyIntervalRec record.go:1327-1338 (ascending) */
    size_t ncol = schema.size() - 1;
    uint32_t num = has_interval ? (uint32_t)((max - min) / iv) : 1;
    n_rows = num; start = has_interval ? min : 0; interval = iv; this->has_interval = has_interval;
    values.assign(ncol, std::vector<uint64_t>(num, 0));
    valid.assign(ncol, std::vector<uint8_t>(num, 0));
    col_times.assign(ncol, std::vector<int64_t>(num, 0));
    times.resize(num);
    for (uint32_t i = 0; i < num; i++) times[i] = has_interval ? min + iv * (int64_t)i : 0;
}

namespace {
inline bool f_tie_le(int64_t t1, int64_t t2) { /* reccord_functions.go:487-488: the time column is read with FloatValue() */
    double a, b; memcpy(&a, &t1, 8); memcpy(&b, &t2, 8); return a <= b;
}
}

void IntervalRecord::update_from(const Record &rec) { /* RecordInit :1069-1093 + UpdateRec :1111-1120 */
    const int64_t *rt = rec.times();
    int nrows = rec.row_nums();
    size_t ncol = schema.size() - 1;
    for (int rr = 0; rr < nrows; rr++) {
        int64_t t = rt[rr];
        uint32_t row = 0;
        if (has_interval) { int64_t d = t - start; if (d < 0) d = -d; row = (uint32_t)(d / interval); } /* GetIndex :1030-1038 */
        if (row >= n_rows) continue; /* the reference would index out of range; never happens for in-range partials */
        for (size_t ci = 0; ci < exprs.size(); ci++) {
            const ExprOpt &e = exprs[ci];
            int ocol = -1;
            for (size_t k = 0; k < ncol; k++) if (schema[k].name == e.out_name) { ocol = (int)k; break; }
            int rcol = rec.field_index(schema[ocol].name);
            if (ocol < 0 || rcol < 0) continue;
            const ColVal &cv = rec.cols[rcol];
            if (cv.is_nil(rr)) continue; /* every Update* returns on a nil partial */
            int type = schema[ocol].type;
            uint64_t v;
            {
                int vi = cv.nil_count == 0 ? rr : cv.valid_count(0, rr);
                if (type == OG_TYPE_BOOL) v = cv.val[vi]; else memcpy(&v, cv.val.data() + 8 * vi, 8);
            }
            uint64_t &sv = values[ocol][row];
            bool src_nil = !valid[ocol][row];
            auto replace_row = [&]() { /* UpdateIntervalRecRow record.go:1360-1364: every column incl. time from the partial's row */
                for (size_t k = 0; k < ncol; k++) {
                    int rc2 = rec.field_index(schema[k].name);
                    if (rc2 < 0) continue;
                    const ColVal &c2 = rec.cols[rc2];
                    if (c2.is_nil(rr)) { if (valid[k][row]) { values[k][row] = 0; valid[k][row] = 0; } continue; }
                    int vi2 = c2.nil_count == 0 ? rr : c2.valid_count(0, rr);
                    uint64_t x; if (schema[k].type == OG_TYPE_BOOL) x = c2.val[vi2]; else memcpy(&x, c2.val.data() + 8 * vi2, 8);
                    values[k][row] = x; valid[k][row] = 1;
                    if (!rec.meta_times[rc2].empty()) col_times[k][row] = rec.meta_times[rc2][rr]; /* updateRecMeta */
                }
                times[row] = t;
            };
            auto lt = [&](uint64_t a, uint64_t b) { return type == OG_TYPE_FLOAT ? as_f(a) < as_f(b) : type == OG_TYPE_INT ? (int64_t)a < (int64_t)b : (a != b && !a); };
            auto gt = [&](uint64_t a, uint64_t b) { return type == OG_TYPE_FLOAT ? as_f(a) > as_f(b) : type == OG_TYPE_INT ? (int64_t)a > (int64_t)b : (a != b && a != 0); };
            auto eq = [&](uint64_t a, uint64_t b) { return type == OG_TYPE_FLOAT ? as_f(a) == as_f(b) : a == b; };
            auto ge = [&](uint64_t a, uint64_t b) { return type == OG_TYPE_FLOAT ? as_f(a) >= as_f(b) : type == OG_TYPE_INT ? (int64_t)a >= (int64_t)b : (a == b || a != 0); };
            auto le = [&](uint64_t a, uint64_t b) { return type == OG_TYPE_FLOAT ? as_f(a) <= as_f(b) : type == OG_TYPE_INT ? (int64_t)a <= (int64_t)b : (a == b || a == 0); };
            switch (e.func) {
            case OG_AGG_COUNT: /* updateCountImpl :757-760 */
                sv = (uint64_t)((int64_t)sv + (int64_t)v); valid[ocol][row] = 1; break;
            case OG_AGG_SUM: /* updateIntegerSumImpl :712-715 / updateFloatSumImpl :730-733 */
                if (type == OG_TYPE_FLOAT) sv = f_bits(as_f(v) + as_f(sv)); else sv = sv + v;
                valid[ocol][row] = 1; break;
            case OG_AGG_MIN:
            case OG_AGG_MAX: {
                bool is_min = e.func == OG_AGG_MIN;
                if (multi) { /* update*Column{Min,Max}Impl :586-660: value only, <= / >= keeps */
                    if ((is_min ? le(sv, v) : ge(sv, v)) && !src_nil) break;
                    sv = v; valid[ocol][row] = 1;
                } else { /* update*{Min,Max}Impl :429-560: strict compare, tie -> earlier time, then whole row */
                    if ((is_min ? lt(sv, v) : gt(sv, v)) && !src_nil) break;
                    int64_t t1 = times[row], t2 = t;
                    bool t_le = type == OG_TYPE_FLOAT ? f_tie_le(t1, t2) : t1 <= t2;
                    if (eq(sv, v) && t_le && !src_nil) break;
                    replace_row();
                }
                break;
            }
            case OG_AGG_FIRST:
            case OG_AGG_LAST: {
                bool is_first = e.func == OG_AGG_FIRST;
                auto cmp = [&](int64_t a, int64_t b) { return is_first ? a > b : a < b; };
                if (multi) { /* update*ColumnFirstLastImp :229-420 */
                    int64_t t1 = col_times[ocol][row], t2 = rec.meta_times[rcol].empty() ? t : rec.meta_times[rcol][rr];
                    if (cmp(t1, t2)) { col_times[ocol][row] = t2; sv = v; valid[ocol][row] = 1; break; }
                    if (!src_nil && cmp(t2, t1)) break;
                    if (ge(sv, v) && !src_nil) break;
                    sv = v; valid[ocol][row] = 1; col_times[ocol][row] = t2;
                } else { /* update*FirstLastImp :47-227 */
                    int64_t t1 = times[row], t2 = t;
                    if (cmp(t1, t2)) { replace_row(); break; }
                    if (!src_nil && cmp(t2, t1)) break;
                    if (ge(sv, v) && !src_nil) break;
                    replace_row();
                }
                break;
            }
            }
        }
    }
}

} // namespace ogo
