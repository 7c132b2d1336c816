/* ogpu::GpuAggCursor — see gpu_cursor.h for the reference interfaces this mirrors. Only C-ABI calls below this line. */
#include "gpu_cursor.h"

namespace ogpu {

static Error make_err(int st, const char *where) {
    Error e; e.code = st;
    if (st != OG_OK) { e.msg = std::string(where) + ": " + og_strerror(st) + " " + og_last_error(); }
    return e;
}

Error GpuShard::Open(const og_shard_desc &desc, std::shared_ptr<GpuShard> *out) {
    auto sh = std::shared_ptr<GpuShard>(new GpuShard());
    int st = og_shard_open(&desc, &sh->s_);
    if (st != OG_OK) return make_err(st, "og_shard_open");
    for (uint32_t c = 0; c < desc.n_columns; c++) sh->schema_.push_back(Field{desc.columns[c].type, desc.columns[c].name ? desc.columns[c].name : ""});
    sh->schema_.push_back(Field{Field_Type_Int, TimeField});
    *out = sh;
    return Error{};
}

Error GpuShard::Synth(const og_synth_desc &desc, std::shared_ptr<GpuShard> *out) {
    auto sh = std::shared_ptr<GpuShard>(new GpuShard());
    int st = og_shard_synth(&desc, &sh->s_);
    if (st != OG_OK) return make_err(st, "og_shard_synth");
    for (uint32_t c = 0; c < desc.n_columns; c++) sh->schema_.push_back(Field{desc.columns[c].type, "f" + std::to_string(c)});
    sh->schema_.push_back(Field{Field_Type_Int, TimeField});
    *out = sh;
    return Error{};
}

GpuShard::~GpuShard() { if (s_) og_shard_close(s_); }

GpuAggCursor::GpuAggCursor(std::shared_ptr<GpuShard> shard) : shard_(std::move(shard)) {}
GpuAggCursor::~GpuAggCursor() { (void)Close(); }

void GpuAggCursor::SetOps(const std::vector<CallOption> &ops) { ops_ = ops; }

static int func_of(const std::string &name) {
    if (name == "count") return OG_AGG_COUNT;
    if (name == "sum") return OG_AGG_SUM;
    if (name == "min") return OG_AGG_MIN;
    if (name == "max") return OG_AGG_MAX;
    if (name == "first") return OG_AGG_FIRST;
    if (name == "last") return OG_AGG_LAST;
    return 0; /* newProcessor panics on an unknown call (series_call_processor.go:80); here: OG_E_UNSUPPORTED */
}

Error GpuAggCursor::SinkPlan(const QueryPlan &plan) {
    if (closed_) return make_err(OG_E_STATE, "SinkPlan on a closed cursor");
    if (q_) { og_query_destroy(q_); q_ = nullptr; ran_ = false; }
    const Schemas &in = shard_->schema();
    auto col_of = [&](const std::string &name) -> int {
        for (size_t i = 0; i + 1 < in.size(); i++) if (in[i].Name == name) return (int)i;
        return -1;
    };
    std::vector<og_call> calls;
    out_schema_.clear();
    for (const CallOption &op : ops_) {
        int f = func_of(op.Call), c = col_of(op.Ref);
        if (!f) { Error e; e.code = OG_E_UNSUPPORTED; e.msg = "unsupported aggregate call: " + op.Call; return e; }
        if (c < 0) { Error e; e.code = OG_E_INVAL; e.msg = "unknown field: " + op.Ref; return e; }
        calls.push_back(og_call{f, c});
        /* output column type: count → integer, everything else keeps the field type (series_call_processor.go:87-283) */
        out_schema_.push_back(Field{f == OG_AGG_COUNT ? (int)Field_Type_Int : in[c].Type, op.Call + "_" + op.Ref});
    }
    out_schema_.push_back(Field{Field_Type_Int, TimeField});
    std::vector<og_filter_item> filt;
    for (const CondItem &ci : plan.Condition) {
        og_filter_item it{};
        it.kind = ci.kind;
        if (ci.kind == OG_F_TERM) {
            it.column = col_of(ci.field);
            if (it.column < 0) { Error e; e.code = OG_E_INVAL; e.msg = "unknown field in condition: " + ci.field; return e; }
            it.op = ci.op; it.const_is_float = ci.is_float; it.fval = ci.f; it.ival = ci.i;
        }
        filt.push_back(it);
    }
    og_query_desc d{};
    d.interval = plan.Interval; d.offset = plan.Offset; d.tmin = plan.StartTime; d.tmax = plan.EndTime;
    d.ascending = plan.Ascending ? 1 : 0;
    d.n_calls = (uint32_t)calls.size(); d.calls = calls.data();
    d.n_filter = (uint32_t)filt.size(); d.filter = filt.data();
    d.group_mode = plan.GroupBy == QueryPlan::GroupAll ? OG_GROUP_ALL : plan.GroupBy == QueryPlan::GroupBySeries ? OG_GROUP_PER_SERIES : OG_GROUP_MAP;
    d.n_groups = plan.NumGroups; d.series_group = plan.SeriesGroup.empty() ? nullptr : plan.SeriesGroup.data();
    d.chunk_size = plan.ChunkSize;
    d.flags = plan.StrictOrder ? OG_Q_STRICT_ORDER : 0;
    return make_err(og_query_create(shard_->handle(), &d, &q_), "og_query_create");
}

Error GpuAggCursor::run_once() {
    if (closed_ || !q_) return make_err(OG_E_STATE, "Next before SinkPlan / after Close");
    if (ran_) return Error{};
    int st = og_query_run(q_);
    if (st != OG_OK) return make_err(st, "og_query_run");
    ran_ = true;
    og_query_stats(q_, &stats_);
    return Error{};
}

Error GpuAggCursor::Next(const Record **rec, const SeriesInfo **info) {
    *rec = nullptr; if (info) *info = nullptr;
    if (Error e = run_once()) return e;
    og_record_view v{};
    int st = og_query_next(q_, &v);
    if (st == OG_EOF) return Error{};            /* (nil, nil, nil) */
    if (st != OG_OK) return make_err(st, "og_query_next");
    Record &r = rec_[ring_]; ring_ ^= 1;
    r.Schema = out_schema_;
    r.ColVals.assign(v.n_cols + 1, ColVal{});
    r.Meta.Times.assign(v.n_cols, nullptr);
    for (uint32_t c = 0; c < v.n_cols; c++) {
        const og_colval_view &cv = v.cols[c];
        ColVal &o = r.ColVals[c];
        o.Val = cv.val; o.ValBytes = cv.val_bytes; o.Bitmap = cv.bitmap; o.BitMapOffset = cv.bitmap_offset;
        o.Len = cv.len; o.NilCount = cv.nil_count;
        r.Meta.Times[c] = cv.times;
    }
    ColVal &t = r.ColVals[v.n_cols];
    t.Val = reinterpret_cast<const uint8_t *>(v.times); t.ValBytes = (size_t)v.rows * 8; t.Len = v.rows; t.NilCount = 0;
    info_.sid = v.sid; info_.group = v.group;
    *rec = &r; if (info) *info = &info_;
    return Error{};
}

Error GpuAggCursor::NextAggData(const Record **rec, const FileInfo **info) {
    const SeriesInfo *si = nullptr;
    Error e = Next(rec, &si);
    if (info) *info = nullptr;
    if (e || !*rec) return e;
    finfo_.Info = *si;
    finfo_.MinTime = (*rec)->Times()[0];
    finfo_.MaxTime = (*rec)->Times()[(*rec)->RowNums() - 1];
    if (info) *info = &finfo_;
    return e;
}

void GpuAggCursor::EndSpan() { span_ = false; }

Error GpuAggCursor::Close() {
    if (closed_) return Error{};
    closed_ = true;
    if (q_) { og_query_abort(q_); og_query_destroy(q_); q_ = nullptr; }
    shard_.reset(); /* tsspFile.Unref */
    return Error{};
}

} // namespace ogpu
