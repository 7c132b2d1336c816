/*
 * tssp.cpp — the TSSP container, read side: file bytes -> og_shard_desc (the flattened ChunkMeta directory og_shard_open
 * takes).  Host code only: the container is a few bytes of metadata per segment, walked once per file; the pages it points
 * at are what the GPU reads.
 *
 * What the reference does on this path (engine/immutable):
 *   file     = "53ac2021" | u64 BE version=2 | chunks | chunk-meta blocks | meta index | bloom | id-time | trailer | footer
 *              (msbuilder.go:1355-1425 Flush, table.go:24-28)
 *   footer   = i64 (zig-zag, BE) trailer offset                                  (msbuilder.go:1413-1415)
 *   trailer  = 6 x i64zz {dataOffset,dataSize,indexSize,metaIndexSize,bloomSize,idTimeSize} + TableStat
 *              (trailer.go:59-88, table_stat.go:36-84) with the ExtraData flags (table_stat.go:122-207)
 *   meta idx = metaIndexItemNum x {u64 id, i64zz minT, i64zz maxT, i64zz offset, u32 count, u32 size}  (tssp_file_meta.go:769-802)
 *   block    = count ChunkMetas back to back, then count x u32 BE start offsets   (msbuilder.go:1481-1500, tssp_file.go:606-658)
 *   chunk    = u64 sid, i64zz offset, u32 size, u32 columnCount, u32 segCount, segCount x (i64zz min, i64zz max),
 *              then per column u16 nameLen, name, u8 type, u16 preAggLen, preAgg, segCount x (i64zz offset, u32 size)
 *              (tssp_file_meta.go:566-581, 228-246, 86-104, 129-143); columns sorted by name, time last
 *   a chunk's bytes = per column [u32 crc][pages of its segments]; Segment.offset is the absolute file offset of a page
 *              (chunkdata_builder_ts.go:37-82)
 * Integers of type int64 are zig-zag coded before the big-endian store (lib/numberenc/number.go:155-168).
 *
 * Not handled (refused with OG_E_UNSUPPORTED): compressed chunk metas (ChunkMetaCompressFlag != 0, chunk_meta_codec.go), detached
 * (object-store) files.  The per-column CRC32 is not verified: pages are validated structurally on the device at og_shard_open.
 */
#include <algorithm>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/ogpu.h"

namespace ogpu { void set_error(const char *fmt, ...); }
using ogpu::set_error;

namespace {

struct Rd { /* bounds-checked big-endian reader over [p, end) */
    const uint8_t *p, *end; bool ok = true;
    Rd(const uint8_t *b, uint64_t n) : p(b), end(b + n) {}
    uint64_t left() const { return (uint64_t)(end - p); }
    bool need(uint64_t n) { if (!ok || left() < n) { ok = false; return false; } return true; }
    uint64_t u64() { if (!need(8)) return 0; uint64_t v = 0; for (int i = 0; i < 8; i++) v = (v << 8) | p[i]; p += 8; return v; }
    int64_t i64() { const uint64_t u = u64(); return (int64_t)(u >> 1) ^ -(int64_t)(u & 1); }
    uint32_t u32() { if (!need(4)) return 0; uint32_t v = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; p += 4; return v; }
    uint32_t u16() { if (!need(2)) return 0; uint32_t v = ((uint32_t)p[0] << 8) | p[1]; p += 2; return v; }
    uint32_t u8() { if (!need(1)) return 0; return *p++; }
    const uint8_t *bytes(uint64_t n) { if (!need(n)) return nullptr; const uint8_t *r = p; p += n; return r; }
};

struct Col {
    std::string name; int type;
    std::vector<uint64_t> off; std::vector<uint32_t> len;
};

} // namespace

struct og_tssp {
    const uint8_t *file; uint64_t len;
    std::string measurement;
    int64_t min_time = 0, max_time = 0, id_count = 0; uint64_t min_id = 0, max_id = 0;
    std::vector<uint64_t> sids; std::vector<uint32_t> seg_begin;
    std::vector<int64_t> tmin, tmax;
    std::vector<uint64_t> time_off; std::vector<uint32_t> time_len;
    std::vector<Col> cols;
    std::vector<og_column_desc> col_desc;
};

extern "C" {

OG_API int og_tssp_parse(const uint8_t *file, uint64_t len, og_tssp **out) {
    if (!file || !out) { set_error("null argument"); return OG_E_INVAL; }
    *out = nullptr;
    static const char magic[] = "53ac2021";
    const uint64_t header = 16, footer = 8;
    if (len < header + footer || memcmp(file, magic, 8) != 0) { set_error("not a TSSP file (magic)"); return OG_E_CORRUPT; }
    { Rd r(file + 8, 8); const uint64_t v = r.u64(); if (v != 2) { set_error("TSSP version %llu (this reader knows 2)", (unsigned long long)v); return OG_E_UNSUPPORTED; } }
    int64_t toff; { Rd r(file + len - footer, footer); toff = r.i64(); }
    if (toff < (int64_t)header || (uint64_t)toff > len - footer) { set_error("trailer offset %lld outside the file", (long long)toff); return OG_E_CORRUPT; }

    /* ---- trailer ---- */
    Rd t(file + toff, len - footer - (uint64_t)toff);
    const int64_t data_off = t.i64(), data_size = t.i64(), index_size = t.i64(), mi_size = t.i64(), bloom_size = t.i64(), idtime_size = t.i64();
    std::unique_ptr<og_tssp> f(new og_tssp());
    f->file = file; f->len = len;
    f->id_count = t.i64(); f->min_id = t.u64(); f->max_id = t.u64(); f->min_time = t.i64(); f->max_time = t.i64();
    const int64_t mi_items = t.i64();
    (void)t.u64(); (void)t.u64(); /* bloomM, bloomK */
    uint32_t dlen = t.u16();
    if (!t.ok) { set_error("trailer truncated"); return OG_E_CORRUPT; }
    { /* ExtraData: 1 B / 2 B legacy forms, else 8 B LE flags with the real length in the upper 32 bits (table_stat.go:177-207) */
        if (t.left() < dlen) { set_error("trailer extra data truncated"); return OG_E_CORRUPT; }
        uint32_t compress = 0; uint64_t real = dlen;
        if (dlen == 2) compress = t.p[1];
        else if (dlen >= 8) {
            uint64_t fl = 0; for (int i = 7; i >= 0; i--) fl = (fl << 8) | t.p[i];
            compress = (uint32_t)(fl >> 8) & 0xff;
            if ((fl >> 32) != 0) real = fl >> 32;
        }
        if (compress != 0) { set_error("chunk metas are compressed (mode %u): not supported by this reader", compress); return OG_E_UNSUPPORTED; }
        if (!t.bytes(real)) { set_error("trailer extra data truncated"); return OG_E_CORRUPT; }
    }
    { const uint32_t nl = t.u16(); const uint8_t *nm = t.bytes(nl); if (!nm) { set_error("trailer name truncated"); return OG_E_CORRUPT; } f->measurement.assign((const char *)nm, nl); }
    if (data_off != (int64_t)header || data_size < 0 || index_size < 0 || mi_size < 0 || bloom_size < 0 || idtime_size < 0 || mi_items < 0 ||
        (uint64_t)data_off + (uint64_t)data_size + (uint64_t)index_size + (uint64_t)mi_size + (uint64_t)bloom_size + (uint64_t)idtime_size != (uint64_t)toff) {
        set_error("trailer section sizes do not add up to the trailer offset"); return OG_E_CORRUPT;
    }
    const uint64_t index_off = (uint64_t)data_off + (uint64_t)data_size, mi_off = index_off + (uint64_t)index_size;
    if ((uint64_t)mi_items * 40 != (uint64_t)mi_size) { set_error("meta index: %lld items do not fill %lld bytes", (long long)mi_items, (long long)mi_size); return OG_E_CORRUPT; }

    /* ---- meta index -> chunk-meta blocks -> chunk metas ---- */
    std::map<std::string, size_t> col_of;
    Rd mi(file + mi_off, (uint64_t)mi_size);
    uint64_t prev_sid = 0;
    for (int64_t b = 0; b < mi_items; b++) {
        (void)mi.u64(); (void)mi.i64(); (void)mi.i64();
        const int64_t boff = mi.i64(); const uint32_t count = mi.u32(), bsize = mi.u32();
        if (boff < (int64_t)index_off || (uint64_t)boff + bsize > mi_off || (uint64_t)count * 4 >= bsize) { set_error("meta index item %lld points outside the chunk-meta region", (long long)b); return OG_E_CORRUPT; }
        Rd cm(file + boff, bsize - (uint64_t)count * 4);
        for (uint32_t i = 0; i < count; i++) {
            const uint64_t sid = cm.u64();
            (void)cm.i64(); (void)cm.u32();
            const uint32_t ncol = cm.u32(), nseg = cm.u32();
            if (!cm.ok || ncol == 0 || (uint64_t)nseg * 16 > cm.left()) { set_error("chunk meta %u of block %lld is truncated", i, (long long)b); return OG_E_CORRUPT; }
            if (sid == 0 || (!f->sids.empty() && sid <= prev_sid)) { set_error("chunk metas are not in ascending series-id order (sid %llu)", (unsigned long long)sid); return OG_E_CORRUPT; }
            prev_sid = sid;
            const size_t seg0 = f->tmin.size();
            if (seg0 + nseg > 0xffffffffull) { set_error("more than 2^32 segments"); return OG_E_UNSUPPORTED; }
            f->sids.push_back(sid); f->seg_begin.push_back((uint32_t)seg0);
            for (uint32_t s = 0; s < nseg; s++) { f->tmin.push_back(cm.i64()); f->tmax.push_back(cm.i64()); }
            f->time_off.resize(seg0 + nseg, 0); f->time_len.resize(seg0 + nseg, 0);
            for (Col &c : f->cols) { c.off.resize(seg0 + nseg, 0); c.len.resize(seg0 + nseg, 0); }
            for (uint32_t c = 0; c < ncol; c++) {
                const uint32_t nl = cm.u16(); const uint8_t *nm = cm.bytes(nl);
                const int ty = (int)cm.u8(); const uint32_t pl = cm.u16();
                if (!cm.bytes(pl) || (uint64_t)nseg * 12 > cm.left()) { set_error("column meta %u of series %llu is truncated", c, (unsigned long long)sid); return OG_E_CORRUPT; }
                const std::string name((const char *)nm, nl);
                const bool is_time = c + 1 == ncol;
                if (is_time != (name == "time") || (is_time && ty != OG_TYPE_INT)) { set_error("series %llu: the time column must be the last column", (unsigned long long)sid); return OG_E_CORRUPT; }
                uint64_t *off; uint32_t *ln;
                if (is_time) { off = f->time_off.data() + seg0; ln = f->time_len.data() + seg0; }
                else {
                    auto it = col_of.find(name);
                    if (it == col_of.end()) {
                        it = col_of.emplace(name, f->cols.size()).first;
                        Col nc; nc.name = name; nc.type = ty; nc.off.assign(seg0 + nseg, 0); nc.len.assign(seg0 + nseg, 0);
                        f->cols.push_back(std::move(nc));
                    }
                    Col &col = f->cols[it->second];
                    if (col.type != ty) { set_error("column %s changes type inside the file (%d, %d)", name.c_str(), col.type, ty); return OG_E_TYPE; }
                    off = col.off.data() + seg0; ln = col.len.data() + seg0;
                }
                for (uint32_t s = 0; s < nseg; s++) {
                    const int64_t o = cm.i64(); const uint32_t z = cm.u32();
                    if (o < data_off || (uint64_t)o + z > index_off || z == 0) { set_error("series %llu column %s segment %u lies outside the data region", (unsigned long long)sid, name.c_str(), s); return OG_E_CORRUPT; }
                    off[s] = (uint64_t)o; ln[s] = z;
                }
            }
            if (!cm.ok) { set_error("chunk meta of series %llu is truncated", (unsigned long long)sid); return OG_E_CORRUPT; }
        }
        if (cm.left() != 0) { set_error("chunk-meta block %lld: %llu stray bytes", (long long)b, (unsigned long long)cm.left()); return OG_E_CORRUPT; }
    }
    if (!mi.ok) { set_error("meta index truncated"); return OG_E_CORRUPT; }
    f->seg_begin.push_back((uint32_t)f->tmin.size());
    /* schema order: sorted by name (lib/record/record.go:115-123) */
    std::sort(f->cols.begin(), f->cols.end(), [](const Col &a, const Col &b) { return a.name < b.name; });
    for (const Col &c : f->cols) {
        if (c.type != OG_TYPE_INT && c.type != OG_TYPE_FLOAT && c.type != OG_TYPE_BOOL && c.type != OG_TYPE_STRING) { set_error("column %s has type %d", c.name.c_str(), c.type); return OG_E_UNSUPPORTED; }
        og_column_desc d; d.name = c.name.c_str(); d.type = c.type; d.page_off = c.off.data(); d.page_len = c.len.data();
        f->col_desc.push_back(d);
    }
    *out = f.release();
    return OG_OK;
}

OG_API int og_tssp_desc(const og_tssp *f, og_shard_desc *d) {
    if (!f || !d) { set_error("null argument"); return OG_E_INVAL; }
    memset(d, 0, sizeof *d);
    d->data = f->file; d->data_len = f->len;
    d->n_series = (uint32_t)f->sids.size(); d->sids = f->sids.data(); d->series_seg_begin = f->seg_begin.data();
    d->n_segments = (uint32_t)f->tmin.size(); d->seg_tmin = f->tmin.data(); d->seg_tmax = f->tmax.data();
    d->n_columns = (uint32_t)f->col_desc.size(); d->columns = f->col_desc.data();
    d->time_page_off = f->time_off.data(); d->time_page_len = f->time_len.data();
    d->flags = 0;
    return OG_OK;
}

OG_API const char *og_tssp_measurement(const og_tssp *f) { return f ? f->measurement.c_str() : ""; }

OG_API int og_tssp_time_range(const og_tssp *f, int64_t *min_time, int64_t *max_time) {
    if (!f || !min_time || !max_time) { set_error("null argument"); return OG_E_INVAL; }
    *min_time = f->min_time; *max_time = f->max_time;
    return OG_OK;
}

OG_API void og_tssp_free(og_tssp *f) { delete f; }

} // extern "C"
