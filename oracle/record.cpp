/*
 * record.cpp — CPU ORACLE (test infrastructure, see og_oracle.h): ColVal / Record model and the per-segment page
 * framing of engine/immutable (column header + block).
 */
#include <cstring>

#include "og_oracle.h"

namespace ogo {

/* ===================== lib/record/column.go ===================== */
bool ColVal::is_nil(int i) const { /* IsNil */
    if (nil_count == 0) return false;
    if (bitmap.empty() || len == nil_count) return true;
    int idx = bitmap_offset + i;
    return (bitmap[idx >> 3] & (1 << (idx & 7))) == 0;
}

int ColVal::valid_count(int start, int end) const { /* ValidCount :297-314 */
    if ((int)val.size() + nil_count == 0 || len == nil_count) return 0;
    if (nil_count == 0) return end - start;
    int c = 0;
    for (int i = start + bitmap_offset; i < end + bitmap_offset; i++)
        if (bitmap[i >> 3] & (1 << (i & 7))) c++;
    return c;
}

void ColVal::value_index_range(int bm_start, int bm_end, int *s, int *e) const { /* getValIndexRange :453-458 */
    if (nil_count == 0) { *s = bm_start; *e = bm_end; return; }
    int a = 0;
    for (int i = bitmap_offset; i < bm_start + bitmap_offset; i++) if (bitmap[i >> 3] & (1 << (i & 7))) a++;
    int b = a;
    for (int i = bm_start + bitmap_offset; i < bm_end + bitmap_offset; i++) if (bitmap[i >> 3] & (1 << (i & 7))) b++;
    *s = a; *e = b;
}

void ColVal::append_bit(bool present) { /* setBitMap / resetBitMap :489-506 */
    int idx = bitmap_offset + len;
    if ((idx >> 3) >= (int)bitmap.size()) bitmap.push_back(0);
    if (present) bitmap[idx >> 3] |= (uint8_t)(1 << (idx & 7));
    else bitmap[idx >> 3] &= (uint8_t)~(1 << (idx & 7));
    len++;
}
void ColVal::append_integer(int64_t v) { const uint8_t *p = (const uint8_t *)&v; val.insert(val.end(), p, p + 8); append_bit(true); }
void ColVal::append_float(double v) { const uint8_t *p = (const uint8_t *)&v; val.insert(val.end(), p, p + 8); append_bit(true); }
void ColVal::append_boolean(bool v) { val.push_back(v ? 1 : 0); append_bit(true); }
void ColVal::append_null(int type, bool reserve) {
    if (reserve) { /* AppendXxxNullReserve: a value slot is kept (record.go:1298-1338 interval records) */
        if (type == OG_TYPE_BOOL) val.push_back(0);
        else val.insert(val.end(), 8, 0);
    }
    append_bit(false);
    nil_count++;
}

int Record::field_index(const std::string &n) const {
    for (size_t i = 0; i < schema.size(); i++) if (schema[i].name == n) return (int)i;
    return -1;
}
void Record::reset() {
    for (auto &c : cols) c.init();
    for (auto &t : meta_times) t.clear();
}

/* ===================== engine/immutable/column_builder.go ===================== */
enum { BLOCK_ONE_BEGIN = 16, BLOCK_ONE_END = 21, BLOCK_FULL_BEGIN = 30, BLOCK_FULL_END = 35, BLOCK_EMPTY_BEGIN = 40, BLOCK_EMPTY_END = 45 };
static inline uint8_t type_one(int t) { return t == OG_TYPE_FLOAT ? 17 : t == OG_TYPE_INT ? 18 : t == OG_TYPE_BOOL ? 19 : 20; }   /* encoding.go:40-46 */
static inline uint8_t type_full(int t) { return t == OG_TYPE_FLOAT ? 31 : t == OG_TYPE_INT ? 32 : t == OG_TYPE_BOOL ? 33 : 34; }  /* :48-53 */
static inline uint8_t type_empty(int t) { return t == OG_TYPE_FLOAT ? 41 : t == OG_TYPE_INT ? 42 : t == OG_TYPE_BOOL ? 43 : 44; } /* :55-60 */

/* subBitmapBytes lib/record/record.go:960-966 */
static void sub_bitmap_bytes(const ColVal &c, const uint8_t **p, size_t *n, int *off) {
    int bo = c.bitmap_offset, l = c.len;
    size_t a = (size_t)(bo >> 3);
    size_t b = (size_t)((bo + l) >> 3) + (((bo + l) & 7) ? 1 : 0);
    *p = c.bitmap.data() + a; *n = b - a; *off = bo & 7;
}

static void encode_column_header(const ColVal &col, int type, Bytes &dst) { /* EncodeColumnHeader :428-444 */
    if (col.nil_count == 0) { dst.push_back(type_full(type)); put_u32be(dst, (uint32_t)col.len); return; }
    if (col.nil_count == col.len) { dst.push_back(type_empty(type)); put_u32be(dst, (uint32_t)col.len); return; }
    dst.push_back((uint8_t)type);
    const uint8_t *bm; size_t n; int off;
    sub_bitmap_bytes(col, &bm, &n, &off);
    put_u32be(dst, (uint32_t)n);
    dst.insert(dst.end(), bm, bm + n);
    put_u32be(dst, (uint32_t)off);
    put_u32be(dst, (uint32_t)col.nil_count);
}

int encode_field_page(const ColVal &col, int type, Bytes &out) { /* enc{Integer,Float,Boolean}Column :151-349 */
    if (col.len == 1 && col.val.size() < 16 && col.val.size() > 0) { /* CanEncodeOneRowMode :488 */
        out.push_back(type_one(type));
        out.insert(out.end(), col.val.begin(), col.val.end());
        return E_OK;
    }
    encode_column_header(col, type, out);
    switch (type) {
    case OG_TYPE_FLOAT: return float_block_encode(col.floats(), col.val.size() / 8, out);
    case OG_TYPE_INT: return int_block_encode(col.integers(), col.val.size() / 8, out);
    case OG_TYPE_BOOL: return bool_block_encode(col.booleans(), col.val.size(), out);
    default: return E_UNSUPPORTED;
    }
}

int encode_time_page(const int64_t *t, size_t n, Bytes &out) { /* ChunkDataBuilder.EncodeTime chunkdata_builder.go:65-114 */
    if (n == 1) { /* CanEncodeOneRowMode: Len==1, len(Val)=8 */
        out.push_back(18);
        const uint8_t *p = (const uint8_t *)t;
        out.insert(out.end(), p, p + 8);
        return E_OK;
    }
    out.push_back(32); /* BlockIntegerFull: time columns have no nulls */
    put_u32be(out, (uint32_t)n);
    return time_block_encode(t, n, out);
}

/* DecodeColumnHeader :446-486. Returns the block slice and the bitmap slice. */
static int decode_column_header(const uint8_t *data, size_t len, int col_type, ColVal &col, const uint8_t **block,
                                size_t *block_len, const uint8_t **bm, size_t *bm_len) {
    if (len < 1) return E_CORRUPT;
    uint8_t typ = data[0];
    if (typ > BLOCK_FULL_BEGIN && typ < BLOCK_FULL_END) {
        if (len < 5) return E_CORRUPT;
        col.len = (int)get_u32be(data + 1);
        col.nil_count = 0; col.bitmap_offset = 0;
        col.bitmap.assign((size_t)(col.len + 7) / 8, 0xff); /* FillBitmap(255) + RepairBitmap */
        if (col.len & 7) col.bitmap.back() = (uint8_t)((1 << (col.len & 7)) - 1);
        *block = data + 5; *block_len = len - 5; *bm = col.bitmap.data(); *bm_len = col.bitmap.size();
        return E_OK;
    }
    if (typ > BLOCK_EMPTY_BEGIN && typ < BLOCK_EMPTY_END) {
        if (len < 5) return E_CORRUPT;
        col.len = (int)get_u32be(data + 1);
        col.nil_count = col.len; col.bitmap_offset = 0;
        col.bitmap.assign((size_t)(col.len + 7) / 8, 0);
        *block = data + 5; *block_len = len - 5; *bm = col.bitmap.data(); *bm_len = col.bitmap.size();
        return E_OK;
    }
    if (typ != (uint8_t)col_type) return OG_E_TYPE; /* "type(%v) in table not eq select type(%v)" */
    if (len < 5) return E_CORRUPT;
    size_t pos = 1;
    size_t nb = get_u32be(data + pos);
    if (len - pos < nb + 8) return E_CORRUPT;
    pos += 4;
    *bm = data + pos; *bm_len = nb;
    pos += nb;
    col.bitmap_offset = (int)get_u32be(data + pos); pos += 4;
    col.nil_count = (int)get_u32be(data + pos); pos += 4;
    *block = data + pos; *block_len = len - pos;
    return E_OK;
}

int decode_field_page(const uint8_t *p, size_t len, int type, ColVal &col) { /* decodeColumnData reader.go:674-698 */
    col.init();
    if (len < 1) return E_CORRUPT;
    if (p[0] > BLOCK_ONE_BEGIN && p[0] < BLOCK_ONE_END) { /* DecodeColumnOfOneValue :700-720 */
        col.len = 1; col.nil_count = 0; col.bitmap.assign(1, 0);
        if (len == 1) { col.nil_count = 1; }
        else { col.val.assign(p + 1, p + len); col.bitmap[0] = 1; }
        return E_OK;
    }
    const uint8_t *block, *bm; size_t bl, bml;
    ColVal hdr;
    int rc = decode_column_header(p, len, type, hdr, &block, &bl, &bm, &bml);
    if (rc != E_OK) return rc;
    int nil_count = hdr.nil_count, bm_off = hdr.bitmap_offset;
    /* append{Integer,Float,Boolean}Column reader.go:504-579 */
    if (bl != 0) {
        size_t nvals;
        if (type == OG_TYPE_FLOAT) {
            std::vector<double> v; rc = float_block_decode(block, bl, v); if (rc != E_OK) return rc;
            col.val.resize(v.size() * 8); memcpy(col.val.data(), v.data(), v.size() * 8); nvals = v.size();
        } else if (type == OG_TYPE_INT) {
            std::vector<int64_t> v; rc = int_block_decode(block, bl, v); if (rc != E_OK) return rc;
            col.val.resize(v.size() * 8); memcpy(col.val.data(), v.data(), v.size() * 8); nvals = v.size();
        } else if (type == OG_TYPE_BOOL) {
            std::vector<uint8_t> v; rc = bool_block_decode(block, bl, v); if (rc != E_OK) return rc;
            col.val = v; nvals = v.size();
        } else return E_UNSUPPORTED;
        int rows = (int)nvals + nil_count;
        /* AppendBitmap(nilBitmap, bitmapOffset, rows, 0, rows): re-pack at offset 0 */
        col.bitmap.assign((size_t)(rows + 7) / 8, 0);
        for (int i = 0; i < rows; i++) {
            int si = bm_off + i;
            if ((size_t)(si >> 3) >= bml) return E_CORRUPT;
            if (bm[si >> 3] & (1 << (si & 7))) col.bitmap[i >> 3] |= (uint8_t)(1 << (i & 7));
        }
        col.bitmap_offset = 0;
        col.len = rows;
        col.nil_count = nil_count;
    } else { /* all null: col.Append(nil, nil, nilBitmap, bitmapOffset, rows, nilCount, ...) */
        int rows = nil_count;
        col.bitmap.assign((size_t)(rows + 7) / 8, 0);
        col.len = rows; col.nil_count = nil_count; col.bitmap_offset = 0;
    }
    return E_OK;
}

int decode_time_page(const uint8_t *p, size_t len, ColVal &col) { /* appendTimeColumnData reader.go:638-672 */
    col.init();
    if (len < 1) return E_CORRUPT;
    if (p[0] == 18) { /* BlockIntegerOne */
        col.len = 1; col.nil_count = 0; col.bitmap.assign(1, 1);
        col.val.assign(p + 1, p + len);
        return E_OK;
    }
    const uint8_t *block, *bm; size_t bl, bml;
    ColVal hdr;
    int rc = decode_column_header(p, len, OG_TYPE_INT, hdr, &block, &bl, &bm, &bml);
    if (rc != E_OK) return rc;
    std::vector<int64_t> v;
    rc = time_block_decode(block, bl, v);
    if (rc != E_OK) return rc;
    col.val.resize(v.size() * 8);
    memcpy(col.val.data(), v.data(), v.size() * 8);
    col.len = (int)v.size();
    col.nil_count = 0; col.bitmap_offset = 0;
    col.bitmap.assign((size_t)(col.len + 7) / 8, 0xff);
    if (col.len & 7) col.bitmap.back() = (uint8_t)((1 << (col.len & 7)) - 1);
    return E_OK;
}

} // namespace ogo
