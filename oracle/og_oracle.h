/*
 * og_oracle.h — CPU ORACLE for the openGemini scan/aggregate hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may build, link, load or execute anything under oracle/.  The product
 * (opengemini_b200/, libogpu.so) never includes or calls it.
 *
 * It is a C++17 restatement of the reference's algorithms (the Go toolchain is absent from this image, so the
 * reference itself cannot be compiled or run here or on the GPU box — SURVEY.md §0.4, §8c).  Every function
 * cites the reference file:line it follows.
 *
 * PARITY PINNING STATUS
 *   - aggregate cursor / reducers / interval record: PINNED by the reference's known-answer tests
 *     (engine/iterators_test.go:748-2043, 2956-3050), restated in oracle/kat_tests.cpp.
 *   - codecs: the reference holds round-trip tests only and no golden encoded bytes, so
 *     byte-level wire parity is pinned only by reading the encoder source ("parity unpinned by vectors");
 *     value-level parity is pinned by the restated round-trip suites (lib/encoding/encoding_test.go:49-843,
 *     timestamp_test.go:25-124 incl. the literal 32-timestamp vector :28-37, lib/compress/float_test.go:67-200).
 *   - Snappy (github.com/golang/snappy v0.0.5-0.20231225225746-43d5d4cd4e0e, klauspost/compress v1.17.11) is a
 *     third-party dependency absent from /root/reference: the decoder follows the published Snappy block
 *     format; the encoder here emits a valid block but not necessarily the same bytes.
 *   - zstd (int tag 3), lz4, MLF (float tag 6): not restated; encode/decode return OGO_E_UNSUPPORTED.
 */
#ifndef OG_ORACLE_H
#define OG_ORACLE_H

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../include/ogpu.h"

namespace ogo {

enum { E_OK = 0, E_INVAL = -1, E_UNSUPPORTED = -4, E_CORRUPT = -5, E_NAN = -20, E_EOF = -21 };

typedef std::vector<uint8_t> Bytes;

/* ---------- lib/numberenc/number.go ---------- */
void put_u16be(Bytes &b, uint16_t v);
void put_u32be(Bytes &b, uint32_t v);
void put_u64be(Bytes &b, uint64_t v);
uint16_t get_u16be(const uint8_t *p);
uint32_t get_u32be(const uint8_t *p);
uint64_t get_u64be(const uint8_t *p);
inline uint64_t zigzag_enc(int64_t v) { return ((uint64_t)v << 1) ^ (uint64_t)(v >> 63); }
inline int64_t zigzag_dec(uint64_t u) { return (int64_t)(u >> 1) ^ -(int64_t)(u & 1); }
int put_uvarint(Bytes &b, uint64_t v);
/* returns bytes consumed (>0) or <=0 on error, like encoding/binary.Uvarint */
int get_uvarint(const uint8_t *p, size_t len, uint64_t *out);

/* ---------- lib/util/lifted/encoding/simple8b/encoding.go ---------- */
static const uint64_t S8B_MAX_VALUE = (1ull << 60) - 1;
/* EncodeAll :350 — in place; returns number of words written to src[0..], or <0 */
long s8b_encode_all(uint64_t *src, size_t n);
/* Decode :419 — returns n values written to dst[240], or <0 */
int s8b_decode(uint64_t dst[240], uint64_t word);

/* ---------- tsm1 Gorilla: lib/util/lifted/influxdb/tsdb/engine/tsm1/batch_float.go ---------- */
static const uint64_t UVNAN = 0x7FF8000000000001ull; /* float.go:17 */
int gorilla_encode(const double *src, size_t n, Bytes &out);       /* FloatArrayEncodeAll :17 (out is replaced) */
int gorilla_decode(const uint8_t *b, size_t len, std::vector<double> &out); /* FloatArrayDecodeAll :278 (out appended) */

/* ---------- lib/compress/compress.go ---------- */
void rle_same_encode(const double *v, size_t n, Bytes &out);          /* SameValueEncoding :38 */
int rle_same_decode(const uint8_t *in, size_t len, std::vector<double> &out); /* :51 */
void rle_encode(const double *v, size_t n, Bytes &out);               /* RLE.Encoding :68 */
int rle_decode(const uint8_t *in, size_t len, std::vector<double> &out);      /* :95 */
int snappy_decode(const uint8_t *in, size_t len, Bytes &out);         /* Snappy block format */
void snappy_encode(const uint8_t *in, size_t len, Bytes &out);        /* valid block, bytes not pinned */
int snappy_decoded_len(const uint8_t *in, size_t len, uint64_t *n, int *hdr);

/* ---------- lib/compress/float.go ---------- */
struct FloatContext { /* Context :163 */
    int value_count = 0, distinct_count = 1;
    bool int_only = true, less_decimal = true, extreme = false;
};
FloatContext float_generate_context(const double *v, size_t n); /* GenerateContext :210 */
int float_block_encode(const double *v, size_t n, Bytes &out);  /* encoding.Float.Encoding lib/encoding/float.go:50 -> adaptiveEncoding float.go:60 (appends) */
int float_block_decode(const uint8_t *in, size_t len, std::vector<double> &out); /* Float.Decoding lib/encoding/float.go:69 -> AdaptiveDecoding float.go:139 */

/* ---------- lib/encoding/int.go, timestamp.go, bool.go ---------- */
int int_block_encode(const int64_t *v, size_t n, Bytes &out);   /* Integer.Encoding :183 (appends) */
int int_block_decode(const uint8_t *in, size_t len, std::vector<int64_t> &out); /* Integer.Decoding :370 */
int time_block_encode(const int64_t *v, size_t n, Bytes &out);  /* Time.Encoding :150 */
int time_block_decode(const uint8_t *in, size_t len, std::vector<int64_t> &out); /* Time.Decoding :310 */
int bool_block_encode(const uint8_t *v, size_t n, Bytes &out);  /* Boolean.Encoding :40 */
int bool_block_decode(const uint8_t *in, size_t len, std::vector<uint8_t> &out); /* Boolean.Decoding :63 */

/* ---------- lib/record/column.go ---------- */
struct ColVal {
    Bytes val;                 /* non-null values only, densely packed LE (8 B int/float, 1 B bool) */
    std::vector<uint32_t> offset;
    Bytes bitmap;              /* LSB-first, 1 = present, bit BitMapOffset+i */
    int bitmap_offset = 0;
    int len = 0;
    int nil_count = 0;

    void init() { val.clear(); offset.clear(); bitmap.clear(); bitmap_offset = 0; len = 0; nil_count = 0; }
    bool is_nil(int i) const;              /* IsNil */
    int valid_count(int start, int end) const; /* ValidCount :297 */
    void value_index_range(int bm_start, int bm_end, int *s, int *e) const; /* getValIndexRange :453 */
    void append_bit(bool present);         /* setBitMap/resetBitMap via AppendXxx */
    void append_integer(int64_t v);
    void append_float(double v);
    void append_boolean(bool v);
    void append_null(int type, bool reserve); /* AppendXxxNull / AppendXxxNullReserve */
    const int64_t *integers() const { return (const int64_t *)val.data(); }
    const double *floats() const { return (const double *)val.data(); }
    const uint8_t *booleans() const { return val.data(); }
    size_t n_values(int type) const { return type == OG_TYPE_BOOL ? val.size() : val.size() / 8; }
};

struct Field { std::string name; int type; };
struct Record { /* lib/record/record.go:57-61 ; last column is time */
    std::vector<Field> schema;
    std::vector<ColVal> cols;
    std::vector<std::vector<int64_t>> meta_times; /* RecMeta.Times */
    explicit Record(const std::vector<Field> &s = {}) : schema(s), cols(s.size()), meta_times(s.size()) {}
    int row_nums() const { return cols.empty() ? 0 : cols.back().len; }
    const int64_t *times() const { return cols.back().integers(); }
    int64_t time(int i) const { return times()[i]; }
    void append_time(int64_t t) { cols.back().append_integer(t); }
    int field_index(const std::string &n) const;
    void reset();
};

/* ---------- engine/immutable/column_builder.go (segment = page framing) ---------- */
/* EncodeColumnHeader :428 + enc*Column :151-349 for one segment (appends the page to out) */
int encode_field_page(const ColVal &col, int type, Bytes &out);
/* chunkdata_builder.go:65 EncodeTime for one segment */
int encode_time_page(const int64_t *t, size_t n, Bytes &out);
/* reader.go:674 decodeColumnData -> append{Integer,Float,Boolean}Column :504-579 */
int decode_field_page(const uint8_t *p, size_t len, int type, ColVal &col);
/* reader.go:638 appendTimeColumnData */
int decode_time_page(const uint8_t *p, size_t len, ColVal &col);

/* ---------- lib/util/lifted/influx/query/select.go:579 Window (Location == nil) ---------- */
struct WindowOpt { int64_t interval = 0, offset = 0, start_time = 0, end_time = 0; };
void window(const WindowOpt &o, int64_t t, int64_t *start, int64_t *end);

/* ---------- engine/aggregate_cursor.go + series_agg_reducer.gen.go + series_agg_func.gen.go ---------- */
struct ExprOpt { int func; std::string in_name; std::string out_name; }; /* hybridqp.ExprOptions restricted to Call(VarRef) */
struct AggCursor; /* opaque */
AggCursor *agg_cursor_new(const std::vector<Field> &in_schema, const std::vector<Field> &out_schema,
                          const std::vector<ExprOpt> &exprs, const WindowOpt &w, int chunk_size);
void agg_cursor_free(AggCursor *c);
/* input records are pulled through this callback (KeyCursor.Next of the child); return nullptr at end */
typedef const Record *(*NextFn)(void *ctx);
void agg_cursor_set_input(AggCursor *c, NextFn fn, void *ctx);
/* aggregateCursor.Next :267 — returns nullptr at end of stream; the record is owned by the cursor */
const Record *agg_cursor_next(AggCursor *c);

/* ---------- engine/agg_tagset_cursor.go:959-1120 + lib/record/reccord_functions.go (dense interval record) ---------- */
struct IntervalRecord {
    std::vector<Field> schema;   /* output fields + time */
    std::vector<ExprOpt> exprs;
    int64_t start = 0, interval = 0;
    uint32_t n_rows = 0;
    bool multi = false;
    bool has_interval = false;
    /* per field column: dense slots */
    std::vector<std::vector<uint64_t>> values; /* raw 8-byte cells (double bits / int64 / bool 0,1) */
    std::vector<std::vector<uint8_t>> valid;
    std::vector<std::vector<int64_t>> col_times; /* RecMeta.Times (multi-call first/last) */
    std::vector<int64_t> times;                  /* time column (window start unless replaced by a selector row) */
    void build(int64_t min, int64_t max, int64_t interval, bool has_interval); /* BuildEmptyIntervalRec record.go:1327 */
    void update_from(const Record &rec);         /* RecordInit :1069 / UpdateRec :1111 for every row of rec */
};

/* ---------- end-to-end CPU path over a shard (the cpu_baseline / reference arm) ---------- */
struct ScanResult {
    uint32_t n_groups = 0, n_buckets = 0;
    int64_t start = 0, interval = 0;
    /* [col][group*n_buckets + b] */
    std::vector<std::vector<uint64_t>> values;
    std::vector<std::vector<uint8_t>> valid;
    std::vector<std::vector<int64_t>> times; /* selector row times; for sum/count = window start */
    uint64_t rows_decoded = 0, segments = 0, page_bytes = 0;
};
/* Runs the reference-structured pull loop (Location.readData -> decode -> FilterByTime/FilterByField ->
 * aggregateCursor -> AggTagSetCursor merge) with `threads` workers striding series like group cursors do
 * (engine/file_cursor.go:190-195).  series_begin/series_end bound the sample. */
int scan_aggregate(const og_shard_desc &shard, const og_query_desc &q, int threads, uint32_t series_begin,
                   uint32_t series_end, ScanResult &out);

/* The CPU BASELINE leg for the headline query shape (fast_scan.cpp): the same path with the reference's batch Gorilla decoder
 * (64-bit cached bit reader, batch_float.go:308-347) instead of the checker's bit-serial one.  E_UNSUPPORTED for anything but one
 * float column, no WHERE, one tagset, count/sum(/min/max in multi-call queries) over Gorilla/raw pages with const-delta times. */
int fast_scan_aggregate(const og_shard_desc &shard, const og_query_desc &q, int threads, uint32_t series_begin,
                        uint32_t series_end, ScanResult &out);

/* synthetic shard builder (host): same distributions as include/ogpu_synth.h, encoded with the restated encoders */
struct HostShard {
    Bytes data;
    std::vector<uint64_t> sids;
    std::vector<uint32_t> series_seg_begin;
    std::vector<int64_t> seg_tmin, seg_tmax;
    std::vector<std::vector<uint64_t>> page_off; /* [n_columns+1][n_segments], time last */
    std::vector<std::vector<uint32_t>> page_len;
    std::vector<int32_t> col_types;
    std::vector<std::string> col_names;
    std::vector<og_column_desc> col_descs;
    og_shard_desc desc() ;
};
int build_synth_shard(const og_synth_desc &d, HostShard &out, int threads = 1);

} // namespace ogo

#endif
