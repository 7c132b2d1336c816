/*
 * codec.cpp — CPU ORACLE (test infrastructure, see og_oracle.h): restatement of openGemini's column block codecs.
 * Reference files followed (relative to the openGemini tree) are cited at each function.
 */
#include <cmath>
#include <cstring>

#include "og_oracle.h"

namespace ogo {

/* ===================== lib/numberenc/number.go ===================== */
void put_u16be(Bytes &b, uint16_t v) { b.push_back((uint8_t)(v >> 8)); b.push_back((uint8_t)v); }          /* :47 */
void put_u32be(Bytes &b, uint32_t v) { for (int s = 24; s >= 0; s -= 8) b.push_back((uint8_t)(v >> s)); }   /* :58 */
void put_u64be(Bytes &b, uint64_t v) { for (int s = 56; s >= 0; s -= 8) b.push_back((uint8_t)(v >> s)); }   /* :73 */
uint16_t get_u16be(const uint8_t *p) { return (uint16_t)((p[0] << 8) | p[1]); }
uint32_t get_u32be(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
uint64_t get_u64be(const uint8_t *p) { return ((uint64_t)get_u32be(p) << 32) | get_u32be(p + 4); }

/* encoding/binary.PutUvarint / Uvarint (Go stdlib) */
int put_uvarint(Bytes &b, uint64_t v) {
    int n = 0;
    while (v >= 0x80) { b.push_back((uint8_t)v | 0x80); v >>= 7; n++; }
    b.push_back((uint8_t)v);
    return n + 1;
}
int get_uvarint(const uint8_t *p, size_t len, uint64_t *out) {
    uint64_t x = 0; unsigned s = 0;
    for (size_t i = 0; i < len; i++) {
        uint8_t c = p[i];
        if (i == 10) return -(int)(i + 1); /* overflow */
        if (c < 0x80) {
            if (i == 9 && c > 1) return -(int)(i + 1);
            *out = x | ((uint64_t)c << s);
            return (int)i + 1;
        }
        x |= (uint64_t)(c & 0x7f) << s;
        s += 7;
    }
    *out = 0;
    return 0;
}

/* ===================== simple8b (lib/util/lifted/encoding/simple8b/encoding.go) ===================== */
static const int S8B_N[16] = {240, 120, 60, 30, 20, 15, 12, 10, 8, 7, 6, 5, 4, 3, 2, 1};   /* selector table :193-210 */
static const int S8B_BITS[16] = {0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 15, 20, 30, 60};

/* canPack :444 — note the quirk: for bits==0 ALL remaining values must be 1, not only the first n (:455-462) */
static bool s8b_can_pack(const uint64_t *src, size_t len, int n, int bits) {
    if (len < (size_t)n) return false;
    if (bits == 0) {
        for (size_t i = 0; i < len; i++) if (src[i] != 1) return false;
        return true;
    }
    uint64_t max = (1ull << bits) - 1;
    for (int i = 0; i < n; i++) if (src[i] > max) return false;
    return true;
}

long s8b_encode_all(uint64_t *src, size_t n) { /* EncodeAll :350 */
    size_t i = 0, j = 0;
    while (i < n) {
        const uint64_t *rem = src + i;
        size_t rl = n - i;
        int sel = -1;
        for (int s = 0; s < 16; s++) {
            if (s8b_can_pack(rem, rl, S8B_N[s], S8B_BITS[s])) { sel = s; break; }
        }
        if (sel < 0) return E_INVAL; /* "value out of bounds" */
        uint64_t w;
        if (sel == 0) w = 0;                 /* pack240 :476 */
        else if (sel == 1) w = 1ull << 60;   /* pack120 (EncodeAll writes 1<<60 directly :366) */
        else {
            w = (uint64_t)sel << 60;
            int bits = S8B_BITS[sel];
            for (int k = 0; k < S8B_N[sel]; k++) w |= rem[k] << (k * bits); /* packN: value k at bit k*bits */
        }
        src[j++] = w;
        i += S8B_N[sel];
    }
    return (long)j;
}

int s8b_decode(uint64_t dst[240], uint64_t v) { /* Decode :419 + unpackN :739-975 */
    int sel = (int)(v >> 60);
    int n = S8B_N[sel], bits = S8B_BITS[sel];
    if (bits == 0) { for (int i = 0; i < n; i++) dst[i] = 1; return n; }
    uint64_t mask = (bits == 64) ? ~0ull : ((1ull << bits) - 1);
    for (int i = 0; i < n; i++) dst[i] = (v >> (i * bits)) & mask;
    return n;
}

/* ===================== Gorilla (tsm1/batch_float.go) ===================== */
static inline uint64_t f2u(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static inline double u2f(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }

namespace {
struct BitWriter { /* the encoder writes an MSB-first bit stream; bytes beyond the last bit are zero (:66-70,249-253) */
    Bytes &b; uint64_t n; /* bits written */
    void put(uint64_t v, unsigned nbits) { /* low nbits of v, MSB first */
        for (unsigned i = 0; i < nbits; i++) {
            uint64_t bit = (v >> (nbits - 1 - i)) & 1;
            size_t byte = (size_t)(n >> 3);
            while (byte >= b.size()) b.push_back(0);
            if (bit) b[byte] |= (uint8_t)(128 >> (n & 7));
            n++;
        }
    }
    void zero() { /* a zero bit only advances n; the byte is materialised lazily like the Go code (:63) */
        n++;
    }
};
} // namespace

int gorilla_encode(const double *src, size_t len, Bytes &b) { /* FloatArrayEncodeAll :17-254 */
    b.clear();
    b.push_back(1 << 4); /* floatCompressedGorilla << 4 (tsm1 header) :23 */
    double first; bool finished = false; size_t start = 0;
    if (len > 0 && std::isnan(src[0])) return E_NAN;          /* :27 */
    if (len == 0) { first = u2f(UVNAN); finished = true; }    /* :29-31 */
    else { first = src[0]; start = 1; }
    uint64_t prev = f2u(first);
    put_u64be(b, prev);                                       /* :42 */
    BitWriter w{b, 8 + 64};
    uint64_t prev_leading = ~0ull, prev_trailing = 0;
    double sum = 0;
    for (size_t i = start; !finished; i++) {
        double x;
        if (i < len) { x = src[i]; sum += x; }
        else { x = u2f(UVNAN); finished = true; }             /* sentinel :57-60 */
        uint64_t cur = (i < len) ? f2u(x) : UVNAN;
        uint64_t delta = cur ^ prev;
        if (delta == 0) { w.zero(); prev = cur; continue; }   /* :65-69 */
        w.put(1, 1);                                          /* :79-80 */
        uint64_t leading = (uint64_t)__builtin_clzll(delta);
        uint64_t trailing = (uint64_t)__builtin_ctzll(delta);
        leading &= 0x1F;                                      /* :89 — clz >= 32 wraps */
        if (leading >= 32) leading = 31;                      /* :90-92 (dead after the mask; kept for fidelity) */
        if (prev_leading != ~0ull && leading >= prev_leading && trailing >= prev_trailing) { /* :99 */
            w.put(0, 1);
            uint64_t l = 64 - prev_leading - prev_trailing;
            w.put(delta >> prev_trailing, (unsigned)l);       /* l LSBs of (delta>>prevTrailing), MSB first :109-137 */
        } else {
            prev_leading = leading; prev_trailing = trailing; /* :139 */
            w.put(1, 1);
            w.put(leading, 5);                                /* :146-172 */
            uint64_t sigbits = 64 - leading - trailing;       /* :179; 64 is written as 0 in 6 bits */
            w.put(sigbits & 0x3F, 6);
            w.put(delta >> trailing, (unsigned)sigbits);      /* :208-238 */
        }
        prev = cur;
    }
    if (std::isnan(sum)) return E_NAN;                        /* :245-247 (also trips on +Inf + -Inf) */
    size_t length = (size_t)(w.n >> 3) + ((w.n & 7) ? 1 : 0); /* :249-253 */
    b.resize(length, 0);
    return E_OK;
}

namespace {
struct BitReader { /* equivalent of brCachedVal/brValidBits over the whole buffer (:308-347) */
    const uint8_t *p; uint64_t nbits; uint64_t pos;
    bool read(unsigned k, uint64_t *out) {
        if (pos + k > nbits) return false;
        uint64_t v = 0;
        for (unsigned i = 0; i < k; i++) {
            uint64_t bp = pos + i;
            v = (v << 1) | ((p[bp >> 3] >> (7 - (bp & 7))) & 1);
        }
        pos += k; *out = v; return true;
    }
};
} // namespace

int gorilla_decode(const uint8_t *b, size_t len, std::vector<double> &out) { /* FloatArrayDecodeAll :278-514 */
    if (len < 9) return E_OK;                                 /* :279-281 returns empty */
    uint64_t val = get_u64be(b + 1);                          /* first byte = compression type, skipped :290 */
    if (val == UVNAN) return E_OK;                            /* :293-300 */
    size_t mark = out.size();
    out.push_back(u2f(val));
    BitReader br{b + 9, (uint64_t)(len - 9) * 8, 0};
    unsigned trailing = 0, meaningful = 64;                   /* :285-287 */
    if (br.nbits == 0) { out.resize(mark); return E_EOF; }    /* goto ERROR :345 */
    for (;;) {
        uint64_t bit;
        if (!br.read(1, &bit)) { out.resize(mark); return E_EOF; }
        if (bit) {
            if (!br.read(1, &bit)) { out.resize(mark); return E_EOF; }
            if (bit) {                                        /* '11': 5 bits leading + 6 bits meaningful :406-452 */
                uint64_t lm;
                if (!br.read(11, &lm)) { out.resize(mark); return E_EOF; }
                unsigned leading = (unsigned)((lm >> 6) & 0x1f);
                meaningful = (unsigned)(lm & 0x3f);
                if (meaningful > 0) trailing = 64 - leading - meaningful;
                else { trailing = 0; meaningful = 64; }
            }
            uint64_t sbits;
            if (!br.read(meaningful, &sbits)) { out.resize(mark); return E_EOF; }
            val ^= sbits << (trailing & 0x3f);                /* :499 */
            if (val == UVNAN) break;                          /* :500-503 */
        }
        out.push_back(u2f(val));                              /* :506 */
    }
    return E_OK;
}

/* ===================== RLE / Same (lib/compress/compress.go) ===================== */
static const unsigned RLE_BLOCK_LIMIT = 1 << 14; /* :25-27 */

void rle_same_encode(const double *v, size_t n, Bytes &out) { /* SameValueEncoding :38-49 */
    uint16_t size = (uint16_t)n;
    out.push_back((uint8_t)(size >> 8)); out.push_back((uint8_t)(size & 0xff));
    if (v[0] == 0) return; /* float compare: -0.0 == 0 also omits the value */
    uint64_t u = f2u(v[0]);
    for (int i = 0; i < 8; i++) out.push_back((uint8_t)(u >> (8 * i))); /* raw LE bytes of in[:8] */
}

int rle_same_decode(const uint8_t *in, size_t len, std::vector<double> &out) { /* SameValueDecoding :51-66 */
    if (len < 2) return E_CORRUPT;
    uint16_t size = get_u16be(in);
    if (len == 2) { out.insert(out.end(), size, 0.0); return E_OK; }
    if (len < 10) return E_CORRUPT; /* FailedToDecodeFloatArray */
    uint64_t u; memcpy(&u, in + 2, 8);
    out.insert(out.end(), size, u2f(u));
    return E_OK;
}

void rle_encode(const double *v, size_t size, Bytes &out) { /* RLE.Encoding :68-93 (values compared as uint64) */
    uint16_t n = 1;
    for (size_t i = 1; i <= size; i++) {
        if (i < size && f2u(v[i]) == f2u(v[i - 1]) && n < RLE_BLOCK_LIMIT) { n++; continue; }
        if (f2u(v[i - 1]) == 0) {
            uint16_t m = n | (1 << 15);
            out.push_back((uint8_t)(m >> 8)); out.push_back((uint8_t)(m & 0xff));
        } else {
            out.push_back((uint8_t)(n >> 8)); out.push_back((uint8_t)(n & 0xff));
            uint64_t u = f2u(v[i - 1]);
            for (int k = 0; k < 8; k++) out.push_back((uint8_t)(u >> (8 * k)));
        }
        n = 1;
    }
}

int rle_decode(const uint8_t *in, size_t len, std::vector<double> &out) { /* RLE.Decoding :95-121 */
    while (len >= 2) {
        uint16_t n = get_u16be(in);
        if (n >> 15) {
            n -= 1 << 15;
            out.insert(out.end(), n, 0.0);
            in += 2; len -= 2;
            continue;
        }
        if (len < 10) return E_CORRUPT;
        uint64_t u; memcpy(&u, in + 2, 8);
        out.insert(out.end(), n, u2f(u));
        in += 10; len -= 10;
    }
    return E_OK;
}

/* ===================== Snappy block format (third-party; see header note) ===================== */
int snappy_decoded_len(const uint8_t *in, size_t len, uint64_t *n, int *hdr) {
    int k = get_uvarint(in, len, n);
    if (k <= 0 || *n > 0xffffffffull) return E_CORRUPT;
    *hdr = k;
    return E_OK;
}

int snappy_decode(const uint8_t *in, size_t len, Bytes &out) {
    uint64_t dlen; int h;
    if (snappy_decoded_len(in, len, &dlen, &h) != E_OK) return E_CORRUPT;
    size_t base = out.size();
    out.reserve(base + dlen);
    size_t s = (size_t)h;
    while (s < len) {
        uint8_t tag = in[s];
        size_t length, offset;
        switch (tag & 3) {
        case 0: { /* literal */
            size_t x = tag >> 2;
            if (x < 60) { s += 1; }
            else if (x == 60) { if (s + 2 > len) return E_CORRUPT; x = in[s + 1]; s += 2; }
            else if (x == 61) { if (s + 3 > len) return E_CORRUPT; x = in[s + 1] | (in[s + 2] << 8); s += 3; }
            else if (x == 62) { if (s + 4 > len) return E_CORRUPT; x = in[s + 1] | (in[s + 2] << 8) | (in[s + 3] << 16); s += 4; }
            else { if (s + 5 > len) return E_CORRUPT; x = in[s + 1] | (in[s + 2] << 8) | (in[s + 3] << 16) | ((size_t)in[s + 4] << 24); s += 5; }
            length = x + 1;
            if (length > len - s || out.size() - base + length > dlen) return E_CORRUPT;
            out.insert(out.end(), in + s, in + s + length);
            s += length;
            continue;
        }
        case 1:
            if (s + 2 > len) return E_CORRUPT;
            length = 4 + ((tag >> 2) & 7);
            offset = ((size_t)(tag & 0xe0) << 3) | in[s + 1];
            s += 2;
            break;
        case 2:
            if (s + 3 > len) return E_CORRUPT;
            length = 1 + (tag >> 2);
            offset = in[s + 1] | (in[s + 2] << 8);
            s += 3;
            break;
        default:
            if (s + 5 > len) return E_CORRUPT;
            length = 1 + (tag >> 2);
            offset = in[s + 1] | (in[s + 2] << 8) | (in[s + 3] << 16) | ((size_t)in[s + 4] << 24);
            s += 5;
            break;
        }
        size_t d = out.size() - base;
        if (offset == 0 || offset > d || d + length > dlen) return E_CORRUPT;
        for (size_t i = 0; i < length; i++) out.push_back(out[out.size() - offset]);
    }
    if (out.size() - base != dlen) return E_CORRUPT;
    return E_OK;
}

static void snappy_emit_literal(const uint8_t *lit, size_t n, Bytes &out) {
    size_t x = n - 1;
    if (x < 60) out.push_back((uint8_t)(x << 2));
    else if (x < 256) { out.push_back(60 << 2); out.push_back((uint8_t)x); }
    else { out.push_back(61 << 2); out.push_back((uint8_t)x); out.push_back((uint8_t)(x >> 8)); }
    out.insert(out.end(), lit, lit + n);
}
static void snappy_emit_copy(size_t offset, size_t length, Bytes &out) {
    while (length >= 68) { out.push_back((63 << 2) | 2); out.push_back((uint8_t)offset); out.push_back((uint8_t)(offset >> 8)); length -= 64; }
    if (length > 64) { out.push_back((59 << 2) | 2); out.push_back((uint8_t)offset); out.push_back((uint8_t)(offset >> 8)); length -= 60; }
    if (length >= 12 || offset >= 2048) {
        out.push_back((uint8_t)(((length - 1) << 2) | 2)); out.push_back((uint8_t)offset); out.push_back((uint8_t)(offset >> 8));
    } else {
        out.push_back((uint8_t)(((offset >> 8) << 5) | ((length - 4) << 2) | 1)); out.push_back((uint8_t)offset);
    }
}

void snappy_encode(const uint8_t *in, size_t len, Bytes &out) {
    put_uvarint(out, len);
    /* blocks of <= 65536 bytes, greedy 4-byte hash matcher */
    size_t pos = 0;
    while (pos < len) {
        size_t blk = len - pos < 65536 ? len - pos : 65536;
        const uint8_t *src = in + pos;
        if (blk < 17) { snappy_emit_literal(src, blk, out); pos += blk; continue; }
        std::vector<int32_t> table(1 << 14, -1);
        size_t s = 0, lit = 0, limit = blk - 4;
        while (s <= limit) {
            uint32_t x; memcpy(&x, src + s, 4);
            uint32_t hsh = (x * 0x1e35a7bdu) >> 18;
            int32_t cand = table[hsh];
            table[hsh] = (int32_t)s;
            uint32_t y = 0;
            if (cand >= 0) memcpy(&y, src + cand, 4);
            if (cand >= 0 && y == x && s - (size_t)cand <= 65535) {
                if (s > lit) snappy_emit_literal(src + lit, s - lit, out);
                size_t m = 4;
                while (s + m < blk && src[cand + m] == src[s + m]) m++;
                snappy_emit_copy(s - (size_t)cand, m, out);
                s += m; lit = s;
            } else s++;
        }
        if (lit < blk) snappy_emit_literal(src + lit, blk - lit, out);
        pos += blk;
    }
}

/* ===================== adaptive float (lib/compress/float.go) ===================== */
enum { F_NULL = 0, F_OLD_GORILLA = 1, F_SNAPPY = 2, F_GORILLA = 3, F_SAME = 4, F_RLE = 5, F_MLF = 6 }; /* :26-33 */

static bool is_int(double f) { /* isInt :240-246 */
    if (f >= 0 && f < 4294967296.0) return (double)(uint64_t)f == f;
    return std::ceil(f) == f && std::floor(f) == f;
}
static bool less_decimal(double f) { return is_int(f * 1000); } /* :248-250 */

FloatContext float_generate_context(const double *v, size_t n) { /* GenerateContext :210-238 */
    FloatContext c;
    c.value_count = (int)n;
    if (n <= 4) return c;
    int distinct = 1;
    for (size_t i = 0; i < n; i++) {
        if (i > 0 && v[i] != v[i - 1]) distinct++;
        if (!c.extreme && std::isnan(v[i])) c.extreme = true;
    }
    c.distinct_count = distinct;
    if (c.distinct_count <= 8) return c;
    int k = 0, less_total = 0;
    for (int i = 0; i < c.value_count && k < c.value_count / 10; i++) {
        if (v[i] == 0) continue;
        k++;
        if (c.int_only && !is_int(v[i])) c.int_only = false;
        if (less_decimal(v[i])) less_total++;
    }
    c.less_decimal = k > 0 && (100 * less_total / k) > 90;
    return c;
}

static void compress_null(const double *v, size_t n, Bytes &out) { /* compressNull :133-137 */
    out.push_back(F_NULL << 4);
    const uint8_t *p = (const uint8_t *)v;
    out.insert(out.end(), p, p + n * 8);
}

int float_block_encode(const double *v, size_t n, Bytes &dst) { /* Float.Encoding lib/encoding/float.go:50-67 */
    if (n == 0) return E_OK;
    Bytes out; /* adaptiveEncoding float.go:60-101 works on out[pos:] */
    FloatContext ctx = float_generate_context(v, n);
    if (ctx.value_count <= 4) { compress_null(v, n, out); }
    else if (ctx.distinct_count == 1) { out.push_back(F_SAME << 4); rle_same_encode(v, n, out); }
    else if (ctx.distinct_count <= 8) { out.push_back(F_RLE << 4); rle_encode(v, n, out); }
    else {
        if ((!ctx.int_only && ctx.less_decimal) || ctx.extreme) {
            out.push_back(F_SNAPPY << 4);
            snappy_encode((const uint8_t *)v, n * 8, out);
        } else {
            Bytes g;
            int rc = gorilla_encode(v, n, g);
            if (rc != E_OK) return rc;
            out.push_back(F_GORILLA << 4);      /* :87-89: prepend the openGemini tag before tsm1's own 0x10 */
            out.insert(out.end(), g.begin(), g.end());
        }
        if (out.size() > n * 8 * 90 / 100) { out.clear(); compress_null(v, n, out); } /* :96-99 */
    }
    dst.insert(dst.end(), out.begin(), out.end());
    return E_OK;
}

/* legacy tag-1 decoder: lib/encoding/float.go:92-179 ([u32 count] + go-bitstream) */
static int old_gorilla_decode(const uint8_t *in, size_t len, std::vector<double> &out) {
    if (len < 4) return E_CORRUPT;
    size_t count = get_u32be(in);
    BitReader br{in + 4, (uint64_t)(len - 4) * 8, 0};
    std::vector<double> vals(count);
    if (count == 0) return E_OK;
    uint64_t v;
    if (!br.read(64, &v)) return E_EOF;
    vals[0] = u2f(v);
    size_t idx = 1;
    unsigned leading = 0, trailing = 0;
    for (;;) {
        uint64_t bit;
        if (!br.read(1, &bit)) return E_EOF;
        if (!bit) { if (idx >= count) return E_CORRUPT; vals[idx] = vals[idx - 1]; idx++; continue; }
        if (!br.read(1, &bit)) return E_EOF;
        if (bit) {
            uint64_t r;
            if (!br.read(5, &r)) return E_EOF;
            leading = (unsigned)r;
            if (!br.read(6, &r)) return E_EOF;
            unsigned mbits = (unsigned)r;
            if (mbits == 0) mbits = 64;
            trailing = 64 - leading - mbits;
        }
        unsigned mbits = 64 - leading - trailing;
        uint64_t r;
        if (!br.read(mbits, &r)) return E_EOF;
        uint64_t vb = f2u(vals[idx - 1]) ^ (r << trailing);
        double vv = u2f(vb);
        if (std::isnan(vv)) break;
        if (idx >= count) return E_CORRUPT;
        vals[idx++] = vv;
    }
    out.insert(out.end(), vals.begin(), vals.end());
    return E_OK;
}

int float_block_decode(const uint8_t *in, size_t len, std::vector<double> &out) { /* Float.Decoding :69-90 + AdaptiveDecoding :139-161 */
    if (len == 0) return E_OK; /* DecodeFloatBlock encoding.go:361-364 */
    int algo = in[0] >> 4;
    switch (algo) {
    case F_OLD_GORILLA: return old_gorilla_decode(in + 1, len - 1, out);
    case F_NULL: {
        size_t n = (len - 1) / 8;
        size_t base = out.size();
        out.resize(base + n);
        memcpy(out.data() + base, in + 1, n * 8);
        return E_OK;
    }
    case F_GORILLA: return gorilla_decode(in + 1, len - 1, out);
    case F_SNAPPY: {
        Bytes raw;
        int rc = snappy_decode(in + 1, len - 1, raw);
        if (rc != E_OK) return rc;
        size_t n = raw.size() / 8, base = out.size();
        out.resize(base + n);
        memcpy(out.data() + base, raw.data(), n * 8);
        return E_OK;
    }
    case F_SAME: return rle_same_decode(in + 1, len - 1, out);
    case F_RLE: return rle_decode(in + 1, len - 1, out);
    case F_MLF: return E_UNSUPPORTED;
    default: return E_CORRUPT; /* errno.InvalidFloatBuffer */
    }
}

/* ===================== int64 (lib/encoding/int.go) ===================== */
enum { I_CONST = 1, I_S8B = 2, I_ZSTD = 3, I_RAW = 4 }; /* :27-32 */

static void int_uncompressed(const int64_t *v, size_t n, Bytes &out) { /* uncompressedData :168-177 */
    out.push_back(I_RAW << 4);
    put_u32be(out, (uint32_t)(n * 8));
    for (size_t i = 0; i < n; i++) put_u64be(out, zigzag_enc(v[i])); /* MarshalInt64Append zigzags, number.go:156 */
}

int int_block_encode(const int64_t *arr, size_t n, Bytes &out) { /* Integer.Encoding :183-212 + init :73-99 */
    if (n == 0) return E_OK;
    if (n < 3) { int_uncompressed(arr, n, out); return E_OK; }
    std::vector<uint64_t> zz; zz.reserve(n);
    bool is_const = true, is_s8b = true;
    zz.push_back(zigzag_enc(arr[0]));
    uint64_t e = zigzag_enc((int64_t)((uint64_t)arr[1] - (uint64_t)arr[0]));
    if (e > S8B_MAX_VALUE) is_s8b = false;
    zz.push_back(e);
    for (size_t i = 2; i < n; i++) {
        e = zigzag_enc((int64_t)((uint64_t)arr[i] - (uint64_t)arr[i - 1]));
        is_const = is_const && zz[i - 1] == e;
        if (is_s8b && e > S8B_MAX_VALUE) is_s8b = false;
        zz.push_back(e);
    }
    if (is_const) { /* encodingConstDelta :101-121 */
        out.push_back(I_CONST << 4);
        put_u64be(out, zz[0]);
        put_uvarint(out, zz[1]);
        put_uvarint(out, (uint64_t)zz.size() - 1);
        return E_OK;
    }
    if (is_s8b) { /* encodingSimple8b :123-134 */
        long words = s8b_encode_all(zz.data() + 1, zz.size() - 1);
        if (words < 0) return (int)words;
        out.push_back(I_S8B << 4);
        put_u32be(out, (uint32_t)(words + 1));
        put_u32be(out, (uint32_t)zz.size());
        for (long i = 0; i < words + 1; i++) put_u64be(out, zz[i]);
        return E_OK;
    }
    return E_UNSUPPORTED; /* zstd :136-166 — klauspost/compress not restated */
}

int int_block_decode(const uint8_t *in, size_t len, std::vector<int64_t> &out) { /* Integer.Decoding :370-384 */
    if (len == 0) return E_OK; /* DecodeIntegerBlock encoding.go:336-339 */
    if (len < 5) return E_CORRUPT; /* decodeInit :327 */
    int ty = in[0] >> 4;
    in++; len--;
    switch (ty) {
    case I_RAW: { /* decodingUncompressed :316-324 */
        size_t bl = get_u32be(in); in += 4; len -= 4;
        if (len < bl) return E_CORRUPT;
        /* UnmarshalInt64Slice2Bytes converts len(src)/8 values of the *whole* remaining slice (number.go:130-139) */
        for (size_t i = 0; i < len / 8; i++) out.push_back(zigzag_dec(get_u64be(in + 8 * i)));
        return E_OK;
    }
    case I_CONST: { /* decodingConstDelta :214-254 */
        if (len < 8) return E_CORRUPT;
        uint64_t first = get_u64be(in); in += 8; len -= 8;
        uint64_t delta, cnt;
        int k = get_uvarint(in, len, &delta);
        if (k <= 0) return E_CORRUPT;
        in += k; len -= k;
        k = get_uvarint(in, len, &cnt);
        if (k <= 0) return E_CORRUPT;
        int64_t cur = zigzag_dec(first), d = zigzag_dec(delta);
        out.push_back(cur);
        for (uint64_t i = 1; i < cnt + 1; i++) { cur = (int64_t)((uint64_t)cur + (uint64_t)d); out.push_back(cur); }
        return E_OK;
    }
    case I_S8B: { /* decodingSimple8b :256-301 */
        if (len < 16) return E_CORRUPT;
        size_t enc = get_u32be(in), srcn = get_u32be(in + 4);
        in += 8; len -= 8;
        size_t l = enc * 8;
        if (len < l) return E_CORRUPT;
        size_t base = out.size();
        int64_t cur = zigzag_dec(get_u64be(in));
        out.push_back(cur);
        uint64_t vals[240];
        for (size_t pos = 8; pos < l; pos += 8) {
            int n = s8b_decode(vals, get_u64be(in + pos));
            for (int i = 0; i < n; i++) { cur = (int64_t)((uint64_t)cur + (uint64_t)zigzag_dec(vals[i])); out.push_back(cur); }
        }
        if (out.size() - base != srcn) return E_CORRUPT; /* the reference panics "idx != count+1" :296 */
        return E_OK;
    }
    case I_ZSTD: return E_UNSUPPORTED;
    default: return E_CORRUPT;
    }
}

/* ===================== timestamps (lib/encoding/timestamp.go) ===================== */
enum { T_CONST = 1, T_S8B = 2, T_SNAPPY = 3, T_RAW = 4 }; /* :27-32 */
static const uint64_t SCALES[12] = {10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull,
                                    1000000000ull, 10000000000ull, 100000000000ull, 1000000000000ull};
static uint64_t time_scale(uint64_t v) { /* scale :38-46 — index 0 (10) is never tried */
    for (int i = 11; i > 0; i--) if (v % SCALES[i] == 0) return SCALES[i];
    return 1;
}
static const double MIN_COMP_RATIO = 0.85; /* minCompReta, lib/encoding/encoding.go */

static void time_pack_raw(const int64_t *v, size_t n, Bytes &out) { /* packUncompressedData :85-94 */
    out.push_back(T_RAW << 4);
    put_u32be(out, (uint32_t)(n * 8));
    for (size_t i = 0; i < n; i++) put_u64be(out, zigzag_enc(v[i]));
}

int time_block_encode(const int64_t *tv, size_t n, Bytes &out) { /* Time.Encoding :150-164 */
    if (n < 3) { time_pack_raw(tv, n, out); return E_OK; }
    const uint64_t *times = (const uint64_t *)tv;
    std::vector<uint64_t> d(n); /* encodingInit :63-83 */
    bool is_const = true;
    d[n - 1] = times[n - 1] - times[n - 2];
    bool is_s8b = d[n - 1] < S8B_MAX_VALUE;
    uint64_t sc = time_scale(d[n - 1]);
    for (size_t i = n - 2; i > 0; i--) {
        d[i] = times[i] - times[i - 1];
        while (sc > 1 && d[i] % sc != 0) sc /= 10;
        is_const = is_const && d[i] == d[i + 1];
        is_s8b = is_s8b && d[i] < S8B_MAX_VALUE;
    }
    d[0] = times[0];
    if (is_const) { /* constDeltaEncoding :96-110 */
        out.push_back(T_CONST << 4);
        put_u64be(out, d[0]);
        put_uvarint(out, d[1]);
        put_uvarint(out, (uint64_t)n - 1);
        return E_OK;
    }
    if (is_s8b) { /* simple8bEncoding :112-130 */
        if (sc > 1) for (size_t i = 1; i < n; i++) d[i] /= sc;
        long words = s8b_encode_all(d.data() + 1, n - 1);
        if (words < 0) return (int)words;
        out.push_back(T_S8B << 4);
        put_u64be(out, sc);
        put_u32be(out, (uint32_t)(words + 1));
        put_u32be(out, (uint32_t)n);
        for (long i = 0; i < words + 1; i++) put_u64be(out, d[i]);
        return E_OK;
    }
    /* snappyEncoding :132-148 (byte parity unpinned: third-party encoder) */
    Bytes comp;
    snappy_encode((const uint8_t *)tv, n * 8, comp);
    if ((double)(9 + comp.size()) / (double)(n * 8) < MIN_COMP_RATIO) {
        out.push_back(T_SNAPPY << 4);
        put_u32be(out, (uint32_t)(n * 8));
        put_u32be(out, (uint32_t)comp.size());
        out.insert(out.end(), comp.begin(), comp.end());
        return E_OK;
    }
    time_pack_raw(tv, n, out);
    return E_OK;
}

int time_block_decode(const uint8_t *in, size_t len, std::vector<int64_t> &out) { /* Time.Decoding :310-324 */
    if (len < 5) return E_CORRUPT; /* decodingInit :176 */
    int ty = in[0] >> 4;
    in++; len--;
    switch (ty) {
    case T_RAW: { /* unpackUncompressedData :299-308 */
        size_t bl = get_u32be(in); in += 4; len -= 4;
        if (len < bl) return E_CORRUPT;
        for (size_t i = 0; i < len / 8; i++) out.push_back(zigzag_dec(get_u64be(in + 8 * i)));
        return E_OK;
    }
    case T_CONST: { /* constDeltaDecoding :190-225 */
        if (len < 8) return E_CORRUPT;
        uint64_t first = get_u64be(in); in += 8; len -= 8;
        uint64_t delta, cnt;
        int k = get_uvarint(in, len, &delta);
        if (k <= 0) return E_CORRUPT;
        in += k; len -= k;
        k = get_uvarint(in, len, &cnt);
        if (k <= 0) return E_CORRUPT;
        uint64_t cur = first;
        out.push_back((int64_t)cur);
        for (uint64_t i = 1; i < cnt + 1; i++) { cur += delta; out.push_back((int64_t)cur); }
        return E_OK;
    }
    case T_S8B: { /* simple8bDecoding :227-272 */
        if (len < 24) return E_CORRUPT;
        uint64_t sc = get_u64be(in);
        size_t enc = get_u32be(in + 8), srcn = get_u32be(in + 12);
        in += 16; len -= 16;
        size_t l = enc * 8;
        if (len < l) return E_CORRUPT;
        size_t base = out.size();
        uint64_t cur = get_u64be(in);
        out.push_back((int64_t)cur);
        uint64_t vals[240];
        for (size_t pos = 8; pos < l; pos += 8) {
            int n = s8b_decode(vals, get_u64be(in + pos));
            for (int i = 0; i < n; i++) { cur += vals[i] * sc; out.push_back((int64_t)cur); }
        }
        if (out.size() - base != srcn) return E_CORRUPT; /* panic "idx != srcCount" :268 */
        return E_OK;
    }
    case T_SNAPPY: { /* snappyDecoding :274-297 */
        if (len < 8) return E_CORRUPT;
        size_t srcl = get_u32be(in), compl_ = get_u32be(in + 4);
        in += 8; len -= 8;
        if (len < compl_) return E_CORRUPT;
        Bytes raw;
        int rc = snappy_decode(in, compl_, raw);
        if (rc != E_OK) return rc;
        if (raw.size() != srcl) return E_CORRUPT;
        size_t n = srcl / 8, base = out.size();
        out.resize(base + n);
        memcpy(out.data() + base, raw.data(), n * 8);
        return E_OK;
    }
    default: return E_CORRUPT;
    }
}

/* ===================== bool (lib/encoding/bool.go) ===================== */
int bool_block_encode(const uint8_t *v, size_t n, Bytes &out) { /* Boolean.Encoding :40-61 */
    if (n == 0) return E_OK; /* EncodeBooleanBlock encoding.go:376-379 */
    out.push_back(1 << 4);
    put_u32be(out, (uint32_t)n);
    uint8_t cur = 0; int cnt = 8;
    for (size_t i = 0; i < n; i++) { /* bitstream.WriteBit: MSB first */
        if (v[i]) cur |= (uint8_t)(1 << (cnt - 1));
        if (--cnt == 0) { out.push_back(cur); cur = 0; cnt = 8; }
    }
    if (cnt != 8) out.push_back(cur); /* Flush(Zero) */
    return E_OK;
}

int bool_block_decode(const uint8_t *in, size_t len, std::vector<uint8_t> &out) { /* Boolean.Decoding :63-96 */
    if (len == 0) return E_OK;
    if (len < 5) return E_CORRUPT;
    int ty = in[0] >> 4;
    size_t count = get_u32be(in + 1);
    if (ty != 1) return E_CORRUPT;
    in += 5; len -= 5;
    if (len * 8 < count) return E_EOF;
    for (size_t i = 0; i < count; i++) out.push_back((in[i >> 3] >> (7 - (i & 7))) & 1);
    return E_OK;
}

} // namespace ogo
