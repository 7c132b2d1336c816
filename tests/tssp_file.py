"""Test infrastructure: a TSSP container WRITER in plain Python, restated from the Go marshal code (not from
opengemini_b200/csrc/tssp.cpp, which is the reader under test):

  engine/immutable/msbuilder.go:1260-1275,1355-1425   header, section order, footer
  engine/immutable/chunkdata_builder_ts.go:37-82        chunk bytes: per column [u32 BE crc32 (IEEE) of its pages][pages]
  engine/immutable/tssp_file_meta.go:566-581,228-246,86-91,129-133   ChunkMeta / ColumnMeta / Segment / SegmentRange
  engine/immutable/tssp_file_meta.go:769-778            MetaIndex
  engine/immutable/msbuilder.go:1481-1500               chunk-meta block = metas + u32 BE start offsets
  engine/immutable/trailer.go:59-66, table_stat.go:36-49,122-143   trailer, TableStat, ExtraData
  lib/numberenc/number.go:155-160                       int64 = zig-zag, big endian

The bloom filter and id-time sections are opaque to a scan; they are written as filler of the sizes the trailer states.
"""
import struct
import zlib

M64 = (1 << 64) - 1


def i64(v):
    return struct.pack(">Q", ((v << 1) ^ (v >> 63)) & M64)


def u64(v):
    return struct.pack(">Q", v & M64)


def u32(v):
    return struct.pack(">I", v)


def u16(v):
    return struct.pack(">H", v)


def build(chunks, measurement=b"mst_0000", metas_per_block=3, bloom_bytes=24, idtime_bytes=11, compress_flag=0, legacy_extra=None):
    """chunks: list of dicts {sid, tmin:[...], tmax:[...], columns:[(name bytes, type, [page bytes per segment])...], time:[page bytes...]}
    in ascending sid order.  Returns (file bytes, directory) where directory[sid][name] = [(offset, size) per segment]."""
    out = bytearray(b"53ac2021" + u64(2))
    directory = {}
    metas = []
    for ch in chunks:
        nseg = len(ch["time"])
        chunk_off = len(out)
        cols = list(ch["columns"]) + [(b"time", 1, ch["time"])]
        col_meta = bytearray()
        entry = {}
        for name, ty, pages in cols:
            assert len(pages) == nseg
            body = b"".join(pages)
            out += u32(zlib.crc32(body) & 0xFFFFFFFF)
            segs = []
            for p in pages:
                segs.append((len(out), len(p)))
                out += p
            entry[name] = segs
            pre = b"\x07preagg-of-time" if name == b"time" else b""
            col_meta += u16(len(name)) + name + bytes([ty]) + u16(len(pre)) + pre
            for off, size in segs:
                col_meta += i64(off) + u32(size)
        directory[ch["sid"]] = entry
        meta = u64(ch["sid"]) + i64(chunk_off) + u32(len(out) - chunk_off) + u32(len(cols)) + u32(nseg)
        for a, b in zip(ch["tmin"], ch["tmax"]):
            meta += i64(a) + i64(b)
        metas.append((ch["sid"], min(ch["tmin"]), max(ch["tmax"]), bytes(meta + col_meta)))
    data_end = len(out)
    # chunk-meta blocks and their index entries
    index = bytearray()
    mi = []
    for b0 in range(0, len(metas), metas_per_block):
        blk = metas[b0:b0 + metas_per_block]
        start = len(index)
        offs, cur = [], 0
        for _, _, _, m in blk:
            offs.append(cur)
            index += m
            cur += len(m)
        for o in offs:
            index += u32(o)
        mi.append((blk[0][0], min(x[1] for x in blk), max(x[2] for x in blk), data_end + start, len(blk), len(index) - start))
    out += index
    mi_bytes = b"".join(u64(i) + i64(a) + i64(b) + i64(off) + u32(cnt) + u32(size) for i, a, b, off, cnt, size in mi)
    out += mi_bytes
    out += bytes(bloom_bytes)
    out += bytes(range(idtime_bytes))
    trailer_off = len(out)
    tr = i64(16) + i64(data_end - 16) + i64(len(index)) + i64(len(mi_bytes)) + i64(bloom_bytes) + i64(idtime_bytes)
    sids = [c["sid"] for c in chunks]
    tr += i64(len(sids)) + u64(min(sids)) + u64(max(sids)) + i64(min(m[1] for m in metas)) + i64(max(m[2] for m in metas)) + i64(len(mi)) + u64(bloom_bytes * 8) + u64(4)
    if legacy_extra is not None:            # 1- or 2-byte forms of older files (table_stat.go:181-195)
        tr += u16(len(legacy_extra)) + legacy_extra
    else:
        extra_len = 8 + 2                   # flags + "number of header values" (no ChunkMetaHeader)
        flags = 1 | (compress_flag << 8) | (extra_len << 32)
        tr += u16(8) + struct.pack("<Q", flags) + u16(0)
    tr += u16(len(measurement)) + measurement
    out += tr
    out += i64(trailer_off)
    return bytes(out), directory
