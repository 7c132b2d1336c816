"""The C-ABI library loads and exports every symbol include/ogpu.h declares (no compute calls: no GPU needed)."""
import ctypes
import os
import re

import pytest

from opengemini_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "ogpu.h")).read()
    return sorted(set(re.findall(r"OG_API\s+[\w\s\*]+?\b(og_\w+)\s*\(", txt)))


def test_header_and_binding_agree():
    assert _header_symbols() == sorted(L.EXPORTS)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(L.LIB_PATH):
        pytest.fail(f"{L.LIB_PATH} not built: run __graft_entry__.build()")
    lib = ctypes.CDLL(L.LIB_PATH)
    for sym in _header_symbols():
        assert hasattr(lib, sym), f"missing export {sym}"


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the product must fail loudly, not compute on the CPU."""
    lib = L.lib()
    if lib.og_device_count() > 0:
        pytest.skip("GPU present")
    assert lib.og_init(0) == L.OG_E_CUDA
    cols = (L.SynthColumn * 1)()
    cols[0].type, cols[0].dist = L.TYPE_FLOAT, L.SYNTH_F_HI
    d = L.SynthDesc(1, 10, 1000, 0, 1, 1, 1, cols)
    h = ctypes.c_void_p()
    assert lib.og_shard_synth(ctypes.byref(d), ctypes.byref(h)) == L.OG_E_CUDA
    assert b"no CUDA device" in lib.og_last_error() or b"CUDA" in lib.og_last_error()


def test_product_never_references_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "opengemini_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                if "oracle" in open(os.path.join(dirpath, f), errors="replace").read().lower().replace("the oracle's", "").replace("oracle", "oracle") and \
                   re.search(r"(#include|import|from)\s+[\"<\.\w/]*oracle", open(os.path.join(dirpath, f), errors="replace").read()):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, f"product files include/import the oracle: {bad}"


def test_ctypes_struct_layouts_match_the_header(tmp_path):
    """sizeof / offsetof of every ABI struct as the C compiler sees include/ogpu.h == the ctypes mirror in _lib.py."""
    import ctypes as C
    import subprocess
    pairs = {"og_filter_item": L.FilterItem, "og_call": L.Call, "og_query_desc": L.QueryDesc, "og_column_desc": L.ColumnDesc,
             "og_shard_desc": L.ShardDesc, "og_colval_view": L.ColValView, "og_record_view": L.RecordView, "og_dense_col": L.DenseCol,
             "og_dense_view": L.DenseView, "og_stats": L.Stats, "og_synth_column": L.SynthColumn, "og_synth_desc": L.SynthDesc,
             "og_shard_layout": L.ShardLayout}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "ogpu.h")}"', "int main(void) {"]
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _t in cls._fields_:
            cfield = {"func": "func"}.get(fname, fname)
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {cfield}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    seen = dict(line.split() for line in out.strip().splitlines())
    for cname, cls in pairs.items():
        assert int(seen[cname]) == C.sizeof(cls), cname
        for fname, _t in cls._fields_:
            assert int(seen[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"
