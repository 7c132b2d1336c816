"""configs[2] shape: int64 (Simple8b) + float64 (Gorilla walk) + bool columns, count/sum with a WHERE filter (generic tile path).
usage: prof_mixed.py [series] [rows] [null_permille]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opengemini_b200 import AggQuery, Shard, _lib as L
T0, SEC = 1_700_000_000_000_000_000, 1_000_000_000
series = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
npm = int(sys.argv[3]) if len(sys.argv) > 3 else 0
Shard.init(0)
cols = [(L.TYPE_INT, L.SYNTH_INT_WALK, npm), (L.TYPE_FLOAT, L.SYNTH_F_LO, npm), (L.TYPE_BOOL, L.SYNTH_BOOL, npm)]
sh = Shard.synth(series, rows, cols, t0=T0, dt=SEC, seed=77)
info = sh.info()
calls = [("count", 0), ("sum", 0), ("sum", 1), ("count", 2)]
for name, flt in (("WHERE f > 1000", [("term", 1, ">", 1000.0)]), ("no filter", None)):
    q = AggQuery(sh, calls, 60 * SEC, T0, T0 + (rows - 1) * SEC, filter=flt)
    for _ in range(3):
        q.run()
    st = q.stats()
    print(name, {k: st[k] for k in ("kernel_ms", "main_kernel_ms", "kernel_launches", "path")},
          "rows/s %.3e  page GB/s %.1f  (3 columns, %d rows, %.2f B/row pages)" % (st["rows_decoded"] / st["kernel_ms"] * 1e3, st["page_bytes"] / st["kernel_ms"] / 1e6, info["n_rows"], st["page_bytes"] / info["n_rows"]))
    q.close()
