"""Small driver for ncu: build a synthetic shard, run the headline query a few times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opengemini_b200 import AggQuery, Shard, _lib as L
T0, SEC = 1_700_000_000_000_000_000, 1_000_000_000
series = int(sys.argv[1]) if len(sys.argv) > 1 else 400
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
dist = L.SYNTH_F_LO if (len(sys.argv) > 3 and sys.argv[3] == "lo") else L.SYNTH_F_HI
runs = int(sys.argv[4]) if len(sys.argv) > 4 else 3
flags = int(sys.argv[5]) if len(sys.argv) > 5 else 0
Shard.init(0)
sh = Shard.synth(series, rows, [(L.TYPE_FLOAT, dist, 0)], t0=T0, dt=SEC, seed=1000)
q = AggQuery(sh, [("sum", 0), ("count", 0), ("max", 0)], 60 * SEC, T0, T0 + (rows - 1) * SEC, flags=flags)
for _ in range(runs):
    q.run()
st = q.stats()
print({k: st[k] for k in ("kernel_ms", "main_kernel_ms", "rows_decoded", "page_bytes", "kernel_launches", "path")})
print("rows/s %.3e  main GB/s %.1f" % (st["rows_decoded"] / st["kernel_ms"] * 1e3, st["page_bytes"] / st["main_kernel_ms"] / 1e6))
