"""Secondary measurement of configs[1]: the same scan grouped per series (1 bucket row per series and minute).  usage: prof_per_series.py [series]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opengemini_b200 import AggQuery, Shard, _lib as L
T0, SEC = 1_700_000_000_000_000_000, 1_000_000_000
series = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
rows = 1_000_000
Shard.init(0)
sh = Shard.synth(series, rows, [(L.TYPE_FLOAT, L.SYNTH_F_HI, 0)], t0=T0, dt=SEC, seed=1000)
for group in ("all", "series"):
    q = AggQuery(sh, [("sum", 0), ("count", 0), ("max", 0)], 60 * SEC, T0, T0 + (rows - 1) * SEC, group=group)
    for _ in range(3):
        q.run()
    st = q.stats()
    print(group, {k: st[k] for k in ("kernel_ms", "main_kernel_ms", "kernel_launches")}, "rows/s %.3e" % (st["rows_decoded"] / st["kernel_ms"] * 1e3))
    q.close()
