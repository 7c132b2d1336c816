"""Staged smoke of the fused kernel with short timeouts (run each stage under `timeout`): prints stats or the error."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from opengemini_b200 import AggQuery, Shard
from opengemini_b200 import _lib as L
T0, SEC = 1_700_000_000_000_000_000, 1_000_000_000
stage = sys.argv[1]
ns, rows = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (33, 3000)
Shard.init(0)
cols = [(L.TYPE_FLOAT, L.SYNTH_F_HI, 0)]
sh = Shard.synth(ns, rows, cols, t0=T0, dt=SEC, seed=3)
print("synth ok", sh.info(), flush=True)
flags = {"strict": L.Q_STRICT_ORDER, "fold": 0, "nofast": L.Q_NO_FAST}[stage]
calls = [("sum", 0), ("count", 0), ("max", 0)]
t = time.time()
try:
    q = AggQuery(sh, calls, 60 * SEC, T0, T0 + (rows - 1) * SEC, flags=flags).run()
except Exception as e:
    print("RUN FAILED:", e, flush=True); sys.exit(3)
print("run ok %.3fs" % (time.time() - t), q.stats(), flush=True)
got = q.dense_host()
if ns * rows <= 2_000_000:
    hs = oracle.HostShard(ns, rows, cols, t0=T0, dt=SEC, seed=3)
    ref = oracle.scan(hs.desc, q.desc, threads=1)
    for k in range(3):
        m = ref["cols"][k]["valid"].astype(bool)
        assert np.array_equal(got["cols"][k]["valid"].astype(bool), m), ("valid", k)
        g, r = got["cols"][k]["values"].view(np.uint64)[m], ref["cols"][k]["values"][m]
        nbad = int((g != r).sum())
        rel = np.abs(g.view(np.float64) - r.view(np.float64)).max() if k == 0 else 0
        print("call", k, "bitwise mismatches", nbad, "max abs diff", rel, flush=True)
        assert nbad == 0 or (k == 0 and stage == "fold" and rel < 1e-6), k
print("stage", stage, "OK", flush=True)
