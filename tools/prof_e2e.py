"""Phase timing of the e2e leg (host-resident shard -> og_shard_open -> og_query_run -> og_query_next).  usage: prof_e2e.py [series]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
from opengemini_b200 import AggQuery, Shard, _lib as L
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
T0, SEC, rows = 1_700_000_000_000_000_000, 1_000_000_000, 1_000_000
if len(sys.argv) > 2 and sys.argv[2] == "affinity":
    import pynvml
    pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
    print("cpu affinity before", len(os.sched_getaffinity(0)))
    pynvml.nvmlDeviceSetCpuAffinity(h)
    print("cpu affinity after nvmlDeviceSetCpuAffinity", sorted(os.sched_getaffinity(0))[:4], "...", len(os.sched_getaffinity(0)))
os.system("nvidia-smi topo -m 2>/dev/null | head -4; numactl -H 2>/dev/null | head -3")
Shard.init(0)
small = Shard.synth(ns, rows, [(L.TYPE_FLOAT, L.SYNTH_F_HI, 0)], t0=T0, dt=SEC, seed=1000)
ex = small.export()
class _L: pass
lay = _L(); lay.data_len = ex["data"].size
pinned = torch.empty(lay.data_len, dtype=torch.uint8, pin_memory=True)
pinned.numpy()[:] = ex["data"]
small.close()
host = pinned.numpy()
calls = [("sum", 0), ("count", 0), ("max", 0)]
def sync(): torch.cuda.synchronize()
# raw copy rate for reference
dev = torch.empty(lay.data_len, dtype=torch.uint8, device="cuda"); sync()
t = time.perf_counter(); dev.copy_(pinned, non_blocking=True); sync(); dt = time.perf_counter() - t
print(f"raw pinned->device {lay.data_len/dt/1e9:.1f} GB/s ({dt*1e3:.1f} ms)"); del dev; torch.cuda.empty_cache()
for it in range(3):
    t0 = time.perf_counter()
    s2 = Shard.open(host, ex["sids"], ex["series_seg_begin"], ex["seg_tmin"], ex["seg_tmax"], [("f0", L.TYPE_FLOAT, ex["page_off"][0], ex["page_len"][0])], ex["page_off"][1], ex["page_len"][1])
    t1 = time.perf_counter()
    q2 = AggQuery(s2, calls, 60 * SEC, T0, T0 + (rows - 1) * SEC)
    t2 = time.perf_counter()
    q2.run()
    t3 = time.perf_counter()
    n = sum(r["rows"] for r in q2.records())
    t4 = time.perf_counter()
    st = q2.stats()
    q2.close(); t5 = time.perf_counter(); s2.close(); t6 = time.perf_counter()
    print(f"open {1e3*(t1-t0):.1f}  create {1e3*(t2-t1):.1f}  run {1e3*(t3-t2):.1f} (kernels {st['kernel_ms']:.1f})  records {1e3*(t4-t3):.1f}  qclose {1e3*(t5-t4):.1f}  sclose {1e3*(t6-t5):.1f}  total {1e3*(t6-t0):.1f} ms  rows/s {ns*rows/(t6-t0):.3e}")
