"""Thread scaling of the CPU oracle scan on this host (diagnostic for bench.py's cpu_baseline)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import oracle
from opengemini_b200 import _lib as L
T0, SEC = 1_700_000_000_000_000_000, 1_000_000_000
ns, rows = int(sys.argv[1]) if len(sys.argv) > 1 else 512, 100000
print("nproc", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    if os.path.exists(f): print(f, open(f).read().strip())
os.system("lscpu | grep -E 'Model name|Socket|Core|Thread|^CPU\\(s\\)'")
t = time.time(); hs = oracle.HostShard(ns, rows, [(L.TYPE_FLOAT, L.SYNTH_F_HI, 0)], t0=T0, dt=SEC, seed=1000, threads=os.cpu_count()); print("build s", time.time() - t)
calls = (L.Call * 3)((L.AGG_SUM, 0), (L.AGG_COUNT, 0), (L.AGG_MAX, 0))
q = L.QueryDesc(60 * SEC, 0, T0, T0 + (rows - 1) * SEC, 1, 3, calls, 0, None, L.GROUP_ALL, 1, None, 0, 0)
for th in (1, 4, 16, 32, 64, 128, 256):
    if th > 2 * os.cpu_count(): break
    t = time.time(); oracle.scan(hs.desc, q, threads=th); dt = time.time() - t
    print(th, "threads", round(ns * rows / dt / 1e6, 1), "M rows/s")
