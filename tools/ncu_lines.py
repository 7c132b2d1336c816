"""Per CUDA source line instruction / stall-sample breakdown of an .ncu-rep.  usage: ncu_lines.py file.ncu-rep units [min_pct]
units = number of (warp-)work items to normalise by (e.g. rows/32)."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; units = float(sys.argv[2]); minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = collections.OrderedDict(); fname = ""; h = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": h = r; ia = h.index("Instructions Executed"); isamp = h.index("# Samples"); ith = h.index("Thread Instructions Executed"); continue
    if h is None or len(r) != len(h): continue
    try: c = int(r[ia] or 0)
    except ValueError: continue
    if r[0] != "": cur = int(r[0]); cursrc = r[1].strip()
    k = (fname, cur)
    a = agg.setdefault(k, [cursrc, 0, 0, 0, 0])
    a[1] += c; a[2] += int(r[isamp] or 0); a[3] += int(r[ith] or 0); a[4] += 1
tot = sum(a[1] for a in agg.values()); ts = sum(a[2] for a in agg.values())
print(f"total warp instr {tot}  per unit {tot/units:.1f}  samples {ts}")
for (f, ln), a in agg.items():
    if a[1] > tot * minpct / 100 or a[2] > ts * minpct / 100:
        print(f"{f[:16]:16s}:{ln:4d} {a[1]/units:6.2f}/u {100*a[1]/tot:4.1f}% thr {a[3]/max(1,a[1]):4.1f} samp {100*a[2]/max(1,ts):4.1f}% sass {a[4]:3d} | {a[0][:100]}")
