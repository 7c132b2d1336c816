"""Summarise an .ncu-rep: key metrics + hottest SASS lines.  usage: ncu_summary.py file.ncu-rep [min_pct]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "sm__cycles_elapsed.max",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]
for i, h in enumerate(hdr):
    if h in want or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
        try:
            v = float(vals[i])
            if h.startswith("smsp__average_warps") and v < 0.05: continue
        except ValueError:
            pass
        print(f"{h} = {vals[i]} {rows[1][i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(src)))
print(r[0][1][:120])
h = r[1]; d = r[2:]
ia, isrc, ith, isamp = h.index("Instructions Executed"), h.index("Source"), h.index("Avg. Threads Executed"), h.index("# Samples")
tot = sum(int(x[ia]) for x in d); tsamp = sum(int(x[isamp]) for x in d)
print("total warp instrs", tot, "samples", tsamp)
for i, x in enumerate(d):
    c = int(x[ia])
    if c > tot * minpct / 100:
        print(f"{i:5d} {x[isrc].strip()[:64]:64s} {c:10d} {100*c/tot:4.1f}% thr {x[ith]:>3s} samp {100*int(x[isamp])/tsamp:4.1f}%")
