"""summarise an .ncu-rep (raw page + SASS sampling) into a small text file for profiles/"""
import csv, collections, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "derived__smsp__sass_thread_inst_executed_op_dfma_pred_on_x2", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"]
for k in want:
    if k in d: print(f"{k:75s} {d[k][0]} {d[k][1]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr) and r[ix["# Samples"]].isdigit()]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter(); n = 0
for r in data:
    n += int(r[ix["# Samples"]])
    for c in stall_cols:
        if r[ix[c]].isdigit(): tot[c] += int(r[ix[c]])
print("\nwarp-state samples:", n)
for c, v in tot.most_common(10): print(f"  {c:26s} {100*v/max(n,1):5.1f}%")
# per source function (needs the library that was profiled still built in-tree):
# samples, executed warp instructions, FP64-pipe warp instructions (DFMA/DMUL/DADD, DMMA)
if len(sys.argv) > 2:
    import os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import sass_account
    table = sass_account.line_table(sys.argv[2])
    if len(table) == len(data):
        agg = collections.defaultdict(lambda: [0, 0, 0, 0])
        for (off, fn, text), r in zip(table, data):
            a = agg[fn]
            ex = int(r[ix["Instructions Executed"]])
            a[0] += int(r[ix["# Samples"]]); a[1] += ex
            cl = sass_account.classify(text)
            if cl in ("DFMA", "DMUL", "DADD"): a[2] += ex
            if cl == "DMMA": a[3] += ex
        tot_ex = sum(a[1] for a in agg.values())
        print("\nby source function: % samples | % warp instructions | FP64 warp instr | DMMA")
        for fn, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            print(f"  {fn:28s} {100*a[0]/max(n,1):5.1f}%  {100*a[1]/max(tot_ex,1):5.1f}%  {a[2]:10d} {a[3]:9d}")
    else:
        print(f"\n(per-function table skipped: library has {len(table)} SASS lines, report {len(data)})")
top = sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:12]
print("\nhottest SASS instructions:")
for r in top: print(f"  {int(r[ix['# Samples']]):6d}  {r[ix['Source']][:80]}")
