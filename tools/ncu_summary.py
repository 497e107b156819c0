"""summarise an .ncu-rep (raw page + SASS sampling) into a small text file for profiles/"""
import csv, collections, os, subprocess, sys
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


def write_counters(rep, n_elems, out_path):
    """profiles/kernel_counters.json: what bench.py quotes about the fused kernel, tied to the
    kernel sources it was captured from (sha1) — executed FP64 work, pipe cycles, DRAM bytes"""
    import hashlib, json, os, re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    h = hashlib.sha1()
    for f in ("assemble_kernels.cuh", "mitc4_math.h", "mitc4_tying.h"):
        h.update(open(os.path.join(root, "a2d-shells_b200", "csrc", f), "rb").read())
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    d = {hh: (v, u) for hh, u, v in zip(rows[0], rows[1], rows[2])}
    def val(k, scale={"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0}):
        v, u = d[k]
        return float(v) * scale.get(u, 1.0)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]; ix = {hh: i for i, hh in enumerate(hdr)}
    thr = collections.Counter(); wrp = collections.Counter()
    for r in rows[2:]:
        if len(r) < len(hdr) or not r[ix["# Samples"]].isdigit():
            continue
        m = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
        if not m:
            continue
        op = m.group(1).split(".")[0]
        if op in ("DFMA", "DMUL", "DADD", "DMMA"):
            wrp[op] += int(r[ix["Instructions Executed"]])
            thr[op] += int(r[ix["Predicated-On Thread Instructions Executed"]])
    flops = 2 * thr["DFMA"] + thr["DMUL"] + thr["DADD"] + 512 * wrp["DMMA"]
    pipe_cycles = 2 * (wrp["DFMA"] + wrp["DMUL"] + wrp["DADD"]) + 16 * wrp["DMMA"]   # per SM sub-partition
    n_smsp = 148 * 4
    clock = 1.965e9
    out = {
        "source_hash": h.hexdigest(), "kernel": d["Kernel Name"][0], "n_elems": n_elems,
        "report": "profiles/" + os.path.basename(rep).replace(".ncu-rep", ".txt") +
                  " (ncu --set full --clock-control none)",
        "kernel_time_under_ncu_s": val("gpu__time_duration.sum"),
        "fp64_flops_per_elem": flops / n_elems,
        "fp64_thread_instr_per_elem": {k: thr[k] / n_elems for k in ("DFMA", "DMUL", "DADD")},
        "dmma_per_elem": wrp["DMMA"] / n_elems,
        "warp_instr_per_elem": float(d["smsp__inst_executed.sum"][0]) / n_elems,
        "fp64_pipe_cycles_per_elem": pipe_cycles / n_elems,
        "fp64_pipe_ceiling_elems_per_s": n_smsp * clock / (pipe_cycles / n_elems),
        "fp64_pipe_busy": float(d["sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active"][0]) / 100.0,
        "dram_bytes_per_elem": (val("dram__bytes_read.sum") + val("dram__bytes_write.sum")) / n_elems,
        "warps_active_pct": float(d["sm__warps_active.avg.pct_of_peak_sustained_active"][0]),
    }
    json.dump(out, open(out_path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if "--counters" in sys.argv:
    i = sys.argv.index("--counters")
    write_counters(rep, int(sys.argv[i + 1]),
                   sys.argv[i + 2] if len(sys.argv) > i + 2 else
                   os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "kernel_counters.json"))
