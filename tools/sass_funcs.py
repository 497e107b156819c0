"""per source function (innermost inlined frame) static instruction counts of one kernel,
split into the per-trip part and the per-element loop.  usage: sass_funcs.py substr"""
import collections, os, re, subprocess, sys, tempfile
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
key = sys.argv[1]
lib = os.path.join(root, "a2d-shells_b200", "lib", "liba2ds_b200.so")
csrc = os.path.join(root, "a2d-shells_b200/csrc")
marks = {}
for fn in os.listdir(csrc):
    if not fn.endswith((".h", ".cuh")): continue
    mk = []
    for i, l in enumerate(open(os.path.join(csrc, fn)).read().splitlines(), 1):
        m = re.match(r"^(?:A2DS_HD|__device__ __forceinline__|static inline|inline)\s+\S+\s+([A-Za-z_0-9]+)\(", l)
        if not m: m = re.match(r"^\s{4}(k_assemble(?:_t)?|k_mass)\(const KParams", l)
        if m: mk.append((i, m.group(1)))
    marks[fn] = mk
def func_of(f, line):
    name = f
    for i, n in marks.get(f, []):
        if i <= line: name = n
        else: break
    return name
src = open(os.path.join(csrc, "assemble_kernels.cuh")).read().splitlines()
loops = []
for i, l in enumerate(src, 1):
    if l.strip() == "#pragma unroll 1":
        j = next((k for k in range(i, len(src)) if "drawn = __shfl_sync" in src[k - 1] or src[k - 1].startswith("}")), len(src))
        loops.append((i, j))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = max((f for f in os.listdir(tmp) if f.endswith(".cubin")), key=lambda f: os.path.getsize(os.path.join(tmp, f)))
sass = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
cur = False; outer = 0; inner = None; pend = []
tab = collections.defaultdict(lambda: collections.Counter())
for l in sass:
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        cur = key in m.group(1); continue
    if not cur: continue
    m = re.match(r'\s*//## File ".*/([^/"]+)", line (\d+)(.*)', l)
    if m:
        pend.append((m.group(1), int(m.group(2)), "inlined at" in m.group(3))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", l)
    if m:
        if pend:
            inner = func_of(pend[0][0], pend[0][1])
            o = [p for p in pend if p[0] == "assemble_kernels.cuh" and not p[2]]
            if o: outer = o[-1][1]
            pend = []
        part = "E" if any(a <= outer <= b for a, b in loops) else "T"
        op = m.group(1)
        cls = "fp64" if re.match(r"^(DFMA|DMUL|DADD)", op) else "dmma" if op.startswith("DMMA") else "lds" if re.match(r"^(LDS|STS)", op) else "other"
        tab[(part, inner)][cls] += 1
for part in "TE":
    rows = [(sum(c.values()), n, c) for (p, n), c in tab.items() if p == part]
    print("== per trip" if part == "T" else "== per element", sum(r[0] for r in rows))
    for tot, n, c in sorted(rows, reverse=True)[:22]:
        print(f"  {n:28s} {tot:5d}  fp64 {c['fp64']:4d} dmma {c['dmma']:3d} lds/sts {c['lds']:4d} other {c['other']:4d}")
