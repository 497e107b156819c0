"""static SASS mix of a k_assemble / k_assemble_t instantiation, split into the per-trip part
(gather + batched node / Gauss-point phases: executed once per NB elements) and the
per-element loop, by the outermost source line of each instruction (-lineinfo).
usage: python tools/sass_split.py [lib.so] mangled-substring ...   (development aid)"""
import collections, os, re, subprocess, sys, tempfile
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
lib = os.path.join(root, "a2d-shells_b200", "lib", "liba2ds_b200.so")
if args and args[0].endswith(".so"):
    lib = args.pop(0)
src = open(os.path.join(root, "a2d-shells_b200/csrc/assemble_kernels.cuh")).read().splitlines()
# per-element loops: from '#pragma unroll 1' to the 'drawn = __shfl_sync' that follows it
loops = []
for i, l in enumerate(src, 1):
    if l.strip() == "#pragma unroll 1":
        j = next((k for k in range(i, len(src)) if "drawn = __shfl_sync" in src[k - 1] or src[k - 1].startswith("}")), len(src))
        loops.append((i, j))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = max((f for f in os.listdir(tmp) if f.endswith(".cubin")), key=lambda f: os.path.getsize(os.path.join(tmp, f)))
sass = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
groups = [("DMMA", r"^DMMA"), ("DFMA", r"^DFMA"), ("DMUL", r"^DMUL"), ("DADD", r"^DADD"), ("LDS", r"^LDS"),
          ("STS", r"^STS"), ("SHFL", r"^SHFL"), ("RED", r"^(RED|ATOM)"), ("LDG", r"^(LDG|LDC|ULDC|LDCU)"),
          ("SEL", r"^(SEL|FSEL)"), ("IMAD", r"^(IMAD|IADD|LEA|VIADD)"), ("MOV", r"^(MOV|UMOV|CS2R)"), ("LOP", r"^(LOP3|SHF|PRMT)"),
          ("CTRL", r"^(BRA|BSSY|BSYNC|WARPSYNC|NOP|EXIT)"), ("LDL/STL", r"^(LDL|STL)")]
cur = None; outer = 0; inner_fn = ""
res = {}
for l in sass:
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        cur = m.group(1) if any(a in m.group(1) for a in args) else None
        continue
    if cur is None:
        continue
    m = re.match(r'\s*//## File ".*/([^/"]+)", line (\d+)(.*)', l)
    if m:
        if m.group(1) == "assemble_kernels.cuh" and "inlined at" not in m.group(3):
            outer = int(m.group(2))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", l)
    if m:
        part = "element" if any(a <= outer <= b for a, b in loops) else "trip"
        op = m.group(1)
        g = next((g for g, pat in groups if re.match(pat, op)), "other")
        res.setdefault(cur, {}).setdefault(part, collections.Counter())[g] += 1
for name, parts in res.items():
    print(name)
    for part in ("trip", "element"):
        c = parts.get(part, collections.Counter())
        tot = sum(c.values()); fp = c["DFMA"] + c["DMUL"] + c["DADD"]
        print(f"  {part:8s} {tot:5d} instr, FP64 {fp:4d} " + " ".join(f"{g}={c[g]}" for g, _ in groups if c[g]) + f" other={c['other']}")
