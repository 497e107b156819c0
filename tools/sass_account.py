"""static SASS accounting: FP64 / DMMA / shared-memory instruction counts of one k_assemble
instantiation, attributed to the source function (by line range) they were inlined from.
usage: python tools/sass_account.py [mangled-substring, default ILb1ELb1ELb1ELb0E]
(development aid; counts are static, the batched phases run once per 4 elements)"""
import collections, os, re, subprocess, sys, tempfile
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "a2d-shells_b200", "lib", "liba2ds_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
# the library is linked from several sources: take the cubin that holds the kernels
cubin = max((f for f in os.listdir(tmp) if f.endswith(".cubin")),
            key=lambda f: os.path.getsize(os.path.join(tmp, f)))
sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)],
                      capture_output=True, text=True).stdout.splitlines()
# function line ranges of the two sources
ranges = {}
for fn in ("a2d-shells_b200/csrc/mitc4_math.h", "a2d-shells_b200/csrc/assemble_kernels.cuh"):
    src = open(os.path.join(root, fn)).read().splitlines()
    marks = []
    for i, l in enumerate(src, 1):
        m = re.match(r"^(?:A2DS_HD|__device__ __forceinline__|__global__|template.*__global__|static|inline)?.*?\b([A-Za-z_0-9]+)\(.*[,{(]\s*$", l)
        if m and not l.startswith(" ") and not l.startswith("//") and not l.startswith("#"):
            marks.append((i, m.group(1)))
    if fn.endswith("assemble_kernels.cuh"):  # the kernel body: one bucket per "// ----" phase comment
        k0 = next(i for i, l in enumerate(src, 1) if re.match(r"\s+k_assemble\(const KParams", l))
        k1 = next(i for i, l in enumerate(src, 1) if i > k0 and l.startswith("}"))
        marks = [mk for mk in marks if not (k0 <= mk[0] <= k1)]
        marks.append((k0, "k_assemble: setup/gather"))
        for i in range(k0, k1):
            m = re.match(r"\s*// ---- (.*?)[- ]*$", src[i - 1])
            if m:
                marks.append((i, "k: " + m.group(1)[:24]))
        marks.sort()
    ranges[os.path.basename(fn)] = marks
def func_of(f, line):
    marks = ranges.get(os.path.basename(f))
    if not marks:
        return os.path.basename(f)
    name = "?"
    for i, n in marks:
        if i <= line:
            name = n
        else:
            break
    return name
classes = [("DMMA", r"\bDMMA"), ("DFMA", r"\bDFMA"), ("DMUL", r"\bDMUL"), ("DADD", r"\bDADD"),
           ("Dother", r"\b(DSETP|MUFU|F2F|I2F\.F64|DMNMX)"), ("LDS", r"\bLDS"), ("STS", r"\bSTS"),
           ("SHFL", r"\bSHFL"), ("LDG", r"\bLDG|\bLDL|\bSTL"), ("RED", r"\bRED|\bATOM")]
def classify(text):
    for name, pat in classes:
        if re.search(pat, text):
            return name
    return None


def line_table(key):
    """[(byte offset, source function, instruction text)] of the k_assemble instantiation
    whose mangled name contains `key`"""
    out = []
    inside = False
    cur = ("?", 0)
    for l in sass:
        if l.startswith(".text."):
            inside = key in l and "k_assemble" in l
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", l)
        if not m:
            continue
        out.append((int(m.group(1), 16), func_of(*cur), m.group(2)))
    return out


key = sys.argv[1] if len(sys.argv) > 1 and __name__ == "__main__" else "ILb1ELb1ELb1ELb0E"
tab = collections.defaultdict(lambda: collections.Counter())
total = collections.Counter()
for off, fn, text in (line_table(key) if __name__ == "__main__" else []):
    total["inst"] += 1
    tab[fn]["inst"] += 1
    c = classify(text)
    if c:
        tab[fn][c] += 1
        total[c] += 1
cols = ["inst"] + [c for c, _ in classes]
if __name__ == "__main__":
  print(f"{'function':28s}" + "".join(f"{c:>8s}" for c in cols))
for fn, cnt in sorted(tab.items(), key=lambda kv: -kv[1]["inst"]):
  if __name__ == "__main__":
    print(f"{fn:28s}" + "".join(f"{cnt[c]:8d}" for c in cols))
if __name__ == "__main__":
  print(f"{'TOTAL':28s}" + "".join(f"{total[c]:8d}" for c in cols))
