"""shared-memory wavefronts (total / excessive) and stall samples of an ncu capture attributed to
source lines of the CURRENT build of the library (run right after the capture, same sources).
usage: python tools/ncu_conflicts.py report.ncu-rep mangled-substring [top N]"""
import csv, subprocess, collections, re, os, sys, tempfile
rep, key = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr) and r[ix["# Samples"]].isdigit()]
def num(r, c):
    try: return float(r[ix[c]])
    except ValueError: return 0.0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "a2d-shells_b200/lib/liba2ds_b200.so")], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = max((f for f in os.listdir(tmp) if f.endswith(".cubin")), key=lambda f: os.path.getsize(os.path.join(tmp, f)))
sass = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
cur = False; pend = []; table = []; last = [("?", 0)]
for l in sass:
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m: cur = key in m.group(1); continue
    if not cur: continue
    m = re.match(r'\s*//## File ".*/([^/"]+)", line (\d+)(.*)', l)
    if m: pend.append((m.group(1), int(m.group(2)))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        if pend: last = pend; pend = []
        table.append(last)
assert len(table) == len(data), (len(table), len(data), "library and capture differ")
ex = collections.Counter(); tot = collections.Counter(); smp = collections.Counter(); ins = collections.Counter()
for t, r in zip(table, data):
    inner = f"{t[0][0]}:{t[0][1]}"
    ex[inner] += num(r, "L1 Wavefronts Shared Excessive"); tot[inner] += num(r, "L1 Wavefronts Shared")
    smp[inner] += num(r, "# Samples"); ins[inner] += num(r, "Instructions Executed")
src = {}
def text(k):
    f, ln = k.split(":")
    if f not in src:
        try: src[f] = open(os.path.join(root, "a2d-shells_b200/csrc", f)).read().splitlines()
        except OSError: src[f] = []
    return src[f][int(ln) - 1].strip()[:90] if src[f] and int(ln) <= len(src[f]) else ""
n = 1e6
print("excessive shared wavefronts (M) | total | line")
for k, v in ex.most_common(top):
    print(f"{v/n:8.1f} {tot[k]/n:8.1f}  {k:26s} {text(k)}")
print("total excessive", sum(ex.values()) / n, "of", sum(tot.values()) / n)
print("\nstall samples | executed (M) | line")
for k, v in smp.most_common(top):
    print(f"{v:8.0f} {ins[k]/n:8.1f}  {k:26s} {text(k)}")
