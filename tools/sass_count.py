"""static SASS instruction mix of the kernels whose mangled name contains the given substrings
usage: python tools/sass_count.py [lib.so] substr [substr ...]   (development aid)"""
import collections, os, re, subprocess, sys, tempfile
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
lib = os.path.join(root, "a2d-shells_b200", "lib", "liba2ds_b200.so")
if args and args[0].endswith(".so"):
    lib = args.pop(0)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = max((f for f in os.listdir(tmp) if f.endswith(".cubin")), key=lambda f: os.path.getsize(os.path.join(tmp, f)))
sass = subprocess.run(["nvdisasm", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
cur = None
funcs = collections.OrderedDict()
for l in sass:
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", l)
    if m and cur:
        funcs[cur].append(m.group(1))
groups = [("DMMA", r"^DMMA"), ("DFMA", r"^DFMA"), ("DMUL", r"^DMUL"), ("DADD", r"^DADD"), ("LDS", r"^LDS"),
          ("STS", r"^STS"), ("SHFL", r"^SHFL"), ("RED/ATOM", r"^(RED|ATOM)"), ("LDG/LDC", r"^(LDG|LDC|ULDC)"), ("LDGSTS", r"^LDGSTS"),
          ("SEL", r"^(SEL|FSEL)"), ("IMAD/IADD", r"^(IMAD|IADD|LEA)"), ("MOV", r"^(MOV|UMOV|CS2R)"), ("LOP/SHF", r"^(LOP3|SHF|PRMT)"),
          ("BRA/SYNC", r"^(BRA|BSSY|BSYNC|WARPSYNC|NOP|EXIT)"), ("LDL/STL", r"^(LDL|STL)")]
for name, ops in funcs.items():
    if not any(a in name for a in args):
        continue
    cnt = collections.Counter()
    for o in ops:
        for g, pat in groups:
            if re.match(pat, o):
                cnt[g] += 1; break
        else:
            cnt["other"] += 1
    print(name, len(ops), "instr,", " ".join(f"{g}={cnt[g]}" for g, _ in groups if cnt[g]), f"other={cnt['other']}")
