"""per source line (outermost kernel line + innermost line) aggregation of an ncu source page:
samples, executed instructions, shared-memory wavefronts (ideal / excessive).
usage: python tools/ncu_lines.py report.ncu-rep [top N]    (development aid)"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr) and r[ix["# Samples"]].isdigit()]
def num(r, c):
    v = r[ix[c]]
    try: return float(v)
    except ValueError: return 0.0
tot_s = sum(num(r, "# Samples") for r in data); tot_i = sum(num(r, "Instructions Executed") for r in data)
print("instructions", len(data), "samples", tot_s, "executed", tot_i)
print("shared wavefronts", sum(num(r, "L1 Wavefronts Shared") for r in data), "ideal", sum(num(r, "L1 Wavefronts Shared Ideal") for r in data),
      "excessive", sum(num(r, "L1 Wavefronts Shared Excessive") for r in data))
ex = sorted(data, key=lambda r: -num(r, "L1 Wavefronts Shared Excessive"))[:top]
print("\nmost excessive shared wavefronts:")
for r in ex:
    print(f"  {r[ix['Address']][-5:]} {r[ix['Source']][:60]:60s} exec {num(r,'Instructions Executed'):12.0f} wave {num(r,'L1 Wavefronts Shared'):12.0f} ideal {num(r,'L1 Wavefronts Shared Ideal'):12.0f}")
st = sorted(data, key=lambda r: -num(r, "# Samples"))[:top]
print("\nmost sampled:")
for r in st:
    reasons = sorted(((num(r, c), c) for c in hdr if c.startswith("stall_") and "Not Issued" not in c), reverse=True)[:2]
    print(f"  {r[ix['Address']][-5:]} {r[ix['Source']][:60]:60s} samples {num(r,'# Samples'):7.0f} " + " ".join(f"{c[6:]}={v:.0f}" for v, c in reasons))
