"""Top stall / instruction lines of an ncu report's source page.
  ncu -i rep.ncu-rep --page source --csv --print-source sass > sass.csv ; python tools/ncu_top.py sass.csv [N]
  ncu -i rep.ncu-rep --page source --csv --print-source cuda > cuda.csv ; python tools/ncu_top.py cuda.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
# find the header row
hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
def I(r, k):
    try: return int(r[ix[k]])
    except Exception: return 0
ts = sum(I(r, "# Samples") for r in data) or 1
ti = sum(I(r, "Instructions Executed") for r in data) or 1
print("samples", ts, "warp-instr", ti)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in sorted(data, key=lambda r: -I(r, "# Samples"))[:N]:
    top = sorted(((I(r, s), s) for s in stalls), reverse=True)[:2]
    print("%6s %5.2f%% samp %5.2f%% inst  %-70s %s" % (r[0][-6:], 100 * I(r, "# Samples") / ts, 100 * I(r, "Instructions Executed") / ti,
          r[1].strip()[:70], " ".join("%s=%d" % (s[6:], v) for v, s in top if v)))
