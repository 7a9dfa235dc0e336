"""I-cache footprint of an ncu source page (cuda,sass): distinct 128-byte code lines that are hot, per source region.
  python tools/ncu_lines.py cs.csv threshold file:lo-hi:name ..."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
th = int(sys.argv[2])
specs = []
for spec in sys.argv[3:]:
    f, rng, name = spec.split(":"); lo, hi = map(int, rng.split("-")); specs.append((f, lo, hi, name))
def region(f, l):
    for ff, lo, hi, n in specs:
        if f and f.startswith(ff) and lo <= l <= hi: return n
    return "other:" + (f or "?")
cur = None; curfile = None
addr = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        curfile = r[1].split('/')[-1]; continue
    if len(r) > 7 and r[0].isdigit() and r[2] == "-":
        cur = (curfile, int(r[0]))
    elif len(r) > 7 and r[2].startswith('0x'):
        a = int(r[2], 16)
        try: x = int(r[7])
        except Exception: continue
        if a not in addr: addr[a] = (region(*(cur or (None, 0))), x)
base = min(addr)
lines = collections.defaultdict(lambda: [0, collections.Counter()])
for a, (reg, x) in addr.items():
    l = (a - base) // 128
    lines[l][0] = max(lines[l][0], x); lines[l][1][reg] += 1
hot = {l: v for l, v in lines.items() if v[0] > th}
print("code lines: %d total (%.1f KB), %d hot (%.1f KB) at exec > %d" % (len(lines), len(lines) / 8, len(hot), len(hot) / 8, th))
per = collections.Counter()
for l, v in hot.items():
    tot = sum(v[1].values())
    for reg, n in v[1].items(): per[reg] += n / tot
for reg, n in per.most_common(): print("  %-22s %5.1f lines %5.2f KB" % (reg, n, n / 8))
