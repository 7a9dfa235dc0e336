"""Per-region summary of an ncu source page (cuda,sass): instructions per tile-pair, share of stall samples,
cycles per instruction and the top stall reasons.
  python tools/ncu_roles.py cs.csv n_tile_pairs warps_per_sm kernel_cycles file:lo-hi:name ..."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
ntp = float(sys.argv[2])
specs = []
for spec in sys.argv[3:]:
    f, rng, name = spec.split(":"); lo, hi = map(int, rng.split("-")); specs.append((f, lo, hi, name))
def region(f, l):
    for ff, lo, hi, n in specs:
        if f and f.startswith(ff) and lo <= l <= hi: return n
    return "other"
cur = None; curfile = None; ix = None; stalls = []
agg = collections.defaultdict(collections.Counter)
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        curfile = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No":
        ix = {h: i for i, h in enumerate(r)}; stalls = [h for h in r if h.startswith('stall_') and 'Not Issued' not in h]; continue
    if len(r) > 7 and r[0].isdigit() and r[2] == "-":
        cur = (curfile, int(r[0]))
    elif len(r) > 7 and r[2].startswith('0x'):
        reg = region(*(cur or (None, 0)))
        for st in stalls:
            try: agg[reg][st] += int(r[ix[st]])
            except Exception: pass
        agg[reg]['_inst'] += int(r[7])
tot_s = sum(sum(v for k, v in c.items() if k != '_inst') for c in agg.values())
tot_i = sum(c['_inst'] for c in agg.values())
print("total: %.0f warp instr per tile-pair, %d samples" % (tot_i / ntp, tot_s))
for reg, c in sorted(agg.items(), key=lambda kv: -sum(v for k, v in kv[1].items() if k != '_inst')):
    tot = sum(v for k, v in c.items() if k != '_inst')
    top = sorted(((v, k) for k, v in c.items() if k != '_inst'), reverse=True)[:5]
    print("%-10s samples %5.1f%%  inst/tp %5.0f  " % (reg, 100.0 * tot / tot_s, c['_inst'] / ntp) +
          " ".join("%s=%.0f%%" % (k[6:], 100 * v / max(tot, 1)) for v, k in top))
