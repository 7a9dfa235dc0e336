"""Sum executed warp instructions / stall samples of an ncu source page (cuda,sass) over named line ranges.
  python tools/ncu_sum.py cs.csv ntile_pairs  file:lo-hi:name ..."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
ntp = float(sys.argv[2])
cur = None; curfile = None
agg = collections.Counter(); samp = collections.Counter()
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        curfile = r[1].split('/')[-1]; continue
    if len(r) > 7 and r[0].isdigit() and r[2] == "-":
        cur = (curfile, int(r[0]))
    elif len(r) > 7 and r[2].startswith('0x'):
        try: s = int(r[6]); x = int(r[7])
        except Exception: continue
        if cur is None: cur = ("?", 0)
        agg[cur] += x; samp[cur] += s
tot = sum(agg.values()); ts = sum(samp.values())
print("total warp instr per tile-pair %.0f" % (tot / ntp))
acc = 0
for spec in sys.argv[3:]:
    f, rng, name = spec.split(":")
    lo, hi = map(int, rng.split("-"))
    i = sum(v for (ff, l), v in agg.items() if ff and ff.startswith(f) and lo <= l <= hi)
    s = sum(v for (ff, l), v in samp.items() if ff and ff.startswith(f) and lo <= l <= hi)
    acc += i
    print("%-22s %7.0f instr/tile-pair (%5.1f%%)  samples %5.1f%%" % (name, i / ntp, 100.0 * i / tot, 100.0 * s / ts))
print("%-22s %7.0f" % ("other", (tot - acc) / ntp))
