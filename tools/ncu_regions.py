"""Instruction / stall-sample share per source line of an ncu report.
  ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > cs.csv
  python tools/ncu_regions.py cs.csv [topN]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 50
cur = None; curfile = None
agg = collections.Counter(); samp = collections.Counter(); src = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        curfile = r[1].split('/')[-1]; continue
    if len(r) > 7 and r[0].isdigit() and r[2] == "-":
        cur = (curfile, int(r[0])); src[cur] = r[1]
    elif len(r) > 7 and r[2].startswith('0x'):
        try: s = int(r[6]); x = int(r[7])
        except Exception: continue
        if cur is None: cur = ("?", 0); src[cur] = "(no line info)"
        agg[cur] += x; samp[cur] += s
tot = sum(agg.values()) or 1; ts = sum(samp.values()) or 1
print("total samples", ts, "total warp instr", tot)
for k, v in sorted(agg.items(), key=lambda kv: (kv[0][0] or "", kv[0][1]))[:0]:
    pass
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1])[:N]:
    print("%-16s %4d  instr %5.2f%%  samp %5.2f%%  %s" % (f, l, 100 * v / tot, 100 * samp[(f, l)] / ts, src[(f, l)][:80]))
if len(sys.argv) > 3:   # per-file cumulative by line order
    print("---- by file/line")
    for (f, l), v in sorted(agg.items(), key=lambda kv: (kv[0][0] or "", kv[0][1])):
        if 100 * v / tot >= 0.15 or 100 * samp[(f, l)] / ts >= 0.15:
            print("%-16s %4d  instr %5.2f%%  samp %5.2f%%  %s" % (f, l, 100 * v / tot, 100 * samp[(f, l)] / ts, src[(f, l)][:80]))
