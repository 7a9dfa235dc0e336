"""Code size per source line / function of one kernel, from `nvdisasm -g -c` output (static: no GPU needed).
  cuobjdump -xelf all barkit_b200/libbarkit_b200.so; nvdisasm -g -c bk_ctx.sm_100a.cubin > k.sass
  python tools/sass_lines.py k.sass <kernel-substring> [top]
Also prints the opcode histogram of the kernel (UBLKCP = 1-D TMA bulk copy, SYNCS = mbarrier, UBLKPF = L2 prefetch)."""
import collections
import re
import sys

path, want = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
per_line = collections.Counter()
ops = collections.Counter()
cur = None
inside = False
name = None
for line in open(path):
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
    if m:
        name = m.group(1)
        inside = want in name
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        per_line[cur] += 1
        ops[m.group(1).split(".")[0]] += 1
tot = sum(per_line.values())
print("kernel *%s*: %d instructions = %.1f KB" % (want, tot, tot * 16 / 1024))
src = {}
for (f, l), n in per_line.most_common(top):
    if f not in src:
        try:
            src[f] = open("barkit_b200/csrc/" + f).read().split("\n")
        except OSError:
            src[f] = []
    text = src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ""
    print("%6.2f KB  %-16s %4d | %s" % (n * 16 / 1024, f, l, text))
print("opcodes:", ", ".join("%s %d" % kv for kv in ops.most_common()))
