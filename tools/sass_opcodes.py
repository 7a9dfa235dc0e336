"""SASS opcode histogram of every kernel in the built library (what profiles/<tag>_sass_opcodes.txt holds).
  python tools/sass_opcodes.py > profiles/r2_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "barkit_b200", "libbarkit_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
print("# SASS opcode histograms of the kernels in barkit_b200/libbarkit_b200.so (cuobjdump -sass; tools/sass_opcodes.py)")
print("# UBLKCP = 1-D TMA bulk copy (cp.async.bulk), UBLKPF = L2 bulk prefetch, SYNCS = mbarrier, UTMACMDFLUSH = bulk-group "
      "commit/wait; no tensor-core opcodes: the path is HBM-bound byte work")
print("# extract_kernelILb<PAIRED>ELb<GENERAL>ELb<VAR>ELi<ring depth>ELb<TABLE>; match_kernelILb<VM>")
for part in re.split(r"\n\s*Function : ", txt)[1:]:
    name = part.split("\n", 1)[0].strip()
    ops = collections.Counter(m.group(1) for m in re.finditer(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", part))
    n = sum(ops.values())
    print("\nkernel %s: %d instructions = %.1f KB" % (name, n, n * 16 / 1024))
    print("opcodes: " + ", ".join("%s %d" % kv for kv in ops.most_common()))
