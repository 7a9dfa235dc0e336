"""Patterns that need a search: match_kernel + table (default) against the search inside the extract kernel
(BK_FLAG_MATCH_IN_KERNEL), device-timed on bksyn-v1 data.   python tools/match_speed.py [units]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from barkit_b200 import _lib
from tests.util import lib_syn_config

P = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
rl = 329
CASES = [  # (label, data config, paired, pattern1, pattern2, k)
    ("SE unanchored", "C4", False, "(?P<CB>[ATGCN]{16})atgccat", None, 1),
    ("SE anchored -r", "C1", False, "^(?P<UMI>[ATGCN]{12})", None, 1),
    ("PE one pattern, 3 variants", "C3", True, "^(?P<CB>[ATGCN]{14,16})atgccat(?P<UMI>[ATGCN]{12})", None, 1),
    ("PE one pattern, 9 variants", "C3", True, "^(?P<CB>[ATGCN]{14,16})atgccat(?P<UMI>[ATGCN]{10,12})", None, 1),
    ("PE one pattern, unanchored 3 variants", "C3", True, "(?P<CB>[ATGCN]{14,16})atgccat", None, 1),
    ("PE two anchored patterns", "C4", True, "^(?P<SB>[ATGCN]{8})gtcagtac(?P<UMI>[ATGCN]{12})", "^(?P<CB>[ATGCN]{16})", 1),
    ("PE C4", "C4", True, "^(?P<SB>[ATGCN]{8})gtcagtac(?P<UMI>[ATGCN]{12})", "(?P<CB>[ATGCN]{16})atgccat", 1),
    # the general form (match_kernel<VM>): no in-kernel variant exists
    ("PE one pattern, general anchored", "C3", True, "^(?P<CB>[ATGCN]{16})(atgccat)+(?P<UMI>[ATGCN]{12})", None, 1),
    ("PE one pattern, general unanchored", "C3", True, "(?P<CB>[ATGCN]{14,})(atgccat)+", None, 1),
    ("PE one pattern, general optional group", "C3", True, "^(?P<CB>[ATGCN]{16})(atgccat|gtcagtac)(?P<UMI>[ATGCN]{12})?", None, 1),
]
only = os.environ.get("BK_ONLY")
for label, cfgname, paired, p1, p2, k in CASES:
    if only and only not in label:
        continue
    forms = [(0, "match_kernel"), (_lib.FLAG_MATCH_IN_KERNEL, "in-kernel")]
    if p2:
        forms.append((_lib.FLAG_SPLIT_TWO_PATTERNS, "split"))
    if "general" in label:
        forms = forms[:1]
    for flags, fl in forms:
        rc = label.endswith("-r")
        ctx = _lib.Context(_lib.Program(p1, k), _lib.Program(p2, k) if p2 else None, paired=paired, rc_barcodes=rc,
                           extra_flags=flags)
        ms = []
        for mate in ((1, 2) if paired else (2 if cfgname == "C4" else 1,)):
            dt = ctx.dev_alloc(P * rl + 64); do = ctx.dev_alloc((P + 1) * 4)
            ctx.generate_device(lib_syn_config(cfgname), mate, 20240902, 0, P, dt, do)
            m = _lib.MateIn(); m.text, m.rec_off, m.text_len, m.max_rec_bytes = dt, do, P * rl, rl; ms.append(m)
        cap = P * (rl + 120)
        o1 = ctx.dev_alloc(cap); o2 = ctx.dev_alloc(cap) if paired else None
        for i in range(5):
            r = ctx.extract_device(P, ms[0], ms[1] if paired else None, o1, cap, o2, cap if paired else 0)
        print("%-40s %-12s %8.3f ms %8.1f M units/s  kept %.3f  launches %d" %
              (label, fl, r.kernel_ms, P / r.kernel_ms / 1e3, r.n_out / P, r.launches), flush=True)
        ctx.close()
