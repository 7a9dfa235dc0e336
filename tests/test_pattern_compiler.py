"""Host pattern front end (C ABI) against the reference's known-answer vectors and the oracle.

CPU-only: these tests make no compute calls on a device.  The lowered program is checked by
interpreting its text dump in Python and comparing match spans with the oracle's regex
(pattern.rs:225-230 semantics) on random reads.
"""
import random
import re
import zlib

import pytest

from barkit_b200 import _lib
from oracle import barkit_oracle as bo

# ---- the reference's own vectors, through the C ABI ------------------------------------------


@pytest.mark.parametrize("expected,text,max_error", [   # pattern.rs:244-249, :35-38
    (["."], "a", 1), (["A"], "a", 0), ([], "", 1),
    (["AAA.", "AA.A", "A.AA", ".AAA"], "AAAA", 1), (["..."], "AAA", 3), (["..."], "AAA", 4),
    (["ATG.", "AT.C", "A.GC", ".TGC"], "ATGC", 1),
])
def test_generate_sequences_with_pcr_errors(expected, text, max_error):
    assert _lib.sequence_with_errors(text, max_error) == expected


@pytest.mark.parametrize("expected,pattern,max_error", [   # pattern.rs:263-272, :88-91
    ("^(AA.|A.A|.AA)(?P<UMI>[ATGCN]{3})", "^aaa(?P<UMI>[ATGCN]{3})", 1),
    ("^(...)(?P<UMI>[ATGCN]{3})", "^aaa(?P<UMI>[ATGCN]{3})", 3),
    ("^(...)(?P<UMI>[ATGCN]{3})", "^aaa(?P<UMI>[ATGCN]{3})", 4),
    ("^((...))(?P<UMI>[ATGCN]{3})", "^(aaa)(?P<UMI>[ATGCN]{3})", 4),
    ("^(AA.|A.A|.AA)(?P<UMI>[ATGCN]{3})CCC", "^aaa(?P<UMI>[ATGCN]{3})CCC", 1),
    ("^(?P<UMI>[ATGCN]{3})", "^(?P<UMI>[ATGCN]{3})", 1),
    ("^(ATG.|AT.C|A.GC|.TGC)(?<UMI>[ATGCN]{12})", "^atgc(?<UMI>[ATGCN]{12})", 1),
])
def test_create_fuzzy(expected, pattern, max_error):
    assert _lib.expand(pattern, max_error) == expected


def test_barcode_types_in_pattern_order():   # pattern.rs:191-205, parse.rs:101
    p = _lib.Program("(?P<SB>[ATGC]{4})gattaca(?P<UMI>[ATGCN]{6})(?<CB>[ACGT]{2})", 1)
    assert p.barcode_types() == ["SB", "UMI", "CB"]


def test_error_messages_follow_error_rs():
    with pytest.raises(_lib.BarkitError) as e:     # error.rs:13-14
        _lib.Program("^(?P<XX>[ATGC]{4})", 1)
    assert e.value.status == _lib.BK_E_PATTERN
    assert str(e.value) == "Provided unexpected barcode capture group XX"
    with pytest.raises(_lib.BarkitError) as e:     # error.rs:11-12, pattern.rs:201-203
        _lib.Program("^atgc", 1)
    assert str(e.value) == "^(ATG.|AT.C|A.GC|.TGC) capture group not found in your pattern"
    assert str(e.value) == str(pytest.raises(bo.BarcodeCaptureGroupNotFound, bo.BarcodeRegex, "^atgc", 1).value)
    with pytest.raises(_lib.BarkitError) as e:     # duplicate names are a regex error (pattern.rs:182)
        _lib.Program("(?P<UMI>A)(?P<UMI>C)", 1)
    assert str(e.value).startswith("Regex error: ")
    with pytest.raises(_lib.BarkitError) as e:     # pattern.rs:56-58
        _lib.Program("^" + "a" * 65 + "(?P<UMI>[ATGC]{4})", 1)
    assert str(e.value) == "Failed to choose permutation mask"


@pytest.mark.parametrize("pattern", [
    "^(?P<UMI>[ATGCN]{2,4}+)", "^(?P<UMI>[ATGCN]{2,}*)",       # stacked repetition
    "^(?P<UMI>[ATGCN]{4,2})", "?(?P<UMI>[ATGCN]{4})",
    r"^(?P<UMI>\w{4})", "(?i)(?P<UMI>[ATGCN]{4})",
    "(?=A)(?P<UMI>[ATGCN]{4})", "(?P<UMI>[[:alpha:]]{4})",
    "^(?P<UMI>[^atg-z]{4})", "é(?P<UMI>A)",
    "^(?P<UMI>[ACGT]{4})(A?){2}",                  # a repeated group that can match the empty string
    "^(?P<UMI>[ACGT]{4})(A*|C)+",
    "^(?P<UMI>A)|(?P<XY>C)", "^(?P<UMI>A)|(?P<UMI>C)",
])
def test_unsupported_constructs_are_refused(pattern):
    with pytest.raises(_lib.BarkitError) as e:
        _lib.Program(pattern, 1)
    assert e.value.status in (_lib.BK_E_UNSUPPORTED, _lib.BK_E_PATTERN)


# ---- bounded variable repetition -> fixed-length variants in backtracking order -----------

def test_variable_repetition_lowers_to_variants_in_priority_order():
    p = _lib.Program("^(?P<UMI>[ATGCN]{8,10})atgccat", 1)
    assert p.nvariants() == 3
    assert p.info().total_len == 17 and list(p.info().cap_len)[:1] == [10]        # greedy: the longest first
    assert _lib.Program("^(?P<UMI>[ATGCN]{8,10}?)atgccat", 1).info().total_len == 15   # lazy: the shortest first
    assert _lib.Program("^(?P<UMI>[ATGCN]{12})", 1).nvariants() == 1
    assert _lib.Program("^N?(?P<CB>[ACGT]{4,6})T{0,1}(?P<UMI>[ACGT]{2})", 0).nvariants() == 2 * 3 * 2
    # the text the reference would hand to the regex crate keeps its quantifiers (pattern.rs:93-109)
    assert _lib.expand("^(?P<UMI>[ATGCN]{8,10})at", 1) == bo.BarcodePattern("^(?P<UMI>[ATGCN]{8,10})at", 1).get_pattern_with_errors()
    # [?{] inside a class and the ? of group syntax are not quantifiers
    assert _lib.Program("^(?P<UMI>[?{}*+]{3})(?:A)", 0).nvariants() == 1
    # variable repetition inside a group is fine as long as the group itself is taken once
    assert _lib.Program("(?P<UMI>[ACGT]{4})([AT]{1,2}G)", 0).nvariants() == 2
    assert _lib.Program("(?P<UMI>[ACGT]{4})(A?C){1}(GT){2}", 0).nvariants() == 2


def test_alternation_and_repeated_groups_unfold_into_variants():
    """`(a|b)`, `(X){m,n}`, `(a|b){n}`: one fixed-length program per unfolding, in the regex crate's decision order
    (branches in order; a greedy repetition prefers another round over stopping).  The expanded text the reference would
    compile keeps the construct."""
    assert _lib.Program("(?P<CB>[ACGT]{16})(atgccat|gattaca)", 1).nvariants() == 2
    assert _lib.Program("^(?P<UMI>AA|CCC)G", 0).info().total_len == 3            # first branch first
    p = _lib.Program("^(?P<UMI>[ACGT]{4})(AC){1,3}G", 0)
    assert p.nvariants() == 3 and p.info().total_len == 4 + 6 + 1                # greedy: three rounds first
    assert _lib.Program("^(?P<UMI>[ACGT]{4})(AC){1,3}?G", 0).info().total_len == 4 + 2 + 1
    assert _lib.Program("^(?P<UMI>[ACGT]{4})(A|C){2}", 0).nvariants() == 4
    assert _lib.Program("^(?P<UMI>[ACGT]{4})(A|CG){1,2}?T", 0).nvariants() == 6
    assert _lib.Program("^(?P<UMI>[ACGT]{4})(A?C){3}", 0).nvariants() == 8       # every copy has its own length
    assert _lib.Program("(?P<UMI>[ACGT]{4})([AT]{1,2}G){2}", 0).nvariants() == 4
    assert _lib.Program("(?P<UMI>[ACGT]{4})(A{1,2}?C){0}", 0).info().total_len == 4
    assert _lib.expand("(?P<CB>[ACGT]{4})(at|ga)", 1) == bo.BarcodePattern("(?P<CB>[ACGT]{4})(at|ga)", 1).get_pattern_with_errors()


def test_unbounded_repetition_lowers_to_one_parametric_program():
    """`* + {m,}` and {m,n} with more than 24 length combinations: ONE program laid out at the minimum counts, plus
    insertion points (bk_program.h: bk_ins) that the device matcher fills by backtracking."""
    p = _lib.Program("^N*(?P<CB>[ACGT]{4,})T+?(?P<UMI>[ACGT]{2})$", 0)
    assert p.nvariants() == 1 and p.info().total_len == 4 + 1 + 2
    ins = [l.split() for l in p.dump().splitlines() if l.startswith("ins ")]
    #          pos  max_extra  lazy  capmask
    assert [(int(f[1]), int(f[2]), int(f[3]), int(f[5])) for f in ins] == [(0, 65535, 0, 0), (4, 65535, 0, 1), (5, 65535, 1, 0)]
    assert p.barcode_types() == ["CB", "UMI"]
    # bounded but wide: the same form, with the bound kept
    p = _lib.Program("(?P<UMI>[ACGT]{1,40})gattaca", 1)
    ins = [l.split() for l in p.dump().splitlines() if l.startswith("ins ")]
    assert p.nvariants() == 1 and [(int(f[1]), int(f[2])) for f in ins] == [(1, 39)]
    # up to 24 combinations stay fixed-length variants
    assert _lib.Program("(?P<UMI>[ACGT]{1,24})gattaca", 1).nvariants() == 24
    # the text the reference would hand to the regex crate keeps its quantifiers (pattern.rs:93-109)
    assert _lib.expand("^(?P<UMI>[ATGCN]+)at", 1) == bo.BarcodePattern("^(?P<UMI>[ATGCN]+)at", 1).get_pattern_with_errors()


# ---- the general form: a backtracking program (bk_program.h: bk_vmop) ------------------------------------------------


class DumpedVm:
    """Interpreter of a general-form program's text dump: the statement of what match_vm (bk_kernels.cuh) does."""

    def __init__(self, text):
        self.code, self.cls, self.caps = [], {}, []
        for line in text.strip().split("\n"):
            f = line.split()
            if f[0] == "anchor":
                self.a_start = int(f[1])
            elif f[0] == "total":
                self.tmin = int(f[1])
            elif f[0] == "vmclass":
                self.cls[int(f[1])] = int.from_bytes(bytes.fromhex(f[2]), "little")
            elif f[0] == "vm":
                self.code.append((f[2], int(f[3]), int(f[4]), int(f[5]), bytes.fromhex(f[6]) if len(f) > 6 else b""))
            elif f[0] == "cap":
                self.caps.append(f[1])

    def inclass(self, c, k):
        return c < 0x80 and (self.cls[k] >> c) & 1

    def search(self, seq):
        if len(seq) < self.tmin:
            return None
        for s in range((0 if self.a_start else len(seq) - self.tmin) + 1):
            r = self.run(seq, s)
            if r is not None:
                return r
        return None

    def run(self, seq, s):
        L, pc, pos, cap, stk = len(seq), 0, s, [-1] * 6, []
        while True:
            op, k, a, b, lit = self.code[pc]
            ok = True
            if op == "cls":
                ok = pos + b <= L and all(self.inclass(seq[pos + i], a) for i in range(b))
                pos += b
                pc += 1
            elif op == "lit":
                ok = pos + b <= L
                if ok:
                    bad = [seq[pos + i] for i in range(b) if seq[pos + i] != lit[i]]
                    ok = len(bad) <= k and not any(c >= 0x80 or c == 10 for c in bad)
                pos += b
                pc += 1
            elif op == "split":
                stk.append(("b", b, pos))
                pc = a
            elif op == "jmp":
                pc = a
            elif op == "save":
                stk.append(("r", a, cap[a]))
                cap[a] = pos
                pc += 1
            elif op == "bol":
                ok = pos == 0
                pc += 1
            elif op == "eol":
                ok = pos == L
                pc += 1
            elif op == "star":
                pc += 1
                if k == 0:      # greedy: the whole run, then a byte less per visit
                    n = 0
                    while n < b and pos + n < L and self.inclass(seq[pos + n], a):
                        n += 1
                    if n:
                        stk.append(("g", pc, pos + n - 1, n - 1))
                    pos += n
                elif b:         # lazy: nothing, then a byte more per visit
                    stk.append(("l", pc - 1, pos, b))
            else:
                return s, pos, cap
            if ok:
                continue
            resumed = False
            while stk and not resumed:
                e = stk.pop()
                if e[0] == "r":
                    cap[e[1]] = e[2]
                elif e[0] == "b":
                    pc, pos, resumed = e[1], e[2], True
                elif e[0] == "g":
                    pc, pos, resumed = e[1], e[2], True
                    if e[3]:
                        stk.append(("g", e[1], e[2] - 1, e[3] - 1))
                elif e[2] < L and self.inclass(seq[e[2]], self.code[e[1]][2]):
                    pc, pos, resumed = e[1] + 1, e[2] + 1, True
                    if e[3] > 1:
                        stk.append(("l", e[1], e[2] + 1, e[3] - 1))
            if not resumed:
                return None


GENERAL = [
    "^(?P<UMI>([ATGCN])+)", "^(?P<UMI>[ATGCN]{4})+", "^(?P<UMI>[ATGCN]+)(A*C){2}", "^(?P<UMI>[ATGCN]+)a*",
    "^(?P<UMI>A*C*G*T*N*)", "^((?P<UMI>[ATGCN]{4})A){1,2}", "^(?P<UMI>[ATGCN]{4})?", "(?P<UMI>[ATGCN]{4})^",
    "((?P<UMI>[ATGCN]{4})){2}", "^(?P<UMI>A)|(?P<CB>C)", "^(A|(?P<UMI>C))T", "^(?P<UMI>[ACGT]+)(A|C)",
    "^(?P<UMI>[ACGT]{4})(A|C|G|T|N){3}", "(?P<UMI>[ACGT]{4})(AC)+G", "(?P<CB>[ACGT]{3,})(atgc)+(?P<UMI>[ACGT]{2})",
    "(?P<UMI>A+?)(C|GT)*?T$", "^(?P<CB>[ACGT]{2})((?P<UMI>A[CG])|T)+", "(?P<UMI>[ACGT]{4})(atgc|gatt)+",
    "^(?P<UMI>(AC|A)+)C", "(?P<UMI>(A|AC)(C|CA)*)T", "(?P<SB>[ACGT]{2})(gattaca){1,}?(?P<UMI>[ACGT]{3})",
    "(?P<UMI>[ACGT]{3})(acgt){2,}|^(?P<CB>N+)",
]


@pytest.mark.parametrize("pattern", GENERAL)
def test_general_form_equals_regex(pattern):
    """What the fixed, variant and parametric forms refuse compiles to a backtracking program; interpreted in Python, it
    must give the oracle regex's match span and capture spans (a group that took no part in the match: -1, -1)."""
    rnd = random.Random(zlib.crc32(pattern.encode()))
    for k in (0, 1):
        prog = _lib.Program(pattern, k)
        assert prog.nvariants() == 1 and "vm 0 " in prog.dump(), pattern
        vm = DumpedVm(prog.dump())
        rx = bo.BarcodeRegex(pattern, k)
        assert prog.barcode_types() == rx.get_barcode_types()
        lits = re.findall(r"[atgcn]{3,}", pattern)
        matched = 0
        for _ in range(300):
            L = rnd.choice([0, 3, 8, 17, 23, 35, 36, 60])
            alphabet = rnd.choice([b"ACGTACGTACGTN", b"AC", b"ACGT", b"AAAC"])
            seq = bytearray(rnd.choice(alphabet) for _ in range(L))
            if lits and L > 30 and rnd.random() < 0.7:
                lit = bytearray(rnd.choice(lits).upper().encode())
                for _m in range(rnd.choice([0, 0, 1, 1, 2, 3])):
                    lit[rnd.randrange(len(lit))] = rnd.choice(b"ACGTN a\r")
                at = rnd.randrange(0, L - len(lit))
                seq[at:at + len(lit)] = lit
            seq = bytes(seq)
            m = rx.regex.search(seq)
            r = vm.search(seq)
            assert (r is None) == (m is None), (pattern, k, seq)
            if m:
                matched += 1
                assert r[:2] == m.span(0), (pattern, k, seq)
                for c, name in enumerate(vm.caps):
                    assert (r[2][2 * c], r[2][2 * c + 1]) == m.span(name), (pattern, k, seq, name)
        assert matched or pattern.endswith("^")


# ---- expansion text == oracle on generated patterns ---------------------------------------------

_PIECES = ["^", "$", "atgc", "n", "gattaca", "ATGC", "N", "[ATGCN]", "[atgc]", "{3}", "{12}", "(", ")",
           "(?P<UMI>", "(?<CB>", "(?P<SB>", ".", " ", "_", "x", "[", "]", "\\", "d", "-", "|", "acgt]",
           "[^acg", "9", "rysw", "kmbdhvn"]


def test_expand_matches_oracle_on_random_pattern_soup():
    rnd = random.Random(7)
    for _ in range(1500):
        pat = "".join(rnd.choice(_PIECES) for _ in range(rnd.randint(1, 8)))
        for k in (0, 1, 2, 9):
            assert _lib.expand(pat, k) == bo.BarcodePattern(pat, k).get_pattern_with_errors(), (pat, k)


# ---- lowered program == oracle regex on random reads -----------------------------------------------


class DumpedProgram:
    def __init__(self, text):
        self.ops, self.caps = [], []
        for line in text.strip().split("\n"):
            f = line.split()
            if f[0] == "anchor":
                self.a_start, self.a_end = int(f[1]), int(f[2])
            elif f[0] == "total":
                self.total = int(f[1])
            elif f[0] == "lit":
                self.ops.append(("lit", int(f[1]), int(f[2]), int(f[3]), bytes.fromhex(f[4]) if len(f) > 4 else b""))
            elif f[0] == "class":
                bm = int.from_bytes(bytes.fromhex(f[3]), "little")
                self.ops.append(("class", int(f[1]), int(f[2]), bm))
            elif f[0] == "cap":
                self.caps.append((f[1], int(f[2]), int(f[3])))

    def match_at(self, seq, s):
        for op in self.ops:
            if op[0] == "lit":
                _, off, ln, k, lit = op
                mism = 0
                for i in range(ln):
                    c = seq[s + off + i]
                    if c != lit[i]:
                        if c >= 0x80 or c == 10:
                            return False
                        mism += 1
                        if mism > k:
                            return False
            else:
                _, off, ln, bm = op
                for i in range(ln):
                    c = seq[s + off + i]
                    if c >= 0x80 or not (bm >> c) & 1:
                        return False
        return True

    def search(self, seq):
        L, T = len(seq), self.total
        if L < T:
            return -1
        lo, hi = 0, L - T
        if self.a_start:
            hi = 0
        if self.a_end:
            if self.a_start and L != T:
                return -1
            lo = hi = L - T
        for s in range(lo, hi + 1):
            if self.match_at(seq, s):
                return s
        return -1


SUPPORTED = [
    ("^(?P<UMI>[ATGCN]{12})", [0, 1]),
    ("^(?P<CB>[ATGCN]{16})atgccat", [0, 1, 2, 7, 9]),
    ("^(?P<CB>[ATGCN]{16})atgccat(?P<UMI>[ATGCN]{12})", [0, 1, 2]),
    ("^(?P<SB>[ATGCN]{8})gtcagtac(?P<UMI>[ATGCN]{12})", [0, 1, 2]),
    ("(?P<CB>[ATGCN]{16})atgccat", [0, 1, 2]),
    ("(?P<SB>[ATGC]{4})gattaca(?P<UMI>[ATGCN]{6})", [0, 1, 2]),
    ("(?P<UMI>[ATGCN]{6})acgt$", [0, 1]),
    ("^(?P<UMI>[ATGCN]{6})acgt$", [1]),
    ("^n{3}(?P<UMI>[ATGCN]{4})", [0, 1]),
    ("^(atg){2}(?P<UMI>[ACGT]{3})TT", [0, 1]),
    ("(?P<CB>atgc)(?:N|)".replace("(?:N|)", "") + "[^N]{2}(?P<UMI>.{3})", [0, 1]),
    ("ATGCatgc(?P<UMI>[ACGT]{2})", [1]),
    ("(?P<UMI>[A-C]{3}[G-T])acg", [1]),
    ("^(?P<UMI>)ACG", [0]),
    ("^(?P<UMI>[ATGCN]{0})", [1]),
    ("(?<CB>[ATGC]{3}) (?P<UMI>[ATGC]{2})", [1]),
]


@pytest.mark.parametrize("pattern,ks", SUPPORTED)
def test_lowered_program_equals_regex(pattern, ks):
    rnd = random.Random(zlib.crc32(pattern.encode()))
    for k in ks:
        prog = _lib.Program(pattern, k)
        dp = DumpedProgram(prog.dump())
        rx = bo.BarcodeRegex(pattern, k)
        assert prog.barcode_types() == rx.get_barcode_types()
        lits = re.findall(r"[atgcn]{3,}", pattern)
        for _ in range(400):
            L = rnd.choice([0, 3, 8, 17, 23, 35, 36, 60])
            seq = bytearray(rnd.choice(b"ACGTACGTACGTN") for _ in range(L))
            if lits and L > 30 and rnd.random() < 0.7:          # plant a (mutated) adapter
                lit = bytearray(rnd.choice(lits).upper().encode())
                for _m in range(rnd.choice([0, 0, 1, 1, 2, 3])):
                    lit[rnd.randrange(len(lit))] = rnd.choice(b"ACGTN a\r")
                p = rnd.randrange(0, L - len(lit))
                seq[p:p + len(lit)] = lit
            seq = bytes(seq)
            m = rx.regex.search(seq)
            s = dp.search(seq)
            assert s == (m.start(0) if m else -1), (pattern, k, seq)
            if m:
                assert m.end(0) - m.start(0) == dp.total
                for name, off, ln in dp.caps:
                    assert (m.start(name), m.end(name)) == (s + off, s + off + ln)


def test_every_declared_symbol_is_exported(repo_root):
    import os
    hdr = open(os.path.join(repo_root, "include", "barkit_b200.h")).read()
    declared = set(re.findall(r"\b(bk_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(_lib.lib, name)


def test_tokenize_host():
    fq = b"@a x\nACGT\n+\nIIII\n@b\r\nAC\r\n+b\r\nII\r\n@c\nA\n+\nI"
    off, consumed, maxrec = _lib.tokenize(fq)
    assert list(off) == [0, 17, 33, len(fq)]
    assert consumed == len(fq) and maxrec == 17
    off, consumed, _ = _lib.tokenize(fq, final=False)          # incomplete tail is left for the next chunk
    assert list(off) == [0, 17, 33] and consumed == 33
    for bad in (b"a\nACGT\n+\nIIII\n", b"@a\nACGT\n-\nIIII\n", b"@a\nACGT\n+\nIII\n", b"@a\nACGT\n+\n"):
        with pytest.raises(_lib.BarkitError):
            _lib.tokenize(bad)
    assert [len(r.seq) for r in bo.read_fastq(fq)] == [4, 2, 1]
