"""Shared helpers for the parity tests (inputs only; no product logic)."""
import random

from oracle import bksyn

# BASELINE.json configs -> (pattern1, pattern2, paired)   
CONFIG_PATTERNS = {
    "C1": ("^(?P<UMI>[ATGCN]{12})", None, False),
    "C2": (None, "^(?P<CB>[ATGCN]{16})atgccat", True),
    "C3": ("^(?P<CB>[ATGCN]{16})atgccat(?P<UMI>[ATGCN]{12})", None, True),
    "C4": ("^(?P<SB>[ATGCN]{8})gtcagtac(?P<UMI>[ATGCN]{12})", "(?P<CB>[ATGCN]{16})atgccat", True),
}


def lib_syn_config(name_or_cfg):
    """oracle/bksyn.SynConfig -> the C ABI's bk_syn_config."""
    from barkit_b200 import _lib
    cfg = bksyn.CONFIGS[name_or_cfg] if isinstance(name_or_cfg, str) else name_or_cfg

    def plant(p):
        return None if p is None else (p.lit, p.off, p.jitter)
    return _lib.syn_config(cfg.read_len, plant(cfg.plant1), plant(cfg.plant2))


def ragged_fastq(rnd: random.Random, n: int, mate: int = 1, plant=None, min_len=0, max_len=200,
                 crlf=False, plus_comment=False, final_newline=True, head_lens=None):
    """Random FASTQ text with variable read and header lengths.  `plant` = list of literal bytes
    sprinkled (sometimes mutated) at random positions.  Headers depend only on the record index so
    that two mates generated with the same `head_lens` pair up."""
    eol = b"\r\n" if crlf else b"\n"
    recs = []
    for i in range(n):
        L = rnd.randint(min_len, max_len)
        seq = bytearray(rnd.choice(b"ACGTACGTACGTACGTN") for _ in range(L))
        if rnd.random() < 0.03 and L:
            seq[rnd.randrange(L)] = rnd.choice(b"acgtnRYKM*.-")
        if plant and L > 12:
            for _ in range(rnd.choice([0, 1, 1, 2])):
                lit = bytearray(rnd.choice(plant))
                for _m in range(rnd.choice([0, 0, 0, 1, 1, 2, 3])):
                    lit[rnd.randrange(len(lit))] = rnd.choice(b"ACGTN")
                if len(lit) < L:
                    p = rnd.choice([0, 0, rnd.randrange(0, L - len(lit) + 1)])
                    p = min(p, L - len(lit))
                    seq[p:p + len(lit)] = lit
        qual = bytes(rnd.randint(33, 74) for _ in range(L))
        hl = head_lens[i] if head_lens else rnd.choice([1, 5, 23, 40, 77])
        head = (b"r%d" % i + b" %d:N:0:" % mate + b"x" * 200)[:hl] if hl > 0 else b""
        plus = b"+" + (head if plus_comment and rnd.random() < 0.5 else b"")
        recs.append(b"@" + head + eol + bytes(seq) + eol + plus + eol + qual + eol)
    text = b"".join(recs)
    if not final_newline and text:
        text = text[:-len(eol)]
    return text
