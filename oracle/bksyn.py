"""`bksyn-v1`: counter-based synthetic FASTQ generator (normative Python/numpy statement).

TEST INFRASTRUCTURE ONLY (see oracle/barkit_oracle.py).  The same generator exists in C
(`oracle/cpu_extract.c`, `bksyn_*`) and in CUDA (`barkit_b200/csrc/bk_gen.cu`); tests require all
three to emit identical bytes.  The reference ships no fixtures or generator,
so this layout is ours; it only has to give R1 and R2 equal record byte lengths.

Record `idx`, mate `m` (1|2):
    key(field)  s0 = splitmix64(seed ^ splitmix64(8*idx + 4*(m-1) + field)),  word_j = splitmix64(s0 + j)
    head  = "bk.%012d %d:N:0:1" % (idx, m)                          (23 bytes)
    base i = "ACGT"[(word_{i//32} of field 0 >> 2*(i%32)) & 3]
             -> 'N' when byte (i%8) of word_{i//8} of field 2 is 0   (p = 1/256)
    qual i = 33 + (byte (i%8) of word_{i//8} of field 1) % 41
    plant (at most one per mate): literal `lit` at offset `off + (word_5 of field 3 % (jitter+1))`,
        u = word_0 of field 3 & 0xff:  u<205 exact | <230 1 sub | <243 2 subs | <250 3 subs |
        else nothing is planted.  Substitution t: pos = word_{1+t} % len, advanced by +1 (mod len)
        while already used; new base = "ACGT"[(orig + 1 + ((word_{1+t} >> 8) % 3)) % 4].
    text  = "@" head "\n" bases "\n+\n" quals "\n"                    (23 + 2L + 6 bytes)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

DEFAULT_SEED = 20240902
HEAD_LEN = 23
M64 = (1 << 64) - 1


@dataclass(frozen=True)
class Plant:
    lit: bytes
    off: int
    jitter: int = 0


@dataclass(frozen=True)
class SynConfig:
    """Planted structure for the two mates (None = plain random read)."""
    read_len: int = 150
    plant1: Optional[Plant] = None
    plant2: Optional[Plant] = None


# BASELINE.json configs
CONFIGS = {
    "C1": SynConfig(),
    "C2": SynConfig(plant2=Plant(b"ATGCCAT", 16)),
    "C3": SynConfig(plant1=Plant(b"ATGCCAT", 16)),
    "C4": SynConfig(plant1=Plant(b"GTCAGTAC", 8), plant2=Plant(b"ATGCCAT", 16, jitter=40)),
}


def _sm64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def _words(seed: int, idx: np.ndarray, mate: int, field: int, nwords: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        ctr = idx.astype(np.uint64) * np.uint64(8) + np.uint64(4 * (mate - 1) + field)
        s0 = _sm64(np.uint64(seed) ^ _sm64(ctr))
        j = np.arange(nwords, dtype=np.uint64)
        return _sm64(s0[:, None] + j[None, :])


def record_len(read_len: int) -> int:
    return HEAD_LEN + 2 * read_len + 6


def generate(cfg: SynConfig, mate: int, first_idx: int, n: int, seed: int = DEFAULT_SEED) -> bytes:
    """FASTQ text of records [first_idx, first_idx+n) of one mate."""
    L = cfg.read_len
    idx = np.arange(first_idx, first_idx + n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        w0 = _words(seed, idx, mate, 0, (L + 31) // 32)
        i = np.arange(L)
        code = (w0[:, i // 32] >> (np.uint64(2) * (i % 32).astype(np.uint64))) & np.uint64(3)
        bases = np.frombuffer(b"ACGT", dtype=np.uint8)[code.astype(np.intp)]
        w2 = _words(seed, idx, mate, 2, (L + 7) // 8)
        nmask = (w2[:, i // 8] >> (np.uint64(8) * (i % 8).astype(np.uint64))) & np.uint64(0xFF)
        bases = np.where(nmask == 0, np.uint8(ord("N")), bases).astype(np.uint8)
        w1 = _words(seed, idx, mate, 1, (L + 7) // 8)
        qb = (w1[:, i // 8] >> (np.uint64(8) * (i % 8).astype(np.uint64))) & np.uint64(0xFF)
        quals = (np.uint64(33) + qb % np.uint64(41)).astype(np.uint8)

    plant = cfg.plant1 if mate == 1 else cfg.plant2
    if plant is not None:
        w3 = _words(seed, idx, mate, 3, 6)
        lit = np.frombuffer(plant.lit, dtype=np.uint8)
        ll = len(lit)
        for r in range(n):  # small per-record loop; this generator is for test-sized inputs
            u = int(w3[r, 0]) & 0xFF
            if u >= 250:
                continue
            nsub = 0 if u < 205 else 1 if u < 230 else 2 if u < 243 else 3
            pos0 = plant.off + (int(w3[r, 5]) % (plant.jitter + 1))
            if pos0 + ll > L:
                continue
            win = lit.copy()
            used = []
            for t in range(nsub):
                w = int(w3[r, 1 + t])
                p = w % ll
                while p in used:
                    p = (p + 1) % ll
                used.append(p)
                orig = b"ACGT".index(bytes([plant.lit[p]]))
                win[p] = b"ACGT"[(orig + 1 + ((w >> 8) % 3)) % 4]
            bases[r, pos0:pos0 + ll] = win

    rl = record_len(L)
    out = np.empty((n, rl), dtype=np.uint8)
    out[:, 0] = ord("@")
    for r in range(n):
        out[r, 1:1 + HEAD_LEN] = np.frombuffer(
            b"bk.%012d %d:N:0:1" % (first_idx + r, mate), dtype=np.uint8)
    p = 1 + HEAD_LEN
    out[:, p] = ord("\n")
    out[:, p + 1:p + 1 + L] = bases
    p += 1 + L
    out[:, p] = ord("\n")
    out[:, p + 1] = ord("+")
    out[:, p + 2] = ord("\n")
    out[:, p + 3:p + 3 + L] = quals
    out[:, p + 3 + L] = ord("\n")
    return out.tobytes()


def generate_pair(cfg: SynConfig, first_idx: int, n: int,
                  seed: int = DEFAULT_SEED) -> Tuple[bytes, bytes]:
    return generate(cfg, 1, first_idx, n, seed), generate(cfg, 2, first_idx, n, seed)
