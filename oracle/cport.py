"""ctypes wrapper of oracle/cpu_extract.c  --  TEST INFRASTRUCTURE ONLY (see that file's header)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libbk_oracle.so")


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def _load():
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "cpu_extract.c")):
        build()
    lib = C.CDLL(_SO)
    lib.cpu_extract.restype = C.c_int
    lib.cpu_extract.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t,
                                C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t),
                                C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_int,
                                C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_size_t]
    lib.cpu_expand.restype = C.c_int
    lib.cpu_expand.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
    lib.cpu_reverse_complement.restype = None
    lib.cpu_reverse_complement.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
    lib.cpu_bksyn_generate.restype = C.c_int
    lib.cpu_bksyn_generate.argtypes = [C.c_uint32, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int,
                                       C.c_uint64, C.c_uint64, C.c_size_t, C.c_void_p, C.c_int]
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def expand(pattern: str, k: int) -> str:
    buf = C.create_string_buffer(1 << 20)
    if lib().cpu_expand(pattern.encode("latin-1"), k, buf, 1 << 20) != 0:
        raise RuntimeError("expansion too large")
    return buf.value.decode("latin-1")


def reverse_complement(seq: bytes) -> bytes:
    out = C.create_string_buffer(len(seq) + 1)
    lib().cpu_reverse_complement(seq, len(seq), out)
    return out.raw[:len(seq)]


def _ptr(a):
    return a.ctypes.data if a is not None and a.size else None


def extract(fq1, fq2, pattern1, pattern2, max_error=1, rc_barcodes=False, skip_trimming=False,
            threads=1, want_match=False, off1=None, off2=None, out_bufs=None):
    """-> dict(out1, out2, n_in, n_out, match1, match2).  fq1/fq2: bytes or uint8 arrays."""
    a1 = np.frombuffer(fq1, dtype=np.uint8) if not isinstance(fq1, np.ndarray) else fq1
    paired = fq2 is not None
    a2 = (np.frombuffer(fq2, dtype=np.uint8) if not isinstance(fq2, np.ndarray) else fq2) if paired else None
    flags = (1 if paired else 0) | (2 if skip_trimming else 0) | (4 if rc_barcodes else 0)
    cap1 = a1.size * 2 + 4096
    cap2 = (a2.size * 2 + 4096) if paired else 0
    if out_bufs is not None:      # caller-owned, re-used buffers (bench: no page faults in the timing)
        o1, o2 = out_bufs
        cap1, cap2 = o1.size, (o2.size if paired else 0)
    else:
        o1 = np.empty(cap1, dtype=np.uint8)
        o2 = np.empty(max(cap2, 1), dtype=np.uint8)
    l1, l2, ni, no = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
    nmax = int(np.count_nonzero(a1 == 10)) // 4 + 2 if want_match else 0
    m1 = np.full(nmax, -2, dtype=np.int32) if want_match else None
    m2 = np.full(nmax, -2, dtype=np.int32) if want_match else None
    nrec = (len(off1) - 1) if off1 is not None else 0
    rc = lib().cpu_extract(pattern1.encode() if pattern1 is not None else None,
                           pattern2.encode() if pattern2 is not None else None, max_error, flags,
                           _ptr(a1), a1.size, _ptr(a2) if paired else None, a2.size if paired else 0,
                           o1.ctypes.data, cap1, C.byref(l1), o2.ctypes.data, cap2, C.byref(l2), threads,
                           C.byref(ni), C.byref(no), _ptr(m1), _ptr(m2),
                           off1.ctypes.data if off1 is not None else None,
                           off2.ctypes.data if off2 is not None else None, nrec)
    if rc != 0:
        raise RuntimeError("cpu_extract failed with code %d" % rc)
    if out_bufs is not None:
        return dict(out1_len=l1.value, out2_len=l2.value, n_in=ni.value, n_out=no.value)
    return dict(out1=o1[:l1.value].tobytes(), out2=o2[:l2.value].tobytes() if paired else None,
                n_in=ni.value, n_out=no.value,
                match1=m1[:ni.value] if want_match else None, match2=m2[:ni.value] if want_match else None)


def generate(cfg, mate: int, first_idx: int, n: int, seed: int, threads: int = 8) -> np.ndarray:
    """bksyn-v1 text of one mate as a uint8 array (cfg: oracle.bksyn.SynConfig)."""
    plant = cfg.plant1 if mate == 1 else cfg.plant2
    rl = 23 + 2 * cfg.read_len + 6
    out = np.empty(n * rl + 64, dtype=np.uint8)
    lit = plant.lit if plant else b""
    lib().cpu_bksyn_generate(cfg.read_len, lit, len(lit), plant.off if plant else 0,
                             plant.jitter if plant else 0, mate, seed, first_idx, n, out.ctypes.data, threads)
    return out[:n * rl]
