"""ctypes binding of the C ABI in include/barkit_b200.h (tests / bench / smoke use this; the
product is the shared library and the `barkit` CLI, not this file).

The library is built in-tree (barkit_b200/build.py).  Importing fails loudly when it is missing:
there is no Python or CPU implementation of the extract path to fall back to.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BK_LIB") or os.path.join(_HERE, "libbarkit_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "barkit_b200: %s is missing. Build it with `python -m barkit_b200.build` "
        "(or __graft_entry__.build()); there is no CPU fallback." % LIB_PATH)

lib = C.CDLL(LIB_PATH)

BK_OK = 0
BK_E_PATTERN, BK_E_UNSUPPORTED, BK_E_CUDA, BK_E_CAPACITY, BK_E_IO, BK_E_ARG, BK_E_FORMAT = \
    -1, -2, -3, -4, -5, -6, -7
FLAG_PAIRED, FLAG_SKIP_TRIM, FLAG_RC = 1, 2, 4
FLAG_DEVICE_BOTH_MATES, FLAG_HOST_MATE_PINNED, FLAG_VERBOSE, FLAG_STAGE_BOTH_MATES, FLAG_MATCH_IN_KERNEL, FLAG_SPLIT_TWO_PATTERNS = 8, 16, 32, 64, 128, 256

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)


class MateIn(C.Structure):
    _fields_ = [("text", C.c_void_p), ("rec_off", C.c_void_p), ("text_len", C.c_uint64),
                ("max_rec_bytes", C.c_uint32), ("reserved", C.c_uint32)]


class Result(C.Structure):
    _fields_ = [("out1_len", C.c_uint64), ("out2_len", C.c_uint64), ("n_in", C.c_uint64),
                ("n_out", C.c_uint64), ("n_match1", C.c_uint64), ("n_match2", C.c_uint64),
                ("kernel_ms", C.c_float), ("launches", C.c_uint32),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]


class ProgramInfo(C.Structure):
    _fields_ = [("total_len", C.c_uint32), ("anchor_start", C.c_uint32), ("anchor_end", C.c_uint32),
                ("nops", C.c_uint32), ("ncaps", C.c_uint32), ("hdr_growth", C.c_uint32),
                ("cap_off", C.c_uint32 * 3), ("cap_len", C.c_uint32 * 3)]


class SynConfig(C.Structure):
    _fields_ = [("read_len", C.c_uint32), ("plant_len", C.c_uint32 * 2), ("plant_off", C.c_uint32 * 2),
                ("plant_jitter", C.c_uint32 * 2), ("plant_lit", (C.c_char * 32) * 2)]


class RunArgs(C.Structure):
    _fields_ = [("fq1", C.c_char_p), ("fq2", C.c_char_p), ("pattern1", C.c_char_p),
                ("pattern2", C.c_char_p), ("out_fq1", C.c_char_p), ("out_fq2", C.c_char_p),
                ("max_memory_mb", C.c_size_t), ("threads", C.c_size_t), ("rc_barcodes", C.c_int),
                ("skip_trimming", C.c_int), ("max_error", C.c_size_t), ("compression", C.c_int),
                ("quiet", C.c_int), ("force", C.c_int), ("n_gpus", C.c_int), ("verbose", C.c_int),
                ("ctx_flags", C.c_uint32)]


# every symbol include/barkit_b200.h declares: (restype, argtypes)
SYMBOLS = {
    "bk_sequence_with_errors": (C.c_int, [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "bk_expand": (C.c_int, [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]),
    "bk_compile": (C.c_void_p, [C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.c_char_p, C.c_size_t]),
    "bk_program_destroy": (None, [C.c_void_p]),
    "bk_program_ncaps": (C.c_int, [C.c_void_p]),
    "bk_program_cap_name": (C.c_char_p, [C.c_void_p, C.c_int]),
    "bk_program_get_info": (C.c_int, [C.c_void_p, C.POINTER(ProgramInfo)]),
    "bk_program_dump": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "bk_tokenize": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_uint32, u32p,
                              C.POINTER(C.c_uint64), u32p]),
    "bk_device_count": (C.c_int, []),
    "bk_ctx_create": (C.c_void_p, [C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_int),
                                   C.c_char_p, C.c_size_t]),
    "bk_ctx_destroy": (None, [C.c_void_p]),
    "bk_last_error": (C.c_char_p, [C.c_void_p]),
    "bk_host_alloc": (C.c_void_p, [C.c_size_t]),
    "bk_host_free": (None, [C.c_void_p]),
    "bk_extract_host": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(MateIn), C.POINTER(MateIn),
                                  C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(Result)]),
    "bk_ctx_nslots": (C.c_int, [C.c_void_p]),
    "bk_submit": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32, C.POINTER(MateIn), C.POINTER(MateIn)]),
    "bk_wait": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                          C.POINTER(Result)]),
    "bk_submit_text": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                 C.c_uint32]),
    "bk_out_bound": (C.c_uint64, [C.c_void_p, C.c_int, C.c_uint32, C.c_uint64]),
    "bk_dev_alloc": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]),
    "bk_dev_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bk_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    "bk_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
    "bk_extract_device": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(MateIn), C.POINTER(MateIn),
                                    C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p,
                                    C.c_void_p, C.c_int, C.POINTER(Result)]),
    "bk_tokenize_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_uint32, u32p,
                                     C.POINTER(C.c_uint64), u32p]),
    "bk_ctx_set_host_threads": (C.c_int, [C.c_void_p, C.c_int]),
    "bk_warmup": (C.c_int, [C.c_int]),
    "bk_program_nvariants": (C.c_int, [C.c_void_p]),
    "bk_debug_events": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]),
    "bk_reverse_complement": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "bk_generate_device": (C.c_int, [C.c_void_p, C.POINTER(SynConfig), C.c_int, C.c_uint64, C.c_uint64,
                                     C.c_uint32, C.c_void_p, C.c_void_p]),
    "bk_bus_probe": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                               C.POINTER(C.c_double)]),
    "bk_host_mem_probe": (C.c_int, [C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "bk_run": (C.c_int, [C.POINTER(RunArgs), C.c_char_p, C.c_size_t]),
}

for _name, (_res, _args) in SYMBOLS.items():
    _f = getattr(lib, _name)  # AttributeError = symbol missing = loud failure
    _f.restype = _res
    _f.argtypes = _args


class BarkitError(Exception):
    def __init__(self, status, message):
        super().__init__(message)
        self.status = status
        self.message = message

    def __str__(self):
        return self.message


def device_count() -> int:
    return lib.bk_device_count()


def sequence_with_errors(sequence: str, max_error: int):
    cap = 1 << 16
    while True:
        buf = C.create_string_buffer(cap)
        n = C.c_size_t(0)
        rc = lib.bk_sequence_with_errors(sequence.encode(), max_error, buf, cap, C.byref(n))
        if rc == BK_E_CAPACITY and cap < (1 << 28):
            cap *= 16
            continue
        if rc != BK_OK:
            raise BarkitError(rc, buf.value.decode())
        return buf.value.decode().split("\n") if n.value else []


def expand(pattern: str, max_error: int) -> str:
    cap = 1 << 16
    while True:
        buf = C.create_string_buffer(cap)
        rc = lib.bk_expand(pattern.encode("latin-1"), max_error, buf, cap)
        if rc == BK_E_CAPACITY and cap < (1 << 28):
            cap *= 16
            continue
        if rc != BK_OK:
            raise BarkitError(rc, buf.value.decode())
        return buf.value.decode("latin-1")


class Program:
    """bk_program handle = the reference's BarcodeRegex (pattern.rs:161-235)."""

    def __init__(self, pattern: str, max_error: int):
        err = C.create_string_buffer(1024)
        st = C.c_int(0)
        self.handle = lib.bk_compile(pattern.encode("latin-1"), max_error, C.byref(st), err, 1024)
        if not self.handle:
            raise BarkitError(st.value, err.value.decode("latin-1"))
        self.pattern = pattern
        self.max_error = max_error

    def barcode_types(self):
        return [lib.bk_program_cap_name(self.handle, i).decode()
                for i in range(lib.bk_program_ncaps(self.handle))]

    def nvariants(self) -> int:
        """Fixed-length programs the pattern lowers to (1 unless it has {m,n} / ? repetition)."""
        return lib.bk_program_nvariants(self.handle)

    def info(self) -> ProgramInfo:
        info = ProgramInfo()
        lib.bk_program_get_info(self.handle, C.byref(info))
        return info

    def dump(self) -> str:
        buf = C.create_string_buffer(1 << 16)
        rc = lib.bk_program_dump(self.handle, buf, 1 << 16)
        if rc != BK_OK:
            raise BarkitError(rc, "dump failed")
        return buf.value.decode()

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            lib.bk_program_destroy(h)
            self.handle = None


def tokenize(text, final: bool = True, cap_records: int | None = None):
    """-> (rec_off uint32[n+1], consumed, max_rec_bytes)"""
    if len(text) == 0:
        return np.zeros(1, dtype=np.uint32), 0, 0
    arr = np.frombuffer(text, dtype=np.uint8)
    if cap_records is None:
        cap_records = int(np.count_nonzero(arr == 10)) // 4 + 2
    off = np.zeros(cap_records + 1, dtype=np.uint32)
    n = C.c_uint32(0)
    consumed = C.c_uint64(0)
    maxrec = C.c_uint32(0)
    rc = lib.bk_tokenize(arr.ctypes.data, arr.size, 1 if final else 0, off.ctypes.data, cap_records,
                         C.byref(n), C.byref(consumed), C.byref(maxrec))
    if rc != BK_OK:
        raise BarkitError(rc, "malformed FASTQ text")
    return off[:n.value + 1].copy(), consumed.value, maxrec.value


class Context:
    """bk_ctx handle: one GPU, compiled programs, run.rs flags."""

    def __init__(self, p1: Program | None, p2: Program | None = None, paired: bool = False,
                 skip_trimming: bool = False, rc_barcodes: bool = False, device: int = 0,
                 extra_flags: int = 0, host_threads: int | None = None):
        flags = (FLAG_PAIRED if paired else 0) | (FLAG_SKIP_TRIM if skip_trimming else 0) | \
                (FLAG_RC if rc_barcodes else 0) | extra_flags
        err = C.create_string_buffer(1024)
        st = C.c_int(0)
        self.handle = lib.bk_ctx_create(device, p1.handle if p1 else None, p2.handle if p2 else None,
                                        flags, C.byref(st), err, 1024)
        if not self.handle:
            raise BarkitError(st.value, err.value.decode())
        self.paired = paired
        self._keep = (p1, p2)
        if host_threads is not None:
            self._check(lib.bk_ctx_set_host_threads(self.handle, host_threads))

    def close(self):
        if self.handle:
            lib.bk_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        self.close()

    def _check(self, rc):
        if rc != BK_OK:
            raise BarkitError(rc, lib.bk_last_error(self.handle).decode())

    @staticmethod
    def _mate(text_arr, off_arr, maxrec):
        m = MateIn()
        m.text = text_arr.ctypes.data
        m.rec_off = off_arr.ctypes.data
        m.text_len = int(off_arr[-1])
        m.max_rec_bytes = max(int(maxrec), 1)
        return m

    def out_bound(self, mate, n, text_len):
        return lib.bk_out_bound(self.handle, mate, n, text_len)

    def extract_text(self, fq1: bytes, fq2: bytes | None = None):
        """FASTQ text in (host) -> FASTQ text out (host), through bk_tokenize + bk_extract_host."""
        texts = [fq1] + ([fq2] if fq2 is not None else [])
        mates, keep, n = [], [], None
        for t in texts:
            off, consumed, maxrec = tokenize(t)
            if consumed != len(t):
                raise BarkitError(BK_E_FORMAT, "trailing bytes after the last record")
            if n is None:
                n = len(off) - 1
            n = min(n, len(off) - 1)  # run.rs:172-173 zip truncates to the shorter side
            arr = np.zeros(len(t) + 64, dtype=np.uint8)
            arr[:len(t)] = np.frombuffer(t, dtype=np.uint8)
            keep.append((arr, off))
            mates.append(self._mate(arr, off, maxrec))
        for i, m in enumerate(mates):
            m.text_len = int(keep[i][1][n])
        outs = [np.zeros(int(self.out_bound(i + 1, n, mates[i].text_len)) + 16, dtype=np.uint8)
                for i in range(len(mates))]
        res = Result()
        rc = lib.bk_extract_host(self.handle, n, C.byref(mates[0]),
                                 C.byref(mates[1]) if len(mates) > 1 else None,
                                 outs[0].ctypes.data, outs[0].size,
                                 outs[1].ctypes.data if len(outs) > 1 else None,
                                 outs[1].size if len(outs) > 1 else 0, C.byref(res))
        self._check(rc)
        o1 = outs[0][:res.out1_len].tobytes()
        if len(mates) > 1:
            return o1, outs[1][:res.out2_len].tobytes(), res
        return o1, res

    # ---- device-resident helpers ----
    def dev_alloc(self, nbytes: int) -> int:
        p = C.c_void_p(0)
        self._check(lib.bk_dev_alloc(self.handle, nbytes, C.byref(p)))
        return p.value

    def dev_free(self, ptr: int):
        self._check(lib.bk_dev_free(self.handle, ptr))

    def h2d(self, dptr: int, arr: np.ndarray):
        self._check(lib.bk_memcpy_h2d(self.handle, dptr, arr.ctypes.data, arr.nbytes))

    def d2h(self, dptr: int, nbytes: int, dtype=np.uint8) -> np.ndarray:
        out = np.empty(nbytes // np.dtype(dtype).itemsize, dtype=dtype)
        if nbytes:
            self._check(lib.bk_memcpy_d2h(self.handle, out.ctypes.data, dptr, nbytes))
        return out

    def tokenize_device(self, text: bytes, final: bool = True, cap_records: int | None = None):
        """bk_tokenize_device on a host string (uploaded first): -> (rc, rec_off uint32[n+1], consumed, max_rec_bytes)"""
        arr = np.zeros(len(text) + 64, np.uint8)
        arr[:len(text)] = np.frombuffer(text, np.uint8)
        if cap_records is None:
            cap_records = int(np.count_nonzero(arr == 10)) // 4 + 2
        d_text = self.dev_alloc(arr.size)
        d_off = self.dev_alloc((cap_records + 1) * 4)
        self.h2d(d_text, arr)
        n, consumed, maxrec = C.c_uint32(0), C.c_uint64(0), C.c_uint32(0)
        rc = lib.bk_tokenize_device(self.handle, d_text, len(text), 1 if final else 0, d_off, cap_records,
                                    C.byref(n), C.byref(consumed), C.byref(maxrec))
        off = self.d2h(d_off, (n.value + 1) * 4, np.uint32)
        self.dev_free(d_text)
        self.dev_free(d_off)
        return rc, off, consumed.value, maxrec.value

    def generate_device(self, cfg: SynConfig, mate: int, seed: int, first_idx: int, n: int,
                        d_text: int, d_off: int):
        self._check(lib.bk_generate_device(self.handle, C.byref(cfg), mate, seed, first_idx, n,
                                           d_text, d_off))

    def bus_probe(self, nbytes: int = 256 << 20, iters: int = 6):
        """-> (h2d, d2h, duplex) GB/s of this GPU's host link (pinned memory, see bk_bus_probe)."""
        a, b, c = C.c_double(0), C.c_double(0), C.c_double(0)
        self._check(lib.bk_bus_probe(self.handle, nbytes, iters, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def extract_device(self, n, m1: MateIn, m2: MateIn | None, d_out1, cap1, d_out2=0, cap2=0,
                       d_match1=0, d_match2=0, iters=1) -> Result:
        res = Result()
        rc = lib.bk_extract_device(self.handle, n, C.byref(m1), C.byref(m2) if m2 is not None else None,
                                   d_out1, cap1, d_out2 or None, cap2, d_match1 or None,
                                   d_match2 or None, iters, C.byref(res))
        self._check(rc)
        return res


def syn_config(read_len=150, plant1=None, plant2=None) -> SynConfig:
    """plantN = (literal bytes, offset, jitter) or None; mirrors oracle/bksyn.SynConfig."""
    cfg = SynConfig()
    cfg.read_len = read_len
    for i, p in enumerate((plant1, plant2)):
        if p is not None:
            lit, off, jit = p
            cfg.plant_len[i] = len(lit)
            cfg.plant_off[i] = off
            cfg.plant_jitter[i] = jit
            cfg.plant_lit[i].value = lit
    return cfg
