"""The LZ4 frame codec behind `--lz4` and lz4 inputs (barkit_b200/csrc/bk_lz4.h, written from the format's
specification) against two independent implementations: pyarrow's LZ4 frame codec and the system's liblz4
(through ctypes; linked blocks, block and content checksums, 64 KiB blocks)."""
import ctypes as C
import ctypes.util
import os
import random
import subprocess

import pytest

from oracle import bksyn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("lz4") / "lz4_driver")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "abi", "lz4_driver.cc")], check=True)
    return exe


def run(exe, mode, data):
    r = subprocess.run([exe, mode], input=data, capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    return r.stdout


def samples():
    rnd = random.Random(3)
    fq = bksyn.generate(bksyn.CONFIGS["C3"], 1, 0, 12000)              # ~4 MB of FASTQ: several blocks
    return [b"", b"A", b"ACGT" * 3, fq, bytes(rnd.randrange(256) for _ in range(70000)), b"\n" * 300000,
            fq[:(1 << 20)], fq[:(1 << 20) + 1]]


def test_own_frames_are_read_by_pyarrow_and_by_ourselves(driver):
    pa = pytest.importorskip("pyarrow")
    for d in samples():
        z = run(driver, "c", d)
        assert z[:4] == b"\x04\x22\x4d\x18"
        assert run(driver, "d", z) == d
        if d:
            assert pa.decompress(z, decompressed_size=len(d), codec="lz4", asbytes=True) == d
    fq = samples()[3]
    ours, theirs = len(run(driver, "c", fq)), len(pa.compress(fq, codec="lz4", asbytes=True))
    assert ours < 0.9 * len(fq) and ours < 1.1 * theirs                # random bases and qualities: LZ4 itself gets ~0.8


def test_pyarrow_frames_are_read(driver):
    pa = pytest.importorskip("pyarrow")
    for d in samples():
        if d:
            assert run(driver, "d", pa.compress(d, codec="lz4", asbytes=True)) == d
    two = pa.compress(b"first\n", codec="lz4", asbytes=True) + pa.compress(b"second\n", codec="lz4", asbytes=True)
    assert run(driver, "d", two) == b"first\nsecond\n"                 # concatenated frames


class _FrameInfo(C.Structure):
    _fields_ = [("blockSizeID", C.c_int), ("blockMode", C.c_int), ("contentChecksumFlag", C.c_int), ("frameType", C.c_int),
                ("contentSize", C.c_ulonglong), ("dictID", C.c_uint), ("blockChecksumFlag", C.c_int)]


class _Prefs(C.Structure):
    _fields_ = [("frameInfo", _FrameInfo), ("compressionLevel", C.c_int), ("autoFlush", C.c_uint), ("favorDecSpeed", C.c_uint),
                ("reserved", C.c_uint * 3)]


def test_liblz4_frames_linked_blocks_and_checksums(driver):
    name = ctypes.util.find_library("lz4") or "liblz4.so.1"
    try:
        lz = C.CDLL(name)
    except OSError:
        pytest.skip("liblz4 runtime not present")
    lz.LZ4F_compressFrameBound.restype = C.c_size_t
    lz.LZ4F_compressFrameBound.argtypes = [C.c_size_t, C.c_void_p]
    lz.LZ4F_compressFrame.restype = C.c_size_t
    lz.LZ4F_compressFrame.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
    lz.LZ4F_isError.argtypes = [C.c_size_t]
    fq = samples()[3]
    for block_id, linked, csum, bsum, level in ((4, 0, 1, 0, 0), (4, 0, 1, 1, 3), (5, 1, 0, 0, 0), (7, 1, 1, 1, 9), (0, 0, 1, 0, 0)):
        p = _Prefs()
        p.frameInfo.blockSizeID, p.frameInfo.blockMode = block_id, 1 - linked     # LZ4F_blockLinked = 0, independent = 1
        p.frameInfo.contentChecksumFlag, p.frameInfo.blockChecksumFlag, p.compressionLevel = csum, bsum, level
        cap = lz.LZ4F_compressFrameBound(len(fq), C.byref(p))
        dst = C.create_string_buffer(cap)
        n = lz.LZ4F_compressFrame(dst, cap, fq, len(fq), C.byref(p))
        assert not lz.LZ4F_isError(n)
        assert run(driver, "d", dst.raw[:n]) == fq
    # a damaged stream is an error, not silence
    r = subprocess.run([driver, "d"], input=dst.raw[:n // 2], capture_output=True)
    assert r.returncode == 1 and b"lz4" in r.stderr
