"""The drop-in boundary driven from compiled C (gcc, -std=c11): include/barkit_b200.h must be a valid C header,
its struct sizes must be what the bindings assume (ctypes here, rust/barkit-extract/src/ffi.rs),
and bk_compile -> bk_ctx_create -> bk_submit / bk_wait must work without Python in between."""
import os
import re
import subprocess

import pytest

from barkit_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "barkit_b200")


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("abi") / "abi_driver")
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "abi", "abi_driver.c"), "-o", exe, "-L", LIBDIR, "-lbarkit_b200",
                    "-Wl,-rpath," + LIBDIR], check=True)
    return exe


def test_struct_sizes_agree_with_the_bindings(driver):
    import ctypes as C
    out = subprocess.run([driver, "sizes"], check=True, capture_output=True, text=True).stdout
    sizes = dict(line.split() for line in out.splitlines())
    assert int(sizes["bk_mate_in"]) == C.sizeof(_lib.MateIn) == 32
    assert int(sizes["bk_result"]) == C.sizeof(_lib.Result) == 72
    assert int(sizes["bk_run_args"]) == C.sizeof(_lib.RunArgs) == 104
    assert int(sizes["bk_program_info"]) == C.sizeof(_lib.ProgramInfo) == 48
    assert int(sizes["bk_syn_config"]) == C.sizeof(_lib.SynConfig) == 92


def _rust_struct_fields(text, name):
    body = re.search(r"pub struct %s \{(.*?)\n\}" % name, text, re.S).group(1)
    return re.findall(r"pub (\w+):", body)


def test_rust_ffi_and_integration_doc_match_the_header():
    """rust/barkit-extract/src/ffi.rs cannot be compiled here; keep it honest mechanically: every function it
    declares is exported by the library, and its repr(C) structs list the same fields, in the same order, as the
    ctypes structures that the GPU tests run against."""
    ffi = open(os.path.join(ROOT, "rust", "barkit-extract", "src", "ffi.rs")).read()
    for fn in re.findall(r"pub fn (bk_\w+)\(", ffi):
        assert hasattr(_lib.lib, fn), fn
    assert _rust_struct_fields(ffi, "BkResult") == [f[0] for f in _lib.Result._fields_]
    assert _rust_struct_fields(ffi, "BkMateIn") == [f[0] for f in _lib.MateIn._fields_]
    assert _rust_struct_fields(ffi, "BkRunArgs") == [f[0] for f in _lib.RunArgs._fields_]
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    assert "h2d_bytes" in doc and "d2h_bytes" in doc          # the round-1 doc stopped at `launches`
    hdr = open(os.path.join(ROOT, "include", "barkit_b200.h")).read()
    declared = set(re.findall(r"\b(bk_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)


def test_front_end_from_c_and_loud_failure_without_a_gpu(driver):
    out = subprocess.run([driver, "nogpu"], check=True, capture_output=True, text=True).stdout.splitlines()
    assert out[0] == "expand ^(ATG.|AT.C|A.GC|.TGC)(?<UMI>[ATGCN]{12})"
    assert out[1] == "caps 2 CB UMI"
    assert out[2] == "error Provided unexpected barcode capture group XX"
    if _lib.device_count() == 0:
        assert out[3].startswith("ctx -3 ") and "no CPU path" in out[3]
    else:
        assert out[3] == "ctx ok"


@pytest.mark.gpu
def test_submit_wait_from_compiled_c_matches_the_oracle(driver, tmp_path):
    from oracle import barkit_oracle as bo
    from oracle import bksyn
    fq1, _ = bksyn.generate_pair(bksyn.CONFIGS["C1"], 0, 5000)
    fq1 += b"@last no eol\nACGTACGTACGTACGT\n+\nIIIIIIIIIIIIIIII"
    inp, outp = tmp_path / "in.fq", tmp_path / "out.fq"
    inp.write_bytes(fq1)
    pat = "^(?P<UMI>[ATGCN]{12})"
    r = subprocess.run([driver, "extract", str(inp), str(outp), pat, "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert outp.read_bytes() == bo.extract_se(fq1, pat, 1)
    assert "launches 1" in r.stdout
