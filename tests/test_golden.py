"""Committed golden fixtures (tests/golden/, made by tests/golden/make_golden.py from the Python restatement of
the reference): the C restatement must reproduce them on the CPU, the CUDA path on the GPU, byte for byte."""
import json
import os

import pytest

from oracle import barkit_oracle as bo
from oracle import cport

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = json.load(open(os.path.join(GOLDEN, "manifest.json")))


def load(case):
    def rd(suffix):
        p = os.path.join(GOLDEN, case["tag"] + suffix)
        return open(p, "rb").read() if os.path.exists(p) else None
    return rd("_in1.fq"), rd("_in2.fq"), rd("_out1.fq"), rd("_out2.fq")


@pytest.mark.parametrize("case", CASES, ids=[c["tag"] for c in CASES])
def test_python_oracle_reproduces_golden(case):
    i1, i2, o1, o2 = load(case)
    if case["paired"]:
        assert bo.extract_pe(i1, i2, case["pattern1"], case["pattern2"], case["max_error"],
                             rc_barcodes=case["rc_barcodes"], skip_trimming=case["skip_trimming"]) == (o1, o2)
    else:
        assert bo.extract_se(i1, case["pattern1"], case["max_error"], case["rc_barcodes"], case["skip_trimming"]) == o1


@pytest.mark.parametrize("case", CASES, ids=[c["tag"] for c in CASES])
def test_c_oracle_reproduces_golden(case):
    i1, i2, o1, o2 = load(case)
    r = cport.extract(i1, i2 if case["paired"] else None, case["pattern1"], case["pattern2"], case["max_error"],
                      rc_barcodes=case["rc_barcodes"], skip_trimming=case["skip_trimming"])
    assert bytes(r["out1"]) == o1
    if case["paired"]:
        assert bytes(r["out2"]) == o2


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c["tag"] for c in CASES])
def test_cuda_reproduces_golden(case):
    from barkit_b200 import _lib
    i1, i2, o1, o2 = load(case)
    k = case["max_error"]
    P1 = _lib.Program(case["pattern1"], k) if case["pattern1"] else None
    P2 = _lib.Program(case["pattern2"], k) if case["pattern2"] else None
    ctx = _lib.Context(P1, P2, paired=case["paired"], skip_trimming=case["skip_trimming"], rc_barcodes=case["rc_barcodes"])
    if case["paired"]:
        g1, g2, _ = ctx.extract_text(i1, i2)
        assert g1 == o1 and g2 == o2
    else:
        g1, _ = ctx.extract_text(i1)
        assert g1 == o1
    ctx.close()
