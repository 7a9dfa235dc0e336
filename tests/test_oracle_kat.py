"""The reference's own known-answer vectors, checked against the oracle.

Every case cites the reference test or doctest it is lifted from.
"""
import pytest

from oracle import barkit_oracle as bo


# pattern.rs:244-249 (rstest) + doctest pattern.rs:35-38
@pytest.mark.parametrize("expected,text,max_error", [
    (["."], "a", 1),
    (["A"], "a", 0),
    ([], "", 1),
    (["AAA.", "AA.A", "A.AA", ".AAA"], "AAAA", 1),
    (["..."], "AAA", 3),
    (["..."], "AAA", 4),
    (["ATG.", "AT.C", "A.GC", ".TGC"], "ATGC", 1),
])
def test_generate_sequences_with_pcr_errors(expected, text, max_error):
    assert bo.BarcodePattern("", max_error).get_sequence_with_errors(text) == expected


# pattern.rs:263-272 (rstest) + doctest pattern.rs:88-91
@pytest.mark.parametrize("expected,pattern,max_error", [
    ("^(AA.|A.A|.AA)(?P<UMI>[ATGCN]{3})", "^aaa(?P<UMI>[ATGCN]{3})", 1),
    ("^(...)(?P<UMI>[ATGCN]{3})", "^aaa(?P<UMI>[ATGCN]{3})", 3),
    ("^(...)(?P<UMI>[ATGCN]{3})", "^aaa(?P<UMI>[ATGCN]{3})", 4),
    ("^((...))(?P<UMI>[ATGCN]{3})", "^(aaa)(?P<UMI>[ATGCN]{3})", 4),
    ("^(AA.|A.A|.AA)(?P<UMI>[ATGCN]{3})CCC", "^aaa(?P<UMI>[ATGCN]{3})CCC", 1),
    ("^(?P<UMI>[ATGCN]{3})", "^(?P<UMI>[ATGCN]{3})", 1),
    ("^(ATG.|AT.C|A.GC|.TGC)(?<UMI>[ATGCN]{12})", "^atgc(?<UMI>[ATGCN]{12})", 1),
])
def test_create_fuzzy(expected, pattern, max_error):
    assert bo.BarcodePattern(pattern, max_error).get_pattern_with_errors() == expected


# doctest pattern.rs:130-136, :142
def test_parse_type():
    assert bo.BarcodeType.parse_type("UMI") == bo.BarcodeType.UMI
    assert bo.BarcodeType.parse_type("SB") == bo.BarcodeType.SAMPLE
    assert bo.BarcodeType.parse_type("CB") == bo.BarcodeType.CELL
    with pytest.raises(bo.UnexpectedCaptureGroupName) as e:
        bo.BarcodeType.parse_type("XX")
    assert str(e.value) == "Provided unexpected barcode capture group XX"


# doctest pattern.rs:211-224
def test_get_captures_doctest():
    rx = bo.BarcodeRegex("^atgc(?<UMI>[ATGCN]{6})", 1)
    assert rx.get_captures(b"ATGCNNNNNNCCC").group("UMI") == b"NNNNNN"
    with pytest.raises(bo.PatternNotMatched):
        rx.get_captures(b"CCCCCCCCCCCCC")


# parse.rs:188-193 (rstest)
@pytest.mark.parametrize("seq,rc", [
    (b"", b""),
    (b"GGGCCCAAATTT", b"AAATTTGGGCCC"),
    (b"ATGCN", b"NGCAT"),
    (b"AAP", b"ATT"),
    (b"CCX", b"AGG"),
    (b"PPP", b"AAA"),
])
def test_get_reverse_complement(seq, rc):
    assert bo.get_reverse_complement(seq) == rc


# pattern.rs:201-203
def test_no_capture_group_is_an_error():
    with pytest.raises(bo.BarcodeCaptureGroupNotFound):
        bo.BarcodeRegex("^ATGC", 1)


# ---- derived worked examples (SURVEY.md App. B "derived"; NOT reference vectors) -------------

def _qual(n, start=33):
    return bytes(range(start, start + n))


def test_derived_A1_A2_A3():
    pat = "^(?P<CB>[ATGCN]{16})atgccat"
    assert bo.BarcodePattern(pat, 1).get_pattern_with_errors() == (
        "^(?P<CB>[ATGCN]{16})(ATGCCA.|ATGCC.T|ATGC.AT|ATG.CAT|AT.CCAT|A.GCCAT|.TGCCAT)")
    seq = b"ACGTACGTACGTACGT" + b"ATGACAT" + b"GGGGGTTTTT"
    fq = b"@r1 2:N:0:1\n" + seq + b"\n+\n" + _qual(len(seq)) + b"\n"
    out = bo.extract_se(fq, pat, 1)
    assert out == (b"@r1 2:N:0:1 CB:ACGTACGTACGTACGT:" + _qual(16) + b"\nGGGGGTTTTT\n+\n"
                   + _qual(10, 33 + 23) + b"\n")
    seq2 = b"ACGTACGTACGTACGT" + b"ATGAAAT" + b"GGGGGTTTTT"          # 2 substitutions
    fq2 = b"@r1\n" + seq2 + b"\n+\n" + _qual(len(seq2)) + b"\n"
    assert bo.extract_se(fq2, pat, 1) == b""
    assert bo.extract_se(fq2, pat, 2) != b""
    assert len(bo.BarcodePattern(pat, 2).get_sequence_with_errors("atgccat")) == 21
    seq3 = b"ACGTACGTACGTACGT" + b"ATGCAT" + b"GGGGGTTTTTT"          # 1 deletion: Hamming only
    fq3 = b"@r1\n" + seq3 + b"\n+\n" + _qual(len(seq3)) + b"\n"
    assert bo.extract_se(fq3, pat, 1) == b""


def test_derived_B_unanchored_order_and_hole():
    pat = "(?P<SB>[ATGC]{4})gattaca(?P<UMI>[ATGCN]{6})"
    seq = b"NNACGTGATTACAAAAAAACCCCGATTACATTTTTTGG"
    qual = _qual(len(seq))
    out = bo.extract_se(b"@id desc\n" + seq + b"\n+\n" + qual + b"\n", pat, 1)
    assert out == (b"@id desc SB:ACGT:" + qual[2:6] + b" UMI:AAAAAA:" + qual[13:19] + b"\n"
                   + seq[:2] + seq[19:] + b"\n+\n" + qual[:2] + qual[19:] + b"\n")


def test_derived_C_rc_quirk():
    out = bo.extract_se(b"@q\nTTAAGGGACGT\n+\nABCDEFGHIJK\n", "^(?P<UMI>[ATGCN]{4})ccc", 0,
                        rc_barcodes=True)
    assert out == b"@q UMI:TTAA:ABCD\nACGT\n+\nHIJK\n"


@pytest.mark.parametrize("pattern,expected", [
    ("ATGCatgc", "ATGCatgc"),
    ("[atgc]{3}x", "[atgc]{3}x"),
    (r"\d{3}atg", r"\(.){3}(AT.|A.G|.TG)"),
    ("(?s)atg", "(?(.))(AT.|A.G|.TG)"),
    ("(?P<CB>atgc)", "(?P<CB>(ATG.|AT.C|A.GC|.TGC))"),
    ("^n{3}(?P<UMI>[ATGCN]{4})", "^(.){3}(?P<UMI>[ATGCN]{4})"),
])
def test_derived_finder_edge_cases(pattern, expected):
    assert bo.BarcodePattern(pattern, 1).get_pattern_with_errors() == expected


def test_pe_policy():
    fq1 = b"@a 1\nACGTTTTT\n+\nIIIIHHHH\n@b 1\nNNNNNNNN\n+x\n12345678\n@c 1\nNAAAAAAA\n+\nabcdefgh\n"
    fq2 = b"@a 2\nGGGGCCCC\n+\nJJJJKKKK\n@b 2\nGGGGAAAA\n+\n87654321\n@c 2\nTTTTTTTT\n+\nhgfedcba\n"
    o1, o2 = bo.extract_pe(fq1, fq2, "^(?P<UMI>[ACGT]{4})", "^(?P<CB>GGGG)", 1)
    # pair a: both match; pair b: only mate 2 (mate 1 kept unchanged, '+x' normalised); pair c: dropped
    assert o1 == b"@a 1 UMI:ACGT:IIII\nTTTT\n+\nHHHH\n@b 1\nNNNNNNNN\n+\n12345678\n"
    assert o2 == b"@a 2 CB:GGGG:JJJJ\nCCCC\n+\nKKKK\n@b 2 CB:GGGG:8765\nAAAA\n+\n4321\n"


def test_readme_header_format():
    """README.md:37 of the reference documents the rewritten header: `@SEQ_ID UMI:ATGC:???? CB:ATGC:???? SB:ATGC:????`
    -- the only statement of row H's output format (parse.rs:161-168) that the reference holds outside its code:
    one space in front of every barcode, NAME ':' sequence ':' quality, in pattern order."""
    pat = "^(?P<UMI>[ATGC]{4})(?P<CB>[ATGC]{4})(?P<SB>[ATGC]{4})"
    out = bo.extract_se(b"@SEQ_ID\nATGCATGCATGCGG\n+\n????????????II\n", pat, 0)
    assert out.split(b"\n")[0] == b"@SEQ_ID UMI:ATGC:???? CB:ATGC:???? SB:ATGC:????"
    assert out == b"@SEQ_ID UMI:ATGC:???? CB:ATGC:???? SB:ATGC:????\nGG\n+\nII\n"
