"""CPU oracle for the `barkit extract` hot path  --  TEST INFRASTRUCTURE ONLY.

This module is a plain-Python restatement of the reference algorithm
(nsyzrantsev/barkit v0.1.1, pure Rust).  It exists to check the CUDA path; it is never on the
product path.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu-baseline /
`--impl reference` legs may import anything under `oracle/`.

Why a restatement: the reference is Rust and this image has no `cargo`/`rustc` (and no network
or vendored crates), so it cannot be compiled or imported here.  The regex engine stand-in is
Python's `re` on bytes: like the Rust `regex` crate it reports the leftmost match and, among
matches starting at the same place, prefers earlier alternatives (leftmost-first), which is all
the reference relies on (`pattern.rs:225-230`).

Parity pinning (what anchors this oracle to the reference):
  * PINNED by the reference's own known-answer tests (`tests/test_oracle_kat.py`):
      - `get_sequence_with_errors`   pattern.rs:244-249 (6 rstest cases) + doctest pattern.rs:35-38
      - `get_pattern_with_errors`    pattern.rs:263-272 (6 rstest cases) + doctest pattern.rs:88-91
      - `BarcodeType::parse_type`    doctest pattern.rs:133-135
      - `BarcodeRegex::get_captures` doctest pattern.rs:213-223
      - `get_reverse_complement`     parse.rs:188-193 (6 rstest cases)
  * PARITY UNPINNED (no reference test, fixture or runnable binary exists for these; they follow
    the code line by line and are pinned only by that reading): header rewrite
    (parse.rs:92-116,150-171), trimming (parse.rs:138-148), SE drop / PE merge policy
    (run.rs:63-80,149-195), FASTQ tokenising and framing (seq_io 0.3.2, fastq.rs:259-263).

Third-party behaviour restated (crates absent from /root/reference, pinned in Cargo.lock):
  regex 1.11.1, fancy-regex 0.14.0 (the finder), seq_io 0.3.2 (record fields, write_to framing),
  rayon 1.10.0 (order-preserving collect).
"""
from __future__ import annotations

import re
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence, Tuple

# --------------------------------------------------------------------------------------------
# error.rs:3-22  (message strings are part of the boundary)
# --------------------------------------------------------------------------------------------


class BarkitError(Exception):
    """error.rs:4 `enum Error`; `str(e)` reproduces the `#[error(...)]` text."""


class RegexError(BarkitError):  # error.rs:9-10
    def __str__(self):
        return "Regex error: %s" % (self.args[0],)


class BarcodeCaptureGroupNotFound(BarkitError):  # error.rs:11-12
    def __str__(self):
        return "%s capture group not found in your pattern" % (self.args[0],)


class UnexpectedCaptureGroupName(BarkitError):  # error.rs:13-14
    def __str__(self):
        return "Provided unexpected barcode capture group %s" % (self.args[0],)


class PatternNotMatched(BarkitError):  # error.rs:17-18
    def __str__(self):
        return "No match"


# --------------------------------------------------------------------------------------------
# pattern.rs
# --------------------------------------------------------------------------------------------

FUZZY_CHARACTER = "."  # pattern.rs:10
ADAPTER_PATTERN_REGEX = r"(?<!\[)\b[atgcryswkmbdhvn]+\b(?!\])"  # pattern.rs:11 (valid `re` verbatim)


class BarcodePattern:
    """pattern.rs:13-110."""

    def __init__(self, pattern: str, max_error: int):  # pattern.rs:20-26
        self.adapter_pattern = re.compile(ADAPTER_PATTERN_REGEX)
        self.barcode_pattern = pattern
        self.max_error = max_error

    def get_sequence_with_errors(self, sequence: str) -> List[str]:
        """pattern.rs:40-79: all ways to keep exactly n-k characters (others become '.'),
        enumerated by ascending mask value, bit i <-> character i kept."""
        if self.max_error == 0:  # pattern.rs:41-43
            return [sequence.upper()]
        if len(sequence) == 0:  # pattern.rs:45-47
            return []
        if self.max_error >= len(sequence):  # pattern.rs:49-51
            return [FUZZY_CHARACTER * len(sequence)]
        n = len(sequence)
        upper = sequence.upper()
        keep = n - self.max_error
        cases = []
        # pattern.rs:64-77 walks every mask 0..=2^n-1 and filters on popcount; Gosper's hack
        # visits exactly the surviving masks in the same ascending order.
        mask = (1 << keep) - 1
        limit = 1 << n
        while mask < limit:
            cases.append("".join(upper[i] if (mask >> i) & 1 else FUZZY_CHARACTER for i in range(n)))
            c = mask & -mask
            r = mask + c
            mask = (((r ^ mask) >> 2) // c) | r
        return cases

    def get_pattern_with_errors(self) -> str:
        """pattern.rs:93-109."""
        result = []
        last_end = 0
        for mat in self.adapter_pattern.finditer(self.barcode_pattern):  # pattern.rs:97
            result.append(self.barcode_pattern[last_end:mat.start()])
            fuzzy = self.get_sequence_with_errors(mat.group(0))  # pattern.rs:101
            result.append("(%s)" % "|".join(fuzzy))  # pattern.rs:102
            last_end = mat.end()
        result.append(self.barcode_pattern[last_end:])
        return "".join(result)


class BarcodeType:
    """pattern.rs:112-159."""

    UMI = "UMI"
    SAMPLE = "SB"
    CELL = "CB"

    @staticmethod
    def parse_type(name: str) -> str:  # pattern.rs:137-144
        if name in ("UMI", "SB", "CB"):
            return name
        raise UnexpectedCaptureGroupName(name)


# `.` of regex::bytes::Regex in its default Unicode mode (pattern.rs:6,182): one UTF-8 encoded scalar value other
# than '\n' -- one to four bytes; a byte >= 0x80 that is not part of a well-formed sequence matches nothing.
# (`re` on bytes would take any single byte.)  The well-formed ranges are those of the Unicode standard, table 3-7.
_UTF8_DOT = (r"(?:[\x00-\x09\x0b-\x7f]|[\xc2-\xdf][\x80-\xbf]|\xe0[\xa0-\xbf][\x80-\xbf]|[\xe1-\xec\xee\xef][\x80-\xbf]{2}"
             r"|\xed[\x80-\x9f][\x80-\xbf]|\xf0[\x90-\xbf][\x80-\xbf]{2}|[\xf1-\xf3][\x80-\xbf]{3}|\xf4[\x80-\x8f][\x80-\xbf]{2})")


def _rust_regex_to_py(pattern: str) -> str:
    """The Rust `regex` crate accepts `(?<name>` and `(?P<name>`; `re` only the latter.  `.` outside a class becomes
    the UTF-8 scalar alternation above."""
    pattern = re.sub(r"\(\?<(?![=!])", "(?P<", pattern)
    out, in_class, i = [], False, 0
    while i < len(pattern):
        c = pattern[i]
        if c == "\\" and i + 1 < len(pattern):
            out.append(pattern[i:i + 2])
            i += 2
            continue
        if in_class:
            in_class = c != "]"
        elif c == "[":
            in_class = True
            if pattern[i + 1:i + 2] == "^":          # `[^...]`, `[]...]`: the first member is literal
                out.append("[^")
                i += 2
                c = ""
            if pattern[i + (1 if c else 0):i + (2 if c else 1)] == "]":
                out.append(c + "]")
                i += 2 if c else 1
                continue
            if not c:
                continue
        elif c == ".":
            out.append(_UTF8_DOT)
            i += 1
            continue
        out.append(c)
        i += 1
    return "".join(out)


class BarcodeRegex:
    """pattern.rs:161-235."""

    def __init__(self, pattern: str, max_error: int):  # pattern.rs:179-188
        barcode_pattern = BarcodePattern(pattern, max_error)
        self.fuzzy_pattern = barcode_pattern.get_pattern_with_errors()
        try:
            self.regex = re.compile(_rust_regex_to_py(self.fuzzy_pattern).encode("latin-1"))
        except re.error as e:  # pattern.rs:182 `?` -> Error::Regex
            raise RegexError(str(e))
        self.barcode_types = self._parse_capture_groups()

    def _parse_capture_groups(self) -> List[str]:  # pattern.rs:191-205
        names = sorted(self.regex.groupindex.items(), key=lambda kv: kv[1])  # group-index order
        groups = [BarcodeType.parse_type(name) for name, _ in names]
        if not groups:
            raise BarcodeCaptureGroupNotFound(self.fuzzy_pattern)
        return groups

    def get_captures(self, read_seq: bytes):  # pattern.rs:225-230
        m = self.regex.search(read_seq)
        if m is None:
            raise PatternNotMatched()
        return m

    def get_barcode_types(self) -> List[str]:  # pattern.rs:232-234
        return list(self.barcode_types)


# --------------------------------------------------------------------------------------------
# parse.rs
# --------------------------------------------------------------------------------------------

# parse.rs:12-32: default b'A'; A<->T, G<->C, IUPAC ambiguity codes map to themselves.
TRANSLATION_TABLE = bytearray(b"A" * 256)
for _a, _b in ((b"A", b"T"), (b"T", b"A"), (b"G", b"C"), (b"C", b"G")):
    TRANSLATION_TABLE[_a[0]] = _b[0]
for _c in b"RYSWKMBDHVN":
    TRANSLATION_TABLE[_c] = _c
TRANSLATION_TABLE = bytes(TRANSLATION_TABLE)


def get_reverse_complement(sequence: bytes) -> bytes:  # parse.rs:173-179
    return bytes(sequence[::-1]).translate(TRANSLATION_TABLE)


@dataclass
class Record:
    """seq_io `RefRecord`/`OwnedRecord`: head = whole line after '@', seq, qual (no EOL/CR)."""
    head: bytes
    seq: bytes
    qual: bytes


class BarcodeParser:
    """parse.rs:34-117."""

    def __init__(self, barcode_regex: BarcodeRegex, skip_trimming: bool, rc_barcodes: bool):
        self.barcode_regex = barcode_regex
        self.skip_trimming = skip_trimming
        self.rc_barcodes = rc_barcodes

    @classmethod
    def new(cls, barcode_regex: Optional[BarcodeRegex], skip_trimming: bool,
            rc_barcodes: bool) -> Optional["BarcodeParser"]:  # parse.rs:46-56
        if barcode_regex is None:
            return None
        return cls(barcode_regex, skip_trimming, rc_barcodes)

    def parse_barcodes(self, record: Record) -> Optional[Record]:  # parse.rs:58-68
        captures = self.barcode_regex.regex.search(record.seq)
        if captures is None and self.rc_barcodes:
            # parse.rs:61-63: offsets found on the reverse complement are later applied to the
            # FORWARD seq/qual (parse.rs:97-99,138-141) -- reproduced on purpose.
            captures = self.barcode_regex.regex.search(get_reverse_complement(record.seq))
        return self._create_read(captures, record)

    def _create_read(self, captures, record: Record) -> Optional[Record]:  # parse.rs:70-90
        if captures is None:
            return None  # (Err(_), _) => None
        try:
            new_read = self._create_read_with_new_header(captures, record)
            if self.skip_trimming:
                return new_read
            return trim_adapters(captures, new_read)
        except BarkitError:
            return None  # `.ok()?`

    def _create_read_with_new_header(self, captures, record: Record) -> Record:  # parse.rs:92-116
        head, seq, qual = record.head, record.seq, record.qual
        for barcode_name in self.barcode_regex.get_barcode_types():  # pattern order
            start, end = get_barcode_match_positions(barcode_name, captures)
            head = add_to_the_header(barcode_name, head, seq, qual, start, end)
        return Record(head, seq, qual)


def get_full_match_positions(captures) -> Tuple[int, int]:  # parse.rs:119-125
    return captures.start(0), captures.end(0)


def get_barcode_match_positions(name: str, captures) -> Tuple[int, int]:  # parse.rs:127-136
    if captures.start(name) < 0:
        raise BarcodeCaptureGroupNotFound(name)
    return captures.start(name), captures.end(name)


def trim_adapters(captures, record: Record) -> Record:  # parse.rs:138-148
    start, end = get_full_match_positions(captures)
    return Record(record.head, record.seq[:start] + record.seq[end:],
                  record.qual[:start] + record.qual[end:])


def add_to_the_header(barcode_type: str, head: bytes, seq: bytes, qual: bytes,
                      start: int, end: int) -> bytes:  # parse.rs:150-171
    barcode_seq = seq[start:end]
    barcode_qual = qual[start:end]
    try:
        barcode_seq.decode("utf-8")  # parse.rs:166 `from_utf8(barcode_seq)?`
    except UnicodeDecodeError:
        raise BarkitError("UTF-8 error")
    return head + b" " + barcode_type.encode() + b":" + barcode_seq + b":" + barcode_qual


# --------------------------------------------------------------------------------------------
# run.rs batch policies
# --------------------------------------------------------------------------------------------


def parse_se_reads(records: Sequence[Record], barcode: BarcodeRegex, skip_trimming: bool,
                   rc_barcodes: bool) -> List[Record]:  # run.rs:63-80 (ordered filter_map)
    out = []
    for record in records:
        parser = BarcodeParser.new(barcode, skip_trimming, rc_barcodes)
        new = parser.parse_barcodes(record) if parser is not None else None
        if new is not None:
            out.append(new)
    return out


def get_new_reads(new_records, record1: Record, record2: Record):  # run.rs:149-160
    n1, n2 = new_records
    if n1 is not None and n2 is not None:
        return n1, n2
    if n1 is None and n2 is not None:
        return Record(record1.head, record1.seq, record1.qual), n2
    if n1 is not None and n2 is None:
        return n1, Record(record2.head, record2.seq, record2.qual)
    return None


def parse_pe_reads(records1: Sequence[Record], records2: Sequence[Record],
                   barcode1: Optional[BarcodeRegex], barcode2: Optional[BarcodeRegex],
                   skip_trimming: bool, rc_barcodes: bool):  # run.rs:163-195 (zip truncates)
    out = []
    for record1, record2 in zip(records1, records2):
        p1 = BarcodeParser.new(barcode1, skip_trimming, rc_barcodes)
        p2 = BarcodeParser.new(barcode2, skip_trimming, rc_barcodes)
        new_reads = (p1.parse_barcodes(record1) if p1 is not None else None,
                     p2.parse_barcodes(record2) if p2 is not None else None)
        pair = get_new_reads(new_reads, record1, record2)
        if pair is not None:
            out.append(pair)
    return out


# --------------------------------------------------------------------------------------------
# seq_io 0.3.2 stand-ins (crate not in tree; Cargo.lock:810-811)
# --------------------------------------------------------------------------------------------


def _trim_cr(line: bytes) -> bytes:
    return line[:-1] if line.endswith(b"\r") else line


def read_fastq(text: bytes) -> List[Record]:
    """Four-line FASTQ tokeniser with seq_io's field definitions: head = the '@' line without
    '@' and EOL (CR stripped); seq / qual single lines; the '+' line's content is ignored; the
    final record may lack a trailing newline."""
    records = []
    pos, n = 0, len(text)
    while pos < n:
        lines = []
        for _ in range(4):
            if pos > n:
                raise ValueError("truncated FASTQ record")
            nl = text.find(b"\n", pos)
            if nl < 0:
                lines.append(text[pos:])
                pos = n + 1
            else:
                lines.append(text[pos:nl])
                pos = nl + 1
        if not lines[0].startswith(b"@"):
            raise ValueError("record does not start with '@'")
        if not lines[2].startswith(b"+"):
            raise ValueError("separator line does not start with '+'")
        records.append(Record(_trim_cr(lines[0][1:]), _trim_cr(lines[1]), _trim_cr(lines[3])))
    return records


def write_record(rec: Record) -> bytes:
    """seq_io::fastq::write_to (fastq.rs:261): `@head\\nseq\\n+\\nqual\\n`."""
    return b"@" + rec.head + b"\n" + rec.seq + b"\n+\n" + rec.qual + b"\n"


def write_fastq(records: Iterable[Record]) -> bytes:  # fastq.rs:265-271
    return b"".join(write_record(r) for r in records)


# --------------------------------------------------------------------------------------------
# whole-path helpers (text in -> text out), i.e. run.rs:122-143 / :249-274 minus file I/O
# --------------------------------------------------------------------------------------------


def extract_se(fq: bytes, pattern: str, max_error: int = 1, rc_barcodes: bool = False,
               skip_trimming: bool = False) -> bytes:
    barcode = BarcodeRegex(pattern, max_error)
    return write_fastq(parse_se_reads(read_fastq(fq), barcode, skip_trimming, rc_barcodes))


def extract_pe(fq1: bytes, fq2: bytes, pattern1: Optional[str], pattern2: Optional[str],
               max_error: int = 1, rc_barcodes: bool = False,
               skip_trimming: bool = False) -> Tuple[bytes, bytes]:
    b1 = BarcodeRegex(pattern1, max_error) if pattern1 is not None else None
    b2 = BarcodeRegex(pattern2, max_error) if pattern2 is not None else None
    pairs = parse_pe_reads(read_fastq(fq1), read_fastq(fq2), b1, b2, skip_trimming, rc_barcodes)
    return (write_fastq(p[0] for p in pairs), write_fastq(p[1] for p in pairs))
