"""Oracle restatement of the two CLI drivers -- TEST INFRASTRUCTURE ONLY.

* ``parse_fasta``      : seq_io 0.3.2 ``fasta::Reader`` record rules (used via src/utils.rs:9-27).
* ``cli_canonicalize`` : src/canonicalize.rs:7-51  (bytes of a FASTA file in -> bytes of stdout/-o out).
* ``cli_uniq``         : src/uniq.rs:15-88         (same, plus the ``--table`` bytes).

seq_io is not vendored in the reference; the record rules below restate its published
reader (record starts at a '>' that follows a '\\n'; ``head()`` = header line without '>'
and without one trailing '\\r'; ``seq()`` = everything between the header's '\\n' and the
record's last '\\n' (or EOF), internal line breaks INCLUDED, one trailing '\\r' trimmed;
``id()`` = head up to the first ' ').  The reference's fixtures pin: wrapped lines, missing
final newline, trailing blank line, header with a space.  CR handling and the raw echo
of ``uniq`` without ``-c`` are unpinned by the reference's tests.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import canonicalize_batch, uniq_consume


@dataclass
class Record:
    head: bytes      # header line without '>' / trailing CR
    seq: bytes       # raw sequence bytes, internal line breaks included

    @property
    def id(self) -> bytes:
        return self.head.split(b" ", 1)[0]


class FastaError(ValueError):
    pass


def _trim_cr(b: bytes) -> bytes:
    return b[:-1] if b.endswith(b"\r") else b


def parse_fasta(data: bytes) -> list[Record]:
    # leading empty lines ("" or "\r") are skipped; the first other line must start with '>'
    pos, n = 0, len(data)
    while pos < n:
        nl = data.find(b"\n", pos)
        line = data[pos: nl if nl >= 0 else n]
        if line not in (b"", b"\r"):
            break
        if nl < 0:
            return []
        pos = nl + 1
    if pos >= n:
        return []
    if data[pos: pos + 1] != b">":
        raise FastaError("expected '>' at record start")
    recs: list[Record] = []
    while pos < n:
        # record end: a '\n' directly followed by '>', else EOF
        search = pos + 1
        nxt = data.find(b"\n>", search)
        if nxt >= 0:
            region_end, new_pos = nxt, nxt + 1      # terminating '\n' excluded
        else:
            region_end = n - 1 if data.endswith(b"\n") else n
            new_pos = n
        first_nl = data.find(b"\n", pos, region_end)
        if first_nl < 0:
            # header line only
            head = _trim_cr(data[pos + 1: region_end])
            seq = b""
        else:
            head = _trim_cr(data[pos + 1: first_nl])
            seq = _trim_cr(data[first_nl + 1: region_end])
        recs.append(Record(head, seq))
        pos = new_pos
    return recs


def _batch(recs: list[Record]):
    lens = np.fromiter((len(r.seq) for r in recs), dtype=np.uint64, count=len(recs))
    offsets = np.zeros(len(recs) + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    arena = np.frombuffer(b"".join(r.seq for r in recs), dtype=np.uint8)
    return arena, offsets


def cli_canonicalize(fasta: bytes, threads: int = 1) -> bytes:
    """src/canonicalize.rs:7-51: normalise (:24), canonicalise (:29), write '>' head '\\n' seq '\\n' (:33-37)."""
    recs = parse_fasta(fasta)
    if not recs:
        return b""
    arena, offsets = _batch(recs)
    res = canonicalize_batch(arena, offsets, normalize=True, threads=threads, want_start=False, want_hash=False)
    out = bytearray()
    ob = res["out"].tobytes()
    for i, r in enumerate(recs):
        o = int(offsets[i])
        out += b">" + r.head + b"\n" + ob[o: o + int(res["lens"][i])] + b"\n"
    return bytes(out)


def _csv_field(f: bytes, delim: bytes) -> bytes:
    # csv 1.2.2 QuoteStyle::Necessary, double_quote = true, terminator '\n'
    if any(c in f for c in (delim, b'"', b"\n", b"\r")):
        return b'"' + f.replace(b'"', b'""') + b'"'
    return f


def cli_uniq(fasta: bytes, canonicalize: bool = False, table_ext: str | None = None, threads: int = 1):
    """src/uniq.rs:15-88.  Returns (fasta_out, table_bytes or None)."""
    recs = parse_fasta(fasta)
    delim = b"\t" if table_ext == "tsv" else b","        # src/utils.rs:74-84
    table = bytearray() if table_ext is not None else None
    if not recs:
        return b"", (bytes(table) if table is not None else None)
    arena, offsets = _batch(recs)
    res = canonicalize_batch(arena, offsets, normalize=True, threads=threads, want_start=False, want_hash=False)
    _, first = uniq_consume(res["out"], offsets, res["lens"])
    ob = res["out"].tobytes()
    out = bytearray()
    wrote_header = False
    for i, r in enumerate(recs):
        r.id.decode("utf-8")                               # record.id().unwrap() (:48,:67)
        if int(first[i]) == i:                             # :47-61
            o = int(offsets[i])
            body = ob[o: o + int(res["lens"][i])] if canonicalize else r.seq
            out += b">" + r.head + b"\n" + body + b"\n"
        elif table is not None:                            # :62-71
            if not wrote_header:
                table += b"id" + delim + b"duplicate_id\n"
                wrote_header = True
            table += _csv_field(recs[int(first[i])].id, delim) + delim + _csv_field(r.id, delim) + b"\n"
    return bytes(out), (bytes(table) if table is not None else None)


def full_seq(seq: bytes) -> bytes:
    """seq_io 0.3.2 fasta Record::full_seq: the sequence lines joined, each without its '\\n' and one trailing '\\r'"""
    return b"".join(_trim_cr(line) for line in seq.split(b"\n"))


def cli_monomerize(fasta: bytes, sensitive: bool = False, seed_length: int = 10, max_mismatch=None, min_identity=None,
                   min_overlap=None, min_overlap_percent=None, min_length: int = 0, max_length=None, keep_all: bool = False,
                   table_ext: str | None = None):
    """src/monomerize.rs:16-160 -> (bytes of -o, bytes of --table or None).  Worker (:83-101): normalise (None -> the raw
    seq), too short for the seed or for --min-length -> None, else last_monomer_end_index[_sensitive] of the NORMALISED
    sequence.  Consumer (:102-150): filters on the monomer length / overlap, then '>' head '\\n' full_seq[..end] '\\n' and
    the table row (id = the whole head, original_length, monomer_length)."""
    from . import normalize
    from .monomerize import Monomerizer
    if max_mismatch is not None and min_identity is not None:
        raise ValueError("cannot specify both max_mismatch and min_identity")
    if min_identity is not None and not 0.0 <= min_identity <= 1.0:
        raise ValueError("min_identity must be between 0.0 and 1.0")
    m = Monomerizer(seed_length, max_mismatch, min_identity)
    out, table = bytearray(), (bytearray() if table_ext is not None else None)
    delim = b"\t" if table_ext == "tsv" else b","
    for r in parse_fasta(fasta):
        norm = normalize(r.seq)
        idx = None
        if not (len(norm) < seed_length or len(norm) < min_length):
            idx = m.last_monomer_end_index_sensitive(norm) if sensitive else m.last_monomer_end_index(norm)
        fs = full_seq(r.seq)
        if idx is not None and (idx < min_length or (max_length is not None and idx > max_length)):
            idx = None
        if min_overlap is not None and idx is not None and len(fs) - idx < min_overlap:
            idx = None
        if min_overlap_percent is not None and idx is not None:
            ratio = (len(fs) - idx) / idx if idx else (float("inf") if len(fs) else float("nan"))
            if ratio < min_overlap_percent:
                idx = None
        if idx is not None or keep_all:
            end = len(fs) if idx is None else idx
            out += b">" + r.head + b"\n" + fs[:end] + b"\n"
            if table is not None:
                if not table:
                    table += delim.join([b"id", b"original_length", b"monomer_length"]) + b"\n"
                table += delim.join([_csv_field(r.head, delim), b"%d" % len(fs), b"%d" % end]) + b"\n"
    return bytes(out), (bytes(table) if table is not None else None)
