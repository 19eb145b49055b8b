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
