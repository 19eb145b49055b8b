"""`circkit canonicalize` / `circkit uniq` on the B200 path: the host side of the two CLI drivers.

    python -m circkit_b200 canonicalize [INPUT] [-o OUTPUT] [-t THREADS]
    python -m circkit_b200 uniq [INPUT] [-o OUTPUT] [-c] [--table TABLE] [-t THREADS]

Mirrors src/canonicalize.rs:7-51 and src/uniq.rs:15-88 with the flags of src/commands.rs:112-149: the host keeps what
the reference keeps on the host -- reading (with the compression sniffing of src/utils.rs:9-27), FASTA record splitting
(seq_io 0.3.2 rules), writing in the reference's framing (src/canonicalize.rs:33-37, src/uniq.rs:50-61), the duplicate
table (src/uniq.rs:62-71, src/utils.rs:74-84) and output compression by suffix (src/utils.rs:29-72) -- and hands batches
of raw record bytes to the C ABI (ck_canon_submit / ck_uniq_submit, two slots in flight), where normalisation, LMSR,
the canonical form, XXH3-64 and the first-occurrence table run on the GPU.  `--threads` is accepted and ignored.
Nothing is printed on success (tests/canon_uniq.rs:74-77); an I/O error exits non-zero with the OS message on stderr.

Record splitting and output assembly are numpy-vectorised (no per-record Python on the canonicalize path).
"""
from __future__ import annotations

import argparse
import bz2
import ctypes
import ctypes.util
import gzip
import lzma
import os
import sys

import numpy as np

from .core import Context


class FastaError(ValueError):
    pass


# ------------------------------------------------------------------------------------------------ (de)compression
def _zstd():
    name = ctypes.util.find_library("zstd") or "libzstd.so.1"
    lib = ctypes.CDLL(name)
    lib.ZSTD_getFrameContentSize.restype = ctypes.c_ulonglong
    lib.ZSTD_getFrameContentSize.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    lib.ZSTD_decompress.restype = ctypes.c_size_t
    lib.ZSTD_decompress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
    lib.ZSTD_compressBound.restype = ctypes.c_size_t
    lib.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
    lib.ZSTD_compress.restype = ctypes.c_size_t
    lib.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    lib.ZSTD_isError.restype = ctypes.c_uint
    lib.ZSTD_isError.argtypes = [ctypes.c_size_t]
    return lib


def _zstd_decompress(raw: bytes) -> bytes:
    lib = _zstd()
    size = lib.ZSTD_getFrameContentSize(raw, len(raw))
    if size in (2**64 - 1, 2**64 - 2):                  # error / unknown size: grow until it fits
        size = max(4 * len(raw), 1 << 20)
    while True:
        buf = ctypes.create_string_buffer(int(size) or 1)
        got = lib.ZSTD_decompress(buf, len(buf), raw, len(raw))
        if not lib.ZSTD_isError(got):
            return buf.raw[:got]
        if size > (1 << 36):
            raise OSError("zstd: cannot decompress input")
        size *= 4


def _zstd_compress(data: bytes, level: int) -> bytes:
    lib = _zstd()
    buf = ctypes.create_string_buffer(lib.ZSTD_compressBound(len(data)) or 1)
    got = lib.ZSTD_compress(buf, len(buf), data, len(data), level)
    if lib.ZSTD_isError(got):
        raise OSError("zstd: cannot compress output")
    return buf.raw[:got]


def read_input(path: str | None) -> bytes:
    """src/utils.rs:9-27: file or stdin (refused when it is a terminal), format sniffed from the magic bytes (niffler)."""
    if path is None:
        if sys.stdin.isatty():
            raise OSError("No input file specified and stdin is a terminal")
        raw = sys.stdin.buffer.read()
    else:
        with open(path, "rb") as f:
            raw = f.read()
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    if raw[:3] == b"BZh":
        return bz2.decompress(raw)
    if raw[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(raw)
    if raw[:4] == b"\x28\xb5\x2f\xfd":
        return _zstd_decompress(raw)
    return raw


def write_output(path: str | None, data: bytes) -> None:
    """src/utils.rs:29-72: stdout, or a file compressed by its suffix (gz 6, bz2 9, xz 6, zst 1; anything else plain)."""
    if path is None:
        sys.stdout.buffer.write(data)
        sys.stdout.buffer.flush()
        return
    ext = path.rsplit(".", 1)[-1] if "." in os.path.basename(path) else ""
    if ext == "gz":
        data = gzip.compress(data, compresslevel=6)
    elif ext == "bz2":
        data = bz2.compress(data, compresslevel=9)
    elif ext == "xz":
        data = lzma.compress(data, preset=6)
    elif ext == "zst":
        data = _zstd_compress(data, 1)
    with open(path, "wb") as f:
        f.write(data)


# ------------------------------------------------------------------------------------------------ FASTA records
class Records:
    """Record boundaries of a FASTA buffer as index arrays (seq_io 0.3.2 `fasta::Reader` rules): a record starts at a
    '>' that follows a '\\n'; head = header line without '>' and one trailing '\\r'; seq = everything between the
    header's '\\n' and the record's last '\\n' (or EOF), internal line breaks included, one trailing '\\r' trimmed."""

    def __init__(self, data: bytes):
        d = np.frombuffer(data, dtype=np.uint8)
        n = len(d)
        self.data = d
        nl = np.flatnonzero(d == 10)
        # leading empty lines ("" or "\r") are skipped; the first other line must start with '>'
        pos = 0
        while pos < n:
            k = int(np.searchsorted(nl, pos))
            end = int(nl[k]) if k < len(nl) else n
            line_len = end - pos
            if not (line_len == 0 or (line_len == 1 and d[pos] == 13)):
                break
            if k >= len(nl):
                pos = n
                break
            pos = end + 1
        empty = np.zeros(0, dtype=np.int64)
        if pos >= n:
            self.head_lo = self.head_hi = self.seq_lo = self.seq_hi = empty
            return
        if d[pos] != ord(">"):
            raise FastaError("expected '>' at record start")
        gt = np.flatnonzero(d == ord(">"))
        gt = gt[gt > pos]
        starts = np.concatenate([[pos], gt[d[gt - 1] == 10]]).astype(np.int64)
        # region end: the '\n' before the next record's '>' (excluded), else EOF without one final '\n'
        last_end = n - 1 if d[n - 1] == 10 else n
        region_end = np.concatenate([starts[1:] - 1, [last_end]]).astype(np.int64)
        if len(nl):
            k = np.searchsorted(nl, starts)
            first_nl = np.where(k < len(nl), nl[np.minimum(k, len(nl) - 1)], n).astype(np.int64)
        else:
            first_nl = np.full(len(starts), n, dtype=np.int64)
        has_seq = first_nl < region_end
        head_hi = np.where(has_seq, first_nl, region_end)
        head_lo = starts + 1
        head_hi = np.maximum(head_hi, head_lo)
        cr = (head_hi > head_lo) & (d[np.maximum(head_hi - 1, 0)] == 13)
        self.head_lo, self.head_hi = head_lo, head_hi - cr
        seq_lo = np.where(has_seq, first_nl + 1, region_end)
        seq_hi = np.maximum(region_end, seq_lo)
        cr = (seq_hi > seq_lo) & (d[np.maximum(seq_hi - 1, 0)] == 13)
        self.seq_lo, self.seq_hi = seq_lo, seq_hi - cr

    def __len__(self):
        return len(self.head_lo)

    def gather(self, lo: np.ndarray, hi: np.ndarray) -> np.ndarray:
        """concatenation of data[lo[i]:hi[i]] over i (regions are disjoint and increasing)"""
        return self.data[_region_mask(len(self.data), lo, hi)]

    def ids(self) -> list[bytes]:
        """record.id(): head up to the first ' ' (src/uniq.rs:48,67 unwrap it as UTF-8)"""
        out = []
        for a, b in zip(self.head_lo.tolist(), self.head_hi.tolist()):
            h = self.data[a:b].tobytes().split(b" ", 1)[0]
            h.decode("utf-8")
            out.append(h)
        return out


def _region_mask(n: int, lo: np.ndarray, hi: np.ndarray) -> np.ndarray:
    delta = np.zeros(n + 1, dtype=np.int32)
    np.add.at(delta, lo, 1)
    np.add.at(delta, hi, -1)
    return np.cumsum(delta[:-1]) > 0


def _assemble(heads: np.ndarray, head_len: np.ndarray, bodies: np.ndarray, body_len: np.ndarray) -> bytes:
    """'>' head '\\n' body '\\n' for every record (src/canonicalize.rs:33-37), from the concatenated heads / bodies."""
    m = len(head_len)
    if m == 0:
        return b""
    rec_len = head_len + body_len + 3
    start = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(rec_len, out=start[1:])
    out = np.empty(int(start[-1]), dtype=np.uint8)
    out[start[:-1]] = ord(">")
    out[start[:-1] + 1 + head_len] = 10
    out[start[1:] - 1] = 10
    out[_region_mask(len(out), start[:-1] + 1, start[:-1] + 1 + head_len)] = heads
    out[_region_mask(len(out), start[:-1] + 2 + head_len, start[1:] - 1)] = bodies
    return out.tobytes()


# ------------------------------------------------------------------------------------------------ GPU pump
MAX_BATCH_BYTES = 256 << 20
MAX_BATCH_RECORDS = 1 << 20


def _batches(offsets: np.ndarray):
    """cuts of the record range into batches of <= MAX_BATCH_BYTES / MAX_BATCH_RECORDS"""
    n = len(offsets) - 1
    cuts = [0]
    while cuts[-1] < n:
        lo = cuts[-1]
        hi = min(n, lo + MAX_BATCH_RECORDS)
        limit = int(offsets[lo]) + MAX_BATCH_BYTES
        if int(offsets[hi]) > limit:
            hi = int(np.searchsorted(offsets, limit, side="right")) - 1
        if hi <= lo:
            raise FastaError("record longer than the batch size")
        cuts.append(hi)
    return cuts


def _pump(ctx: Context, arena: np.ndarray, offsets: np.ndarray, uniq: bool, want_bytes: bool = True):
    """All batches through the two slots of the C ABI (the bounded queue of parallel_fasta): submit batch b while batch
    b-1 is collected.  Returns (canonical bytes, normalised lengths, first_index or None), compact layout."""
    n = len(offsets) - 1
    cuts = _batches(offsets)
    lens = np.zeros(n, dtype=np.uint32)
    first = np.zeros(n, dtype=np.uint64) if uniq else None
    outs = []
    nb = len(cuts) - 1
    for b in range(nb + 1):
        if b < nb:
            lo, hi = cuts[b], cuts[b + 1]
            rel = (offsets[lo: hi + 1] - offsets[lo]).astype(np.uint64)
            sub = arena[int(offsets[lo]): int(offsets[hi])]
            if uniq:
                ctx.uniq_submit(b & 1, sub, rel, lo, normalize=True, no_bytes=not want_bytes, aligned=True)
            else:
                ctx.canon_submit(b & 1, sub, rel, normalize=True, aligned=True)
        if b >= 1:
            lo, hi = cuts[b - 1], cuts[b]
            rel = (offsets[lo: hi + 1] - offsets[lo]).astype(np.uint64)
            total = int(rel[-1])
            if uniq:
                r = ctx.uniq_wait((b - 1) & 1, hi - lo, total, want_bytes=want_bytes, aligned=True)
                first[lo:hi] = r["first"]
            else:
                r = ctx.canon_wait((b - 1) & 1, hi - lo, total, aligned=True)
            lens[lo:hi] = r["lens"]
            if want_bytes:
                st = ctx.aligned_starts(rel).astype(np.int64)
                outs.append(r["out"][_region_mask(len(r["out"]), st, st + r["lens"].astype(np.int64))])
    body = np.concatenate(outs) if outs else np.zeros(0, dtype=np.uint8)
    return body, lens, first


def _context(n_records: int, uniq: bool) -> Context:
    return Context(max_batch_bytes=MAX_BATCH_BYTES, max_batch_records=MAX_BATCH_RECORDS,
                   table_capacity=max(n_records, 1) if uniq else 0)


def canonicalize(fasta: bytes) -> bytes:
    """bytes of a FASTA file -> bytes `circkit canonicalize` writes (src/canonicalize.rs:7-51)"""
    recs = Records(fasta)
    if len(recs) == 0:
        return b""
    seq_len = recs.seq_hi - recs.seq_lo
    offsets = np.zeros(len(recs) + 1, dtype=np.uint64)
    np.cumsum(seq_len, out=offsets[1:])
    arena = np.ascontiguousarray(recs.gather(recs.seq_lo, recs.seq_hi))
    ctx = _context(len(recs), False)
    try:
        body, lens, _ = _pump(ctx, arena, offsets, False)
    finally:
        ctx.close()
    heads = recs.gather(recs.head_lo, recs.head_hi)
    return _assemble(heads, recs.head_hi - recs.head_lo, body, lens.astype(np.int64))


def uniq(fasta: bytes, canonical: bool = False, table_ext: str | None = None):
    """bytes of a FASTA file -> (bytes `circkit uniq` writes, bytes of the --table file or None) (src/uniq.rs:15-88)"""
    recs = Records(fasta)
    table = bytearray() if table_ext is not None else None
    if len(recs) == 0:
        return b"", (bytes(table) if table is not None else None)
    ids = recs.ids()                                        # unwrap()s of src/uniq.rs:48,67 happen for every record
    seq_len = recs.seq_hi - recs.seq_lo
    offsets = np.zeros(len(recs) + 1, dtype=np.uint64)
    np.cumsum(seq_len, out=offsets[1:])
    arena = np.ascontiguousarray(recs.gather(recs.seq_lo, recs.seq_hi))
    ctx = _context(len(recs), True)
    try:
        body, lens, first = _pump(ctx, arena, offsets, True, want_bytes=canonical)
    finally:
        ctx.close()
    keep = first == np.arange(len(recs), dtype=np.uint64)   # first occurrence in input order (src/uniq.rs:47-48)
    head_len = (recs.head_hi - recs.head_lo)[keep]
    heads = recs.gather(recs.head_lo[keep], recs.head_hi[keep])
    if canonical:                                           # src/uniq.rs:53-56
        cstart = np.zeros(len(recs) + 1, dtype=np.int64)
        np.cumsum(lens, out=cstart[1:])
        bodies = body[_region_mask(len(body), cstart[:-1][keep], cstart[1:][keep])]
        body_len = lens.astype(np.int64)[keep]
    else:                                                   # raw record.seq(), internal line breaks included (:57-59)
        bodies = recs.gather(recs.seq_lo[keep], recs.seq_hi[keep])
        body_len = seq_len[keep]
    out = _assemble(heads, head_len, bodies, body_len)
    if table is not None:                                   # src/uniq.rs:62-71, csv 1.2.2 defaults
        delim = b"\t" if table_ext == "tsv" else b","

        def field(f: bytes) -> bytes:
            if any(c in f for c in (delim, b'"', b"\n", b"\r")):
                return b'"' + f.replace(b'"', b'""') + b'"'
            return f
        dups = np.flatnonzero(~keep)
        if len(dups):
            table += b"id" + delim + b"duplicate_id\n"
        for i in dups.tolist():
            table += field(ids[int(first[i])]) + delim + field(ids[i]) + b"\n"
    return out, (bytes(table) if table is not None else None)


class CliError(ValueError):
    """argument combinations the reference rejects with `bail!` (src/monomerize.rs:35-45)"""


def monomerize(fasta: bytes, sensitive: bool = False, seed_length: int = 10, max_mismatch=None, min_identity=None,
               min_overlap=None, min_overlap_percent=None, min_length: int = 0, max_length=None, keep_all: bool = False,
               table_ext: str | None = None):
    """bytes of a FASTA file -> (bytes `circkit monomerize` writes, bytes of the --table file or None)
    (src/monomerize.rs:16-160).  Normalisation (`ck_dev_normalize`) and the monomer search (`ck_dev_monomerize`) run on
    the device over the whole file as one batch; the length / overlap filters of the consumer closure (:102-135) and the
    framing (:139-143: '>' head, full_seq[..end]) are vectorised on the host."""
    import torch
    from .monomerize import Monomerizer
    if max_mismatch is not None and min_identity is not None:
        raise CliError("cannot specify both max_mismatch and min_identity")
    if min_identity is not None and not 0.0 <= min_identity <= 1.0:
        raise CliError("min_identity must be between 0.0 and 1.0")
    recs = Records(fasta)
    n = len(recs)
    table = bytearray() if table_ext is not None else None
    if n == 0:
        return b"", (bytes(table) if table is not None else None)
    seq_len = (recs.seq_hi - recs.seq_lo).astype(np.int64)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(seq_len, out=off[1:])
    arena = np.ascontiguousarray(recs.gather(recs.seq_lo, recs.seq_hi))
    m = Monomerizer(seed_length, max_mismatch, min_identity)
    ctx = m._context()
    try:
        dev = torch.device("cuda", ctx.device)
        raw = torch.from_numpy(arena if len(arena) else np.zeros(1, dtype=np.uint8)).to(dev)
        offs = torch.from_numpy(off).to(dev)
        norm, nlens = m.normalize_device(raw, offs, n)
        end = m.end_indices_device(norm, offs, n, sensitive=sensitive, lens=nlens)
        idx = end.cpu().numpy().astype(np.int64)             # -1 = None
        nl = nlens.cpu().numpy().astype(np.int64)
    finally:
        ctx.close()
    idx[(nl < seed_length) | (nl < min_length)] = -1        # worker closure, src/monomerize.rs:92-96
    # record.full_seq(): the lines of seq joined, each without its '\n' and one trailing '\r'
    is_nl, is_cr = arena == 10, arena == 13
    before_nl = np.append(is_nl[1:], False)
    at_end = np.zeros(len(arena), dtype=bool)
    nonempty = seq_len > 0
    at_end[off[1:][nonempty] - 1] = True
    kept = ~(is_nl | (is_cr & (before_nl | at_end)))
    c = np.zeros(len(arena) + 1, dtype=np.int64)
    np.cumsum(kept, out=c[1:])
    full_len = c[off[1:]] - c[off[:-1]]
    ok = idx >= 0                                            # consumer closure, :107-135
    ok &= ~(idx < min_length)
    if max_length is not None:
        ok &= ~(idx > max_length)
    if min_overlap is not None:
        ok &= ~((full_len - idx) < min_overlap)
    if min_overlap_percent is not None:
        with np.errstate(divide="ignore", invalid="ignore"):
            ok &= ~(((full_len - idx).astype(np.float64) / idx.astype(np.float64)) < min_overlap_percent)
    write = ok | keep_all
    end_idx = np.minimum(np.where(ok, idx, full_len), full_len)
    rec_of = np.repeat(np.arange(n, dtype=np.int64), seq_len)
    rank = c[:-1] - c[off[:-1]][rec_of]
    sel = kept & (rank < end_idx[rec_of]) & write[rec_of]
    heads = recs.gather(recs.head_lo[write], recs.head_hi[write])
    out = _assemble(heads, (recs.head_hi - recs.head_lo)[write], arena[sel], end_idx[write])
    if table is not None:                                    # :145-154, csv 1.2.2 defaults: header with the first row
        delim = b"\t" if table_ext == "tsv" else b","

        def field(f: bytes) -> bytes:
            if any(x in f for x in (delim, b'"', b"\n", b"\r")):
                return b'"' + f.replace(b'"', b'""') + b'"'
            return f
        rows = np.flatnonzero(write)
        if len(rows):
            table += delim.join([b"id", b"original_length", b"monomer_length"]) + b"\n"
        for i in rows.tolist():
            head = recs.data[recs.head_lo[i]: recs.head_hi[i]].tobytes()
            head.decode("utf-8")                             # std::str::from_utf8(record.head()).unwrap()
            table += delim.join([field(head), b"%d" % full_len[i], b"%d" % end_idx[i]]) + b"\n"
    return out, (bytes(table) if table is not None else None)


# ------------------------------------------------------------------------------------------------ entry point
def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="circkit", description="canonicalize / uniq on the B200 path")
    ap.add_argument("-v", "--verbose", action="count", default=0)
    ap.add_argument("-q", "--quiet", action="count", default=0)
    sub = ap.add_subparsers(dest="command", required=True)
    c = sub.add_parser("canonicalize", aliases=["canon"])
    c.add_argument("input", nargs="?")
    c.add_argument("-o", "--output")
    c.add_argument("-t", "--threads", type=int, default=os.cpu_count())
    u = sub.add_parser("uniq")
    u.add_argument("input", nargs="?")
    u.add_argument("-o", "--output")
    u.add_argument("-c", "--canonicalize", "--norm", "--canon", dest="canonicalize", action="store_true")
    u.add_argument("--table")
    u.add_argument("-t", "--threads", type=int, default=os.cpu_count())
    mo = sub.add_parser("monomerize")                          # src/commands.rs:19-90
    mo.add_argument("input", nargs="?")
    mo.add_argument("-o", "--output")
    mo.add_argument("--sensitive", action="store_true")
    mo.add_argument("--seed-length", type=int, default=10, choices=range(5, 65), metavar="[5..64]")
    mo.add_argument("--max-mismatch", type=int)
    mo.add_argument("--min-identity", type=float)
    mo.add_argument("--min-overlap", type=int)
    mo.add_argument("--min-overlap-percent", type=float)
    mo.add_argument("--min-length", type=int, default=0)
    mo.add_argument("--max-length", type=int)
    mo.add_argument("-k", "--keep-all", action="store_true")
    mo.add_argument("--table")
    mo.add_argument("-t", "--threads", type=int, default=os.cpu_count())
    mo.add_argument("--batch-size", type=int, default=64, help=argparse.SUPPRESS)
    args = ap.parse_args(argv)
    try:
        data = read_input(args.input)
        if args.command in ("canonicalize", "canon"):
            write_output(args.output, canonicalize(data))
        elif args.command == "monomerize":
            ext = None
            if args.table is not None:
                ext = "tsv" if args.table.endswith(".tsv") else "csv"
            out, table = monomerize(data, args.sensitive, args.seed_length, args.max_mismatch, args.min_identity, args.min_overlap,
                                    args.min_overlap_percent, args.min_length, args.max_length, args.keep_all, ext)
            write_output(args.output, out)
            if args.table is not None:
                with open(args.table, "wb") as f:
                    f.write(table)
        else:
            ext = None
            if args.table is not None:
                ext = "tsv" if args.table.endswith(".tsv") else "csv"
            out, table = uniq(data, args.canonicalize, ext)
            write_output(args.output, out)
            if args.table is not None:
                with open(args.table, "wb") as f:
                    f.write(table)
    except OSError as e:
        msg = e.strerror or str(e)
        print("Error: %s%s" % (msg, " (os error %d)" % e.errno if e.errno else ""), file=sys.stderr)
        return 1
    except (FastaError, UnicodeDecodeError, CliError, ValueError) as e:
        print("Error: %s" % e, file=sys.stderr)
        return 1
    return 0
