"""Host side of the CLI drivers (circkit_b200/cli.py) without a GPU: FASTA record splitting against the oracle's
restatement of seq_io's rules, output framing, (de)compression sniffing."""
import bz2
import gzip
import lzma
import os
import random

import numpy as np
import pytest

from circkit_b200 import cli
from oracle import cli as ocli

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, "golden", "fixtures")


def _records(data):
    r = cli.Records(data)
    d = np.frombuffer(data, dtype=np.uint8)
    return [(d[a:b].tobytes(), d[c:e].tobytes()) for a, b, c, e in
            zip(r.head_lo.tolist(), r.head_hi.tolist(), r.seq_lo.tolist(), r.seq_hi.tolist())]


CASES = [b"", b"\n\n", b">a\nACGT", b">a\nACGT\n", b">a b c\nAC\nGT\n>b\nTT\n\n", b"\r\n>x\r\nAC\r\nGT\r\n>y\r\nA\r\n",
         b">only_header", b">h1\n>h2\nAC\n", b">a\nAC>GT\n>b\nA\n", b"\n\r\n>a\nA", b">a\n\n>b\n\nAC\n\n", b">\n\n", b">a\nAC\r"]


@pytest.mark.parametrize("data", CASES)
def test_record_splitting_matches_the_oracle_parser(data):
    assert _records(data) == [(r.head, r.seq) for r in ocli.parse_fasta(data)]


def test_record_splitting_on_the_fixtures_and_random_files():
    for d in sorted(os.listdir(FIX)):
        p = os.path.join(FIX, d, "in.fasta")
        if os.path.isfile(p):
            data = open(p, "rb").read()
            assert _records(data) == [(r.head, r.seq) for r in ocli.parse_fasta(data)], d
    rng = random.Random(5)
    for _ in range(300):
        parts = []
        for _ in range(rng.randint(0, 6)):
            parts.append(rng.choice([b">", b">id ", b">x\r", b"AC", b"GT\r", b"\n", b"\n", b"\r\n", b"N", b" ", b">>"]))
        data = b">" + b"".join(parts) if rng.random() < 0.7 else b"\n" * rng.randint(0, 2) + b">" + b"".join(parts)
        assert _records(data) == [(r.head, r.seq) for r in ocli.parse_fasta(data)], data


def test_bad_start_is_an_error():
    with pytest.raises(cli.FastaError):
        cli.Records(b"ACGT\n>a\nAC\n")


def test_assemble_framing():
    heads = np.frombuffer(b"ab c", dtype=np.uint8)
    bodies = np.frombuffer(b"ACGTTT", dtype=np.uint8)
    out = cli._assemble(heads, np.array([1, 3]), bodies, np.array([4, 2]))
    assert out == b">a\nACGT\n>b c\nTT\n"
    assert cli._assemble(heads[:0], np.array([], dtype=np.int64), bodies[:0], np.array([], dtype=np.int64)) == b""


def test_compressed_inputs_are_sniffed(tmp_path):
    plain = open(os.path.join(FIX, "compressed_input", "in.fasta"), "rb").read()
    for ext in ("gz", "bz2", "xz", "zst"):
        assert cli.read_input(os.path.join(FIX, "compressed_input", "in.fasta." + ext)) == plain, ext
    for ext, dec in (("gz", gzip.decompress), ("bz2", bz2.decompress), ("xz", lzma.decompress), ("zst", cli._zstd_decompress),
                     ("fasta", lambda b: b)):
        p = str(tmp_path / ("out." + ext))
        cli.write_output(p, plain)
        assert dec(open(p, "rb").read()) == plain, ext


def test_missing_input_exits_non_zero(capsys):
    assert cli.main(["canonicalize", "/nonexistent/in.fasta"]) == 1
    assert "No such file or directory" in capsys.readouterr().err       # tests/canon_uniq.rs:8-17


def test_streaming_reader_splits_like_the_whole_file_parser(tmp_path):
    """iter_input + iter_records (pieces of a few hundred bytes) must see exactly the records of a whole-buffer parse, for
    every hot-path fixture, compressed inputs included, and for records that straddle many pieces"""
    import random
    rng = random.Random(4)
    big = b"".join(b">r%d some description\r\n" % i + b"\n".join(bytes(rng.choice(b"ACGTN") for _ in range(rng.randint(0, 70)))
                                                                  for _ in range(rng.randint(0, 6))) + b"\n" for i in range(300))
    cases = {"big": big, "no_final_newline": big[:-1], "leading_blank": b"\n\r\n" + big[:2000]}
    for d in sorted(os.listdir(FIX)):
        p = os.path.join(FIX, d, "in.fasta")
        if os.path.exists(p):
            cases[d] = open(p, "rb").read()
    for name, data in cases.items():
        want = _records(data)
        for chunk in (7, 64, 1000, 1 << 20):
            f = tmp_path / "in.fa"
            f.write_bytes(data)
            got = []
            for recs in cli.iter_records(cli.iter_input(str(f), chunk_bytes=chunk)):
                for a, b, c, e in zip(recs.head_lo, recs.head_hi, recs.seq_lo, recs.seq_hi):
                    got.append((recs.data[a:b].tobytes(), recs.data[c:e].tobytes()))
            assert got == want, (name, chunk)
    plain = open(os.path.join(FIX, "compressed_input", "out.fasta"), "rb").read() if False else None
    for ext in ("gz", "bz2", "xz", "zst"):
        p = os.path.join(FIX, "compressed_input", "in.fasta." + ext)
        data = cli.read_input(p)
        got = b"".join(cli.iter_input(p, chunk_bytes=13))
        assert got == data, ext


def test_streaming_writer_round_trips_every_suffix(tmp_path):
    pieces = [b">a\nACGT\n", b"", b">b\n" + b"ACGT" * 5000 + b"\n", b">c\nTTTT\n"]
    for ext in ("fa", "gz", "bz2", "xz", "zst"):
        p = str(tmp_path / ("out." + ext))
        w = cli.Writer(p)
        for x in pieces:
            w.write(x)
        w.close()
        assert cli.read_input(p) == b"".join(pieces), ext
