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
