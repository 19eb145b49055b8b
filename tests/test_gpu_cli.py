"""GPU parity of the CLI drivers (circkit_b200/cli.py): byte-identical output to the oracle's restatement of
`circkit canonicalize` / `circkit uniq` (src/canonicalize.rs:7-51, src/uniq.rs:15-88) on every hot-path fixture of the
reference (tests/examples, goldens under tests/golden/oracle_cli) and on a larger synthetic FASTA file."""
import os
import random

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, "golden", "fixtures")
GOLD = os.path.join(HERE, "golden", "oracle_cli")
FIXTURES = sorted(d for d in os.listdir(FIX) if os.path.isfile(os.path.join(FIX, d, "in.fasta")))


def _gold(name):
    return open(os.path.join(GOLD, name), "rb").read()


@pytest.mark.parametrize("d", FIXTURES)
def test_fixture_outputs_are_byte_identical(d):
    from circkit_b200 import cli
    data = open(os.path.join(FIX, d, "in.fasta"), "rb").read()
    assert cli.canonicalize(data) == _gold(d + ".canonicalize.out")
    out, table = cli.uniq(data, canonical=False, table_ext="csv")
    assert out == _gold(d + ".uniq.out")
    assert table == _gold(d + ".uniq.table.csv")
    out, table = cli.uniq(data, canonical=True, table_ext="csv")
    assert out == _gold(d + ".uniq_c.out")
    assert table == _gold(d + ".uniq.table.csv")


def test_command_line_files_compression_and_silence(tmp_path, capsys):
    """tests/canon_uniq.rs:33-89 and tests/compression.rs: file in, file out (plain and compressed by suffix), compressed
    input, --table; nothing on stdout / stderr."""
    from circkit_b200 import cli
    src = os.path.join(FIX, "compressed_input")
    want = _gold("compressed_input.canonicalize.out")
    for ext in ("", ".gz", ".bz2", ".xz", ".zst"):
        out = str(tmp_path / ("o" + ext.replace(".", "_") + ".fasta"))
        assert cli.main(["canonicalize", os.path.join(src, "in.fasta" + ext), "-o", out, "-t", "2"]) == 0
        assert open(out, "rb").read() == want
    for ext in ("gz", "bz2", "xz", "zst"):
        out = str(tmp_path / ("c.fasta." + ext))
        assert cli.main(["canonicalize", os.path.join(src, "in.fasta"), "-o", out]) == 0
        assert cli.read_input(out) == want
    rep = os.path.join(FIX, "repeated", "in.fasta")
    out, tab = str(tmp_path / "u.fasta"), str(tmp_path / "t.tsv")
    assert cli.main(["uniq", rep, "-o", out, "-c", "--table", tab]) == 0
    assert open(out, "rb").read() == _gold("repeated.uniq_c.out")
    assert open(tab, "rb").read() == _gold("repeated.uniq.table.csv").replace(b",", b"\t")
    cap = capsys.readouterr()
    assert cap.out == "" and cap.err == ""
    # stdout when there is no -o (tests/canon_uniq.rs:19-31)
    p = tmp_path / "s.fasta"
    p.write_bytes(b">seq1\nATGCA")
    assert cli.main(["canonicalize", str(p)]) == 0
    assert ">seq1\nAATGC" in capsys.readouterr().out


def test_synthetic_file_matches_the_oracle_cli():
    """20 k records: wrapped lines, CRLF, lower case, RNA, IUPAC, duplicates as rotations / reverse complements."""
    from circkit_b200 import cli
    import oracle
    from oracle import cli as ocli
    rng = random.Random(17)
    recs, lines = [], []
    for i in range(20000):
        if recs and rng.random() < 0.3:
            s = rng.choice(recs)
            r = rng.randrange(len(s))
            s = s[r:] + s[:r]
            if rng.random() < 0.5:
                s = oracle.revcomp(s)
        else:
            s = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(1, 600)))
            recs.append(s)
        if rng.random() < 0.1:
            s = s.lower()
        if rng.random() < 0.05:
            s = s.replace(b"T", b"U")
        if rng.random() < 0.05:
            k = rng.randrange(len(s))
            s = s[:k] + rng.choice([b"N", b"R", b"y", b"-", b"."]) + s[k + 1:]
        eol = b"\r\n" if rng.random() < 0.1 else b"\n"
        w = rng.choice([0, 60, 70])
        body = eol.join(s[k:k + w] for k in range(0, len(s), w)) if w else s
        lines.append(b">r%d some description" % i + eol + body + eol)
    data = b"".join(lines)
    assert cli.canonicalize(data) == ocli.cli_canonicalize(data, threads=4)
    for canon in (False, True):
        got, gt = cli.uniq(data, canonical=canon, table_ext="csv")
        want, wt = ocli.cli_uniq(data, canonicalize=canon, table_ext="csv", threads=4)
        assert got == want and gt == wt
