"""`python -m circkit_b200 monomerize` (circkit_b200/cli.py: normalise + monomer search on the device, filters and framing
on the host) against the reference's own CLI fixtures (tests/monomerize.rs:46-86) and byte for byte against the oracle's
restatement of the driver on option combinations and awkward framing."""
import os
import random

import pytest

from oracle import cli as ocli

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, "golden", "monomerize_fixtures")
CASES = [("min_overlap", dict(min_overlap=81)), ("min_overlap_percent_0.51", dict(min_overlap_percent=0.51)),
         ("min_overlap_percent_1.0", dict(min_overlap_percent=1.0)), ("min_overlap_percent_1.5", dict(min_overlap_percent=1.5))]


def seq_map(fasta: bytes):
    return {r.id: ocli.full_seq(r.seq) for r in ocli.parse_fasta(fasta)}


@pytest.mark.parametrize("d,kw", CASES)
def test_reference_cli_fixtures(d, kw):
    from circkit_b200 import cli
    data = open(os.path.join(FIX, d, "in.fasta"), "rb").read()
    out, table = cli.monomerize(data, **kw)
    assert seq_map(out) == seq_map(open(os.path.join(FIX, d, "out.fasta"), "rb").read()) and table is None
    assert out == ocli.cli_monomerize(data, **kw)[0]


def _random_fasta(rng, n):
    recs = []
    for i in range(n):
        unit = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(8, 120)))
        kind = rng.randrange(5)
        if kind == 0:
            s = unit * rng.randint(2, 3)
        elif kind == 1:
            s = unit * 2 + unit[: rng.randint(0, len(unit))]
        elif kind == 2:
            s = bytearray(unit * 2)
            s[rng.randrange(len(s))] = rng.choice(b"ACGTN")
            s = bytes(s)
        elif kind == 3:
            s = bytes(rng.choice(b"ACGTacgtnNRYu") for _ in range(rng.randint(0, 200)))
        else:
            s = unit
        width = rng.choice([0, 7, 60])
        eol = rng.choice([b"\n", b"\r\n"])
        body = s if not width else eol.join(s[k: k + width] for k in range(0, len(s), width))
        recs.append(b">r%d some, \"text\"" % i + eol + body + (eol if body or rng.random() < 0.5 else b""))
    data = b"".join(recs)
    return data if data.endswith(b"\n") or rng.random() < 0.5 else data + b"\n"


@pytest.mark.parametrize("kw", [
    dict(), dict(sensitive=True), dict(seed_length=5, max_mismatch=2), dict(min_identity=0.9, keep_all=True, table_ext="csv"),
    dict(seed_length=12, min_identity=0.95, min_length=30, max_length=100, table_ext="tsv"),
    dict(min_overlap=20, min_overlap_percent=0.6, keep_all=True, table_ext="csv"), dict(seed_length=63, keep_all=True),
])
def test_options_and_framing_match_the_oracle_driver(kw):
    from circkit_b200 import cli
    rng = random.Random(len(repr(kw)))
    data = _random_fasta(rng, 400)
    assert cli.monomerize(data, **kw) == ocli.cli_monomerize(data, **kw)


def test_main_entry_and_errors(tmp_path):
    from circkit_b200 import cli
    x = b"ATGACAGGTACAGCATA"
    src = tmp_path / "in.fa"; dst = tmp_path / "out.fa"; tab = tmp_path / "t.tsv"
    src.write_bytes(b">d\n" + x + x + b"\n>p\nTTTTTTTTTTTTAAAAAAAAAA\n")
    assert cli.main(["monomerize", str(src), "-o", str(dst), "--seed-length", "6", "--table", str(tab), "-k"]) == 0
    assert dst.read_bytes() == b">d\n" + x + b"\n>p\nTTTTTTTTTTTTAAAAAAAAAA\n"
    assert tab.read_bytes() == b"id\toriginal_length\tmonomer_length\nd\t34\t17\np\t22\t22\n"
    assert cli.main(["monomerize", str(src), "-o", str(dst), "--max-mismatch", "1", "--min-identity", "0.9"]) == 1
    assert cli.main(["monomerize", str(src), "-o", str(dst), "--min-identity", "1.5"]) == 1
    assert cli.monomerize(b"", table_ext="csv") == (b"", b"")
