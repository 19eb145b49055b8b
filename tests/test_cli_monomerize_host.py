"""`circkit monomerize` (src/monomerize.rs:16-160): the oracle's restatement of the driver (oracle/cli.py) against the four
fixtures of the reference's own CLI tests (tests/monomerize.rs:46-86: --min-overlap 81, --min-overlap-percent 0.51 / 1.0 /
1.5; compared like tests/common.rs `sequences_are_identical`: id -> sequence maps), and argument rules."""
import os

import pytest

from oracle import cli as ocli

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, "golden", "monomerize_fixtures")
CASES = [("min_overlap", dict(min_overlap=81)), ("min_overlap_percent_0.51", dict(min_overlap_percent=0.51)),
         ("min_overlap_percent_1.0", dict(min_overlap_percent=1.0)), ("min_overlap_percent_1.5", dict(min_overlap_percent=1.5))]


def seq_map(fasta: bytes):
    return {r.id: ocli.full_seq(r.seq) for r in ocli.parse_fasta(fasta)}


@pytest.mark.parametrize("d,kw", CASES)
def test_oracle_cli_on_reference_fixtures(d, kw):
    out, table = ocli.cli_monomerize(open(os.path.join(FIX, d, "in.fasta"), "rb").read(), **kw)
    want = open(os.path.join(FIX, d, "out.fasta"), "rb").read()
    assert seq_map(out) == seq_map(want) and table is None


def test_oracle_cli_options():
    x = b"ATGACAGGTACAGCATA"
    fasta = b">dimer desc\n" + x + b"\n" + x + b"\n>plain\nTTTTTTTTTTTTAAAAAAAAAA\n>short\nACG\n"
    out, table = ocli.cli_monomerize(fasta, seed_length=6, table_ext="csv")
    assert out == b">dimer desc\n" + x + b"\n"
    assert table == b"id,original_length,monomer_length\ndimer desc,34,17\n"
    out, table = ocli.cli_monomerize(fasta, seed_length=6, keep_all=True, table_ext="tsv")
    assert out == b">dimer desc\n" + x + b"\n>plain\nTTTTTTTTTTTTAAAAAAAAAA\n>short\nACG\n"
    assert table.splitlines()[0] == b"id\toriginal_length\tmonomer_length" and table.splitlines()[3] == b"short\t3\t3"
    assert ocli.cli_monomerize(fasta, seed_length=6, min_length=18)[0] == b""
    assert ocli.cli_monomerize(fasta, seed_length=6, max_length=16)[0] == b""
    assert ocli.cli_monomerize(b"", table_ext="csv") == (b"", b"")
    with pytest.raises(ValueError, match="both max_mismatch and min_identity"):
        ocli.cli_monomerize(fasta, max_mismatch=1, min_identity=0.9)
    with pytest.raises(ValueError, match="between 0.0 and 1.0"):
        ocli.cli_monomerize(fasta, min_identity=1.5)


def test_product_side_validation_needs_no_gpu():
    """the builder's rules (lib/src/monomerize.rs:20-40) and the CLI's bail!s (src/monomerize.rs:35-45) are checked on the host
    before anything touches the device"""
    from circkit_b200.monomerize import Monomerizer
    from circkit_b200 import cli
    for bad in (0, 64):
        with pytest.raises(ValueError, match="at least 1 and at most 63"):
            Monomerizer(bad)
    with pytest.raises(ValueError, match="overlap_dist and overlap_min_identity"):
        Monomerizer(4, overlap_dist=0, overlap_min_identity=0.9)
    with pytest.raises(ValueError, match="seed_len"):
        Monomerizer()
    with pytest.raises(cli.CliError, match="both max_mismatch and min_identity"):
        cli.monomerize(b">a\nACGT\n", max_mismatch=1, min_identity=0.9)
    with pytest.raises(cli.CliError, match="between 0.0 and 1.0"):
        cli.monomerize(b">a\nACGT\n", min_identity=-0.1)
    assert cli.monomerize(b"") == (b"", None) and cli.monomerize(b"\n\n", table_ext="tsv") == (b"", b"")
