"""GPU parity of `monomerize` (csrc/ck_monomerize.cuh through ck_dev_monomerize) with the reference's own unit tests
(lib/src/monomerize.rs:155-554, tests/golden/monomerize_kats.json), its property tests and the oracle on random batches."""
import json
import os
import random

import numpy as np
import pytest

from oracle.monomerize import Monomerizer as OracleMonomerizer

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ctx():
    import circkit_b200
    c = circkit_b200.Context(max_batch_bytes=0, max_batch_records=0)
    yield c
    c.close()


def _kats():
    d = json.load(open(os.path.join(HERE, "golden", "monomerize_kats.json")))
    return [dict(k, seq=d["strings"][k["seq"]].encode(), expected=d["strings"][k["expected"]].encode()) for k in d["kats"]]


def test_reference_unit_tests_on_device(ctx):
    from circkit_b200.monomerize import Monomerizer
    kats = _kats()
    assert len(kats) == 277
    groups = {}
    for k in kats:                                            # one launch per parameter set
        groups.setdefault((k["seed_len"], k["overlap_dist"], k["min_identity"], k["sensitive"]), []).append(k)
    for (s, d, ident, sens), ks in groups.items():
        m = Monomerizer(s, d, ident, ctx=ctx)
        ends = m.end_indices([k["seq"] for k in ks], sensitive=sens)
        for k, e in zip(ks, ends):
            got = k["seq"] if e is None else k["seq"][:e]
            assert got == k["expected"], (k["name"], k["ref"], s, d, ident)
    # the per-sequence drop-ins
    m = Monomerizer(4, overlap_min_identity=0.95, ctx=ctx)
    assert m.monomerize_sensitive(b"ATGCCCATGCGCCAGCGCAAATGCCCATGCGCCAGCGCAG") == b"ATGCCCATGCGCCAGCGCAA"
    assert Monomerizer(4, overlap_dist=0, ctx=ctx).monomerize(b"TGCCAATGCATGCCAATGC") == b"TGCCAATGCA"
    assert Monomerizer(4, overlap_dist=0, ctx=ctx).first_monomer_end_index(b"ACGT") is None


def test_builder_validation_matches_reference(ctx):
    from circkit_b200.monomerize import Monomerizer
    import circkit_b200
    for bad in (0, 64, 100):
        with pytest.raises(ValueError, match="at least 1 and at most 63"):
            Monomerizer(bad)
    with pytest.raises(ValueError, match="overlap_dist and overlap_min_identity"):
        Monomerizer(4, overlap_dist=1, overlap_min_identity=0.95)
    with pytest.raises(ValueError, match="seed_len"):
        Monomerizer()
    # the C ABI enforces the same rules
    import torch
    z = torch.zeros(2, dtype=torch.int64, device="cuda"); o = torch.zeros(1, dtype=torch.int32, device="cuda")
    for seed, dist, ident in ((0, 0, -1.0), (64, 0, -1.0), (4, 1, 0.95)):
        rc = ctx._lib.ck_dev_monomerize(ctx.handle, None, None, z.data_ptr(), None, 1, seed, dist, ident, 0, o.data_ptr())
        assert rc == -2                                       # CK_ERR_ARG


def test_property_tests_on_device(ctx):
    """proptests of lib/src/monomerize.rs:514-551 as two batches"""
    from circkit_b200.monomerize import Monomerizer
    rng = random.Random(5)
    m = Monomerizer(10, overlap_min_identity=0.95, ctx=ctx)
    xs = [bytes(rng.choice(b"ACGT") for _ in range(rng.randint(12, 100))) for _ in range(2000)]
    assert m.end_indices([x * 3 for x in xs]) == [len(x) for x in xs]
    nxt = {65: 67, 67: 71, 71: 84, 84: 65}
    seqs, want = [], []
    for _ in range(2000):
        x = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(100, 200)))
        c = bytearray(x + x)
        i = rng.randrange(10, 90)
        c[i] = nxt[c[i]]
        seqs.append(bytes(c)); want.append(len(x))
    assert m.end_indices(seqs) == want


@pytest.mark.parametrize("seed_len,dist,ident", [(4, 0, None), (6, 2, None), (10, None, 0.95), (1, 1, None), (63, 3, None), (12, None, 0.80)])
@pytest.mark.parametrize("sensitive", [False, True])
def test_random_batches_match_oracle(ctx, seed_len, dist, ident, sensitive):
    """concatemers with mutations, partial copies, reverse-complement palindromes, low-complexity and IUPAC / lower-case
    bytes (library semantics: bytes as they are, bio's complement table in the sensitive pass), empty and short records"""
    from circkit_b200.monomerize import Monomerizer
    import oracle
    rng = random.Random(seed_len * 1000 + (dist or 0) * 10 + int(sensitive))
    seqs = [b"", b"A", b"ACGT" * 3, b"A" * 200, b"AC" * 150]
    for _ in range(1500):
        alpha = rng.choice([b"ACGT", b"ACGT", b"AC", b"ACGTNRYacgtn-", b"A"])
        unit = bytes(rng.choice(alpha) for _ in range(rng.randint(1, 150)))
        kind = rng.randrange(6)
        if kind == 0:
            s = unit * rng.randint(1, 4)
        elif kind == 1:
            s = unit * rng.randint(2, 3) + unit[: rng.randint(0, len(unit))]
        elif kind == 2:
            s = bytearray(unit * rng.randint(2, 4))
            for _ in range(rng.randint(0, 4)):
                s[rng.randrange(len(s))] = rng.choice(alpha)
            s = bytes(s)
        elif kind == 3:
            s = unit + oracle.revcomp(unit) + unit[: rng.randint(0, len(unit))]
        elif kind == 4:
            s = bytes(rng.choice(alpha) for _ in range(rng.randint(0, 400)))
        else:
            s = unit[: rng.randint(0, len(unit))] + unit * 2
        seqs.append(s)
    om = OracleMonomerizer(seed_len, dist, ident)
    want = [om.last_monomer_end_index_sensitive(s) if sensitive else om.last_monomer_end_index(s) for s in seqs]
    m = Monomerizer(seed_len, dist, ident, ctx=ctx)
    got = m.end_indices(seqs, sensitive=sensitive)
    bad = [i for i in range(len(seqs)) if got[i] != want[i]]
    assert not bad, (bad[:5], [(seqs[i][:80], got[i], want[i]) for i in bad[:3]])
    first_want = [om.first_monomer_end_index(s) for s in seqs[:300]]
    assert m.end_indices(seqs[:300], first_only=True) == first_want
