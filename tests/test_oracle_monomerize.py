"""The monomerize oracle (oracle/monomerize.py) against the reference's own unit tests (lib/src/monomerize.rs:155-554,
transcribed by tests/golden/make_monomerize_kats.py) and its two property tests (:514-551)."""
import json
import os
import random

import pytest

from oracle.monomerize import Monomerizer

HERE = os.path.dirname(os.path.abspath(__file__))


def load_kats():
    d = json.load(open(os.path.join(HERE, "golden", "monomerize_kats.json")))
    out = []
    for k in d["kats"]:
        k = dict(k)
        k["seq"], k["expected"] = d["strings"][k["seq"]].encode(), d["strings"][k["expected"]].encode()
        out.append(k)
    return out


def test_reference_unit_tests():
    kats = load_kats()
    assert len(kats) == 277
    for k in kats:
        m = Monomerizer(k["seed_len"], k["overlap_dist"], k["min_identity"])
        got = m.monomerize_sensitive(k["seq"]) if k["sensitive"] else m.monomerize(k["seq"])
        assert got == k["expected"], (k["name"], k["ref"], k["seed_len"], k["overlap_dist"], k["min_identity"])


def test_compiled_restatement_on_reference_unit_tests_and_against_the_python_one():
    """oracle/ck_oracle.c: ck_o_monomer_end (the CPU baseline of tools/time_monomerize.py)"""
    import numpy as np
    from oracle.monomerize import c_end_index, c_end_indices_batch
    for k in load_kats():
        e = c_end_index(k["seq"], k["seed_len"], k["overlap_dist"], k["min_identity"], sensitive=k["sensitive"])
        assert (k["seq"] if e is None else k["seq"][:e]) == k["expected"], (k["name"], k["ref"])
    rng = random.Random(9)
    seqs = [b"", b"A", b"ACGT" * 5]
    for _ in range(800):
        alpha = rng.choice([b"ACGT", b"AC", b"ACGTNRYacgtn-"])
        unit = bytes(rng.choice(alpha) for _ in range(rng.randint(1, 90)))
        s = bytearray(unit * rng.randint(1, 3) + unit[: rng.randint(0, len(unit))])
        for _ in range(rng.randint(0, 3)):
            if s:
                s[rng.randrange(len(s))] = rng.choice(alpha)
        seqs.append(bytes(s))
    for seed_len, dist, ident in ((4, 0, None), (7, 2, None), (10, None, 0.95), (1, 0, None)):
        m = Monomerizer(seed_len, dist, ident)
        for sens in (False, True):
            want = [m.last_monomer_end_index_sensitive(s) if sens else m.last_monomer_end_index(s) for s in seqs]
            off = np.zeros(len(seqs) + 1, dtype=np.uint64); np.cumsum([len(s) for s in seqs], out=off[1:])
            got = c_end_indices_batch(np.frombuffer(b"".join(seqs), dtype=np.uint8), off, seed_len, dist, ident, sens, threads=3)
            assert [None if g == 0xffffffff else int(g) for g in got] == want
        assert [c_end_index(s, seed_len, dist, ident, first_only=True) for s in seqs[:200]] == \
               [m.first_monomer_end_index(s) for s in seqs[:200]]


def test_builder_validation():
    """lib/src/monomerize.rs:20-40 and the `validation` tests (:394-419), overlap_percentage_and_dist_panics (:479-492)"""
    for bad in (0, 64, 100):
        with pytest.raises(ValueError, match="at least 1 and at most 63"):
            Monomerizer(bad)
    with pytest.raises(ValueError, match="overlap_dist and overlap_min_identity"):
        Monomerizer(4, overlap_dist=1, overlap_min_identity=0.95)
    with pytest.raises(ValueError, match="seed_len"):
        Monomerizer(None)


def test_concatenated_always_monomerizes():
    """proptest concatenated_always_monomerizes (:518-528): XXX -> X"""
    rng = random.Random(1)
    m = Monomerizer(10, overlap_min_identity=0.95)
    for _ in range(300):
        x = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(12, 100)))
        assert m.monomerize(x * 3) == x


def test_small_mutations_outside_seed_still_monomerize():
    """proptest small_mutations_outside_seed_still_monomerize (:530-550)"""
    rng = random.Random(2)
    m = Monomerizer(10, overlap_min_identity=0.95)
    nxt = {65: 67, 67: 71, 71: 84, 84: 65}
    for _ in range(300):
        x = bytes(rng.choice(b"ACGT") for _ in range(rng.randint(100, 200)))
        c = bytearray(x + x)
        i = rng.randrange(10, 90)
        c[i] = nxt[c[i]]
        assert m.monomerize(bytes(c)) == bytes(c[: len(x)])


def test_short_and_degenerate_inputs():
    m = Monomerizer(4, overlap_dist=0)
    assert m.first_monomer_end_index(b"") is None and m.first_monomer_end_index(b"ACGT") is None   # len <= seed_len (:53-56)
    assert m.monomerize(b"ACGTA") == b"ACGTA"
    assert m.monomerize(b"AAAAAAAAAAAA") == b"A" * 4      # homopolymer: shrinks until len <= seed_len stops it
    assert Monomerizer(1, overlap_dist=0).monomerize(b"AAAAAAAAAAAA") == b"A"
