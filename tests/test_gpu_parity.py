"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the committed goldens.

Bit-exact bar: canonical bytes, (start, strand), XXH3-64 and first-occurrence indices must be
identical.  Mirrors the reference's own tests (lib/src/canonicalize.rs:65-232 KATs + properties,
tests/canon_uniq.rs fixtures) and adds the cases the reference leaves unpinned.
"""
import json
import os
import random

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
KATS = json.load(open(os.path.join(GOLDEN, "kats.json")))


@pytest.fixture(scope="module")
def ctx():
    import circkit_b200
    c = circkit_b200.Context(max_batch_bytes=256 << 20, max_batch_records=1 << 20, table_capacity=1 << 21)
    yield c
    c.close()


def _batch(seqs):
    lens = np.array([len(s) for s in seqs], dtype=np.uint64)
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    arena = np.frombuffer(b"".join(seqs), dtype=np.uint8) if int(off[-1]) else np.zeros(0, dtype=np.uint8)
    return arena, off


def _check_batch(ctx, seqs, normalize, tag="", aligned=None):
    if aligned is None:                       # both output layouts
        _check_batch(ctx, seqs, normalize, tag + " [same offsets]", False)
        _check_batch(ctx, seqs, normalize, tag + " [aligned arena]", True)
        return
    arena, off = _batch(seqs)
    want = oracle.canonicalize_batch(arena, off, normalize=normalize, threads=4)
    got = ctx.canonicalize_batch(arena, off, normalize=normalize, aligned=aligned)
    gstart = ctx.aligned_starts(off) if aligned else off[:-1]
    bad = []
    for i in range(len(seqs)):
        o = int(off[i])
        wl, gl = int(want["lens"][i]), int(got["lens"][i])
        wb = want["out"][o:o + wl].tobytes()
        go = int(gstart[i])
        gb = got["out"][go:go + gl].tobytes()
        if (wl != gl or wb != gb or int(want["start"][i]) != int(got["start"][i])
                or int(want["strand"][i]) != int(got["strand"][i]) or int(want["hash"][i]) != int(got["hash"][i])):
            bad.append((i, seqs[i][:80], wb[:60], gb[:60], int(want["start"][i]), int(got["start"][i]),
                        int(want["strand"][i]), int(got["strand"][i]), hex(int(want["hash"][i])), hex(int(got["hash"][i]))))
            if len(bad) >= 5:
                break
    assert not bad, f"{tag}: {len(bad)}+ mismatches, first: {bad[:3]}"


# ---------------------------------------------------------------- reference KATs through the lib drop-ins
def test_lmsr_index_kats(ctx):
    for k in KATS["lmsr_index"]:
        assert ctx.lmsr_index(k["in"].encode()) == k["out"], k


def test_lmsr_kats(ctx):
    for k in KATS["lmsr"]:
        assert ctx.lmsr(k["in"].encode()) == k["out"].encode(), k
    for s in KATS["lmsr_idempotent"]:
        t = ctx.lmsr(s.encode())
        assert ctx.lmsr(t) == t


def test_canonicalize_kats(ctx):
    for k in KATS["canonicalize"]:
        assert ctx.canonicalize(k["in"].encode()) == k["out"].encode(), k
    for a, b in KATS["same_circle"]:
        assert ctx.lmsr(a.encode()) == ctx.lmsr(b.encode())
        assert ctx.canonicalize(a.encode()) == ctx.canonicalize(b.encode()) == b"AAACGCTGCTAAATCAATTTCCTCCATCACCTAGTTTATGTAG"


def test_module_level_drop_ins():
    import circkit_b200
    assert circkit_b200.lmsr_index(b"banana") == 5
    assert circkit_b200.lmsr(b"TAA") == b"AAT"
    assert circkit_b200.canonicalize(b"ATT") == b"AAT"
    assert circkit_b200.canonicalize(b"") == b""
    assert circkit_b200.lmsr_index(b"") == 0


# ---------------------------------------------------------------- proptest-style properties (:216-231)
def test_lmsr_index_matches_three_way_on_printable(ctx):
    rng = random.Random(1)
    seqs = [bytes(rng.randrange(0x20, 0x7F) for _ in range(rng.randint(1, 100))) for _ in range(3000)]
    # low-entropy strings exercise ties
    seqs += [bytes(rng.choice(b"ab") for _ in range(rng.randint(1, 100))) for _ in range(2000)]
    arena, off = _batch(seqs)
    got = ctx.lmsr_index_batch(arena, off)
    for i, s in enumerate(seqs):
        w = oracle.lmsr_index(s)
        assert int(got[i]) == w, (s, w, int(got[i]))
        if i < 300:
            assert w == oracle.lmsr_index_2(s) == oracle.lmsr_index_simple(s)


def test_canonicalize_idempotent_on_device(ctx):
    rng = random.Random(2)
    seqs = [bytes(rng.choice(b"ATGC") for _ in range(rng.randint(1, 100))) for _ in range(2000)]
    arena, off = _batch(seqs)
    r1 = ctx.canonicalize_batch(arena, off)
    r2 = ctx.canonicalize_batch(r1["out"], off)
    assert np.array_equal(r1["out"], r2["out"])
    assert np.array_equal(r1["hash"], r2["hash"])


# ---------------------------------------------------------------- lanes x sizes x shapes
def _rand_dna(rng, n, alpha=b"ACGT"):
    return bytes(rng.choice(alpha) for _ in range(n))


def _with_structure(rng, n, alpha):
    """periodic, multimer, palindromic and low-complexity circles: every tie path."""
    kind = rng.randrange(7)
    if kind == 0:
        return bytes([rng.choice(alpha)]) * n                                   # homopolymer
    if kind == 1:
        p = rng.randint(1, max(1, min(40, n)))
        u = _rand_dna(rng, p, alpha)
        return (u * (n // p + 1))[:max(1, (n // p) * p)]                        # exact power of a short word
    if kind == 2:
        h = _rand_dna(rng, max(1, n // 2), alpha)
        return h + h                                                             # dimer
    if kind == 3:
        h = _rand_dna(rng, max(1, n // 2), alpha)
        return h + oracle.revcomp(h)                                             # reverse-palindromic circle
    if kind == 4:
        u = _rand_dna(rng, rng.randint(1, 5), alpha)
        s = bytearray((u * (n // len(u) + 1))[:n])
        for _ in range(rng.randint(0, 2)):
            s[rng.randrange(n)] = rng.choice(alpha)                              # near-periodic
        return bytes(s)
    if kind == 5:
        s = bytearray(_rand_dna(rng, n, alpha))
        a = rng.randrange(n); L = rng.randint(1, n)
        for k in range(L):
            s[(a + k) % n] = alpha[0]                                            # long run of the smallest symbol
        return bytes(s)
    r = _rand_dna(rng, max(1, n // 3), alpha)
    return r * 3                                                                 # trimer


@pytest.mark.parametrize("alpha,normalize", [
    (b"ACGT", False),                      # 2-bit lane
    (b"ACGTN-", True),                     # 4-bit lane, CLI alphabet
    (b"ACGTNRYKMSWBDHV-", False),          # 4-bit lane, library semantics (bio complement table)
    (b"ACGTacgtNnUuRYx*.", False),         # byte lane, library semantics
    (b"ACGTacgtNnUuRYKMSWBDHVryx*.~ \t\r\n", True),   # CLI normalisation of everything
])
def test_random_and_structured_records_all_lanes(ctx, alpha, normalize):
    rng = random.Random(hash(alpha) & 0xffff)
    seqs = []
    for n in list(range(1, 70)) + [127, 128, 129, 239, 240, 241, 255, 256, 257, 325, 400, 511, 512, 513, 700,
                                   1023, 1024, 1025, 1100, 2047, 2048, 2049, 3000]:
        seqs.append(_rand_dna(rng, n, alpha))
        seqs.append(_with_structure(rng, n, alpha))
    for _ in range(1500):
        n = rng.randint(1, 600)
        seqs.append(_rand_dna(rng, n, alpha) if rng.random() < 0.5 else _with_structure(rng, n, alpha))
    seqs.append(b"")
    _check_batch(ctx, seqs, normalize, tag=repr(alpha))


@pytest.mark.parametrize("normalize", [True, False], ids=["cli-semantics", "lib-semantics"])
def test_config3_shape_iupac_records(ctx, normalize):
    """BASELINE config 3 (mixed IUPAC / N records, byte-level lanes) at 20 k records: CLI semantics (needletail
    normalisation: IUPAC -> N, lowercase -> upper, U -> T; 4-bit lane) and library semantics (bytes as they are, bio's
    complement table; 4-bit lane for upper-case IUPAC, byte lane for soft-masked / RNA records)."""
    from circkit_b200 import synth_host
    arena, off = synth_host.make_iupac_records(20000, 250, 400, seed=3)
    seqs = [arena[int(off[i]): int(off[i + 1])].tobytes() for i in range(len(off) - 1)]
    _check_batch(ctx, seqs, normalize, tag="config 3 " + ("cli" if normalize else "lib"))


def test_long_records_cta_shape(ctx):
    rng = random.Random(9)
    seqs = []
    for n in [8191, 8192, 8193, 20000, 65536, 65537, 100003, 200000]:
        seqs.append(_rand_dna(rng, n))
    for n in [9000, 30000, 70000]:
        for _ in range(3):
            seqs.append(_with_structure(rng, n, b"ACGT"))
    seqs.append(_rand_dna(rng, 5000, b"ACGTN"))        # 4-bit CTA shape
    seqs.append(_with_structure(rng, 6000, b"ACGTN"))
    seqs.append(_rand_dna(rng, 3000, b"ACGTacgtn"))    # byte CTA shape
    seqs.append(_with_structure(rng, 2500, b"ACgt"))
    _check_batch(ctx, seqs, False, tag="long")


def test_records_beyond_the_staged_length_classes(ctx):
    """The reference has no length limit (lib/src/canonicalize.rs:5-36).  Records beyond the classes whose strands fit shared
    memory (2-bit > 425 984, 4-bit > 212 992, bytes > 106 496 symbols) are staged in global memory (k_canon_cta<BITS, true>):
    host-buffer batches (ASCII and host-packed), uniq, and the library drop-ins, every symbol lane, with ties."""
    rng = np.random.default_rng(77)

    def rnd(n, alpha):
        return bytes(rng.choice(np.frombuffer(alpha, np.uint8), n).astype(np.uint8))
    u = rnd(1013, b"ACGT")
    seqs = [rnd(425985, b"ACGT"), rnd(700001, b"ACGT"), rnd(250000, b"ACGTN"), rnd(120000, b"ACGTacgtn"),
            (u * 600)[:600000],                       # a cut power: every 1013th rotation ties on the key
            rnd(300000, b"ACGT") * 2,                 # an exact dimer
            rnd(1000, b"ACGT"), rnd(300, b"ACGTN"), b"", rnd(425984, b"ACGT")]
    _check_batch(ctx, seqs, False, tag="huge")
    arena, off = _batch(seqs)
    want = oracle.canonicalize_batch(arena, off, normalize=False, threads=8)
    got = ctx.canonicalize_batch_packed(arena, off, normalize=False)
    assert np.array_equal(got["hash"], want["hash"]) and np.array_equal(got["start"], want["start"])
    ctx.uniq_reset()
    first = ctx.uniq_batch(np.concatenate([arena, arena]), np.concatenate([off, off[1:] + off[-1]]), want_bytes=False)["first"]
    n = len(seqs)
    assert [int(x) for x in first[n:]] == [int(x) for x in first[:n]]                # the second copies point at the first ones
    assert sorted(set(int(x) for x in first[:n])) == [i for i in range(n) if int(first[i]) == i]
    assert ctx.canonicalize(seqs[1]) == oracle.canonicalize(seqs[1])
    assert ctx.lmsr_index(seqs[4]) == oracle.lmsr_index(seqs[4])
    assert ctx.canonicalize(seqs[2]) == oracle.canonicalize(seqs[2])


def test_segment_kernel_lengths_and_rotations(ctx):
    """k_canon_seg (one warp per record, a lane per segment): lengths around its oct / block / class boundaries, every record
    also as a rotated and as a reverse-complemented copy (the same canonical form and hash must come back), and records whose
    minimal 16-mer occurs twice (they take the CTA kernel's duel path)."""
    rng = random.Random(12)
    seqs = []
    for n in [8193, 8200, 8224, 8320, 9000, 12345, 16384, 16385, 20480, 32768, 33000, 65535, 65536, 65537, 65600, 100001, 131072, 150000]:
        s = _rand_dna(rng, n)
        r = rng.randrange(n)
        seqs += [s, s[r:] + s[:r], oracle.revcomp(s[r:] + s[:r])]
    core = _rand_dna(rng, 20)
    for n in [9000, 40000]:                                     # the same 20-mer planted twice, and a planted poly-A run
        s = bytearray(_rand_dna(rng, n, b"CGT"))
        s[100:120] = b"A" * 20; s[5000:5020] = b"A" * 20
        seqs.append(bytes(s))
        s2 = bytearray(_rand_dna(rng, n, b"CGT"))
        s2[n - 10: n] = b"A" * 10; s2[0:15] = b"A" * 15            # the minimal rotation starts in the record's last bases
        seqs.append(bytes(s2))
    _check_batch(ctx, seqs, False, tag="segments", aligned=True)
    arena, off = _batch(seqs)
    got = ctx.canonicalize_batch(arena, off, normalize=False, aligned=True)
    for i in range(0, 54, 3):
        assert int(got["hash"][i]) == int(got["hash"][i + 1]) == int(got["hash"][i + 2]), i


def test_adversarial_long_periodic(ctx):
    rng = random.Random(10)
    seqs = [b"A" * 50000, b"AC" * 40000, _rand_dna(rng, 171) * 300, b"A" * 2000 + _rand_dna(rng, 30000) + b"A" * 1500,
            b"T" * 9001, (b"ACGT" * 5000)]
    _check_batch(ctx, seqs, False, tag="adversarial")


def test_xxh3_every_length_class(ctx):
    v = json.load(open(os.path.join(GOLDEN, "xxh3_vectors.json")))
    # canonical form of an already-canonical sequence is itself, so hashes of goldens can be checked
    # through the device: feed oracle-canonicalised DNA of each length and compare to python-xxhash goldens
    seqs, want = [], []
    for e in v["vectors"]:
        L = e["len"]
        if e["kind"] != "dna":
            continue
        dna = bytes(b"ACGT"[(i * 7 + (i >> 3) * 3 + L) & 3] for i in range(L))
        seqs.append(dna)
        want.append(oracle.xxh3_64(oracle.canonicalize(dna)))
    arena, off = _batch(seqs)
    got = ctx.canonicalize_batch(arena, off)
    for i, s in enumerate(seqs):
        assert int(got["hash"][i]) == want[i], (len(s), hex(want[i]), hex(int(got["hash"][i])))
    # and the goldens themselves, where the input happens to be its own canonical form
    for e, s in zip([e for e in v["vectors"] if e["kind"] == "dna"], seqs):
        if oracle.canonicalize(s) == s:
            i = seqs.index(s)
            assert "%016x" % int(got["hash"][i]) == e["xxh3_64"]


def test_real_monomer_corpus(ctx):
    from oracle import cli
    data = open(os.path.join(GOLDEN, "fixtures", "real_monomers", "in.fasta"), "rb").read()
    recs = cli.parse_fasta(data)
    _check_batch(ctx, [r.seq for r in recs], True, tag="real monomers")


# ---------------------------------------------------------------- uniq
def test_uniq_first_occurrence_matches_serial_consumer(ctx):
    rng = random.Random(12)
    base = [_rand_dna(rng, rng.randint(200, 500)) for _ in range(3000)]
    seqs = []
    for i in range(9000):
        if i < 50 or rng.random() < 0.6:
            seqs.append(rng.choice(base) if rng.random() < 0.2 else _rand_dna(rng, rng.randint(200, 500)))
        else:
            s = rng.choice(seqs)
            r = rng.randrange(len(s))
            s = s[r:] + s[:r]
            seqs.append(oracle.revcomp(s) if rng.random() < 0.5 else s)
    arena, off = _batch(seqs)
    want = oracle.canonicalize_batch(arena, off, normalize=True, threads=4)
    wh, wfirst = oracle.uniq_consume(want["out"], off, want["lens"])
    ctx.uniq_reset()
    # two batches through both slots, to exercise the cross-batch table and the ordering event
    cut = 4000
    a0, o0 = arena[: int(off[cut])], off[: cut + 1].copy()
    a1, o1 = arena[int(off[cut]):], (off[cut:] - off[cut]).copy()
    ctx.uniq_submit(0, np.ascontiguousarray(a0), o0, 0, normalize=True)
    ctx.uniq_submit(1, np.ascontiguousarray(a1), o1, cut, normalize=True)
    r0 = ctx.uniq_wait(0, cut, int(o0[-1]))
    r1 = ctx.uniq_wait(1, len(seqs) - cut, int(o1[-1]))
    gfirst = np.concatenate([r0["first"], r1["first"]])
    ghash = np.concatenate([r0["hash"], r1["hash"]])
    assert np.array_equal(ghash, wh)
    assert np.array_equal(gfirst, wfirst)
    assert (gfirst == np.arange(len(seqs))).sum() < len(seqs)      # there were duplicates
    ctx.uniq_reset()


def test_repeated_fixture_collapses(ctx):
    from oracle import cli
    data = open(os.path.join(GOLDEN, "fixtures", "repeated", "in.fasta"), "rb").read()
    recs = cli.parse_fasta(data)
    arena, off = _batch([r.seq for r in recs])
    ctx.uniq_reset()
    r = ctx.uniq_batch(arena, off, 0, normalize=True)
    assert list(r["first"]) == [0, 0, 0, 0, 0]
    assert r["out"][: int(r["lens"][0])].tobytes() == b"AAAAAAAT"
    ctx.uniq_reset()


def test_errors_are_codes_not_crashes(ctx):
    import circkit_b200
    arena, off = _batch([b"ACGT"])
    with pytest.raises(circkit_b200.CircKitError):
        ctx.canon_wait(1, 1, 4)                     # wait without submit
    bad = off.copy(); bad[0] = 1
    with pytest.raises(circkit_b200.CircKitError):
        ctx.canon_submit(0, arena, bad, normalize=False)
