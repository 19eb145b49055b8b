"""GPU parity of the packed host path (CK_F_PACKED_IN): ck_pack2_host + ck_canon_submit_packed / ck_uniq_submit_packed must
give exactly what the ASCII path and the oracle give -- canonical bytes, lengths, (start, strand), XXH3-64, first-occurrence
indices -- for both semantics (CLI: needletail normalisation first, src/canonicalize.rs:24-29; library: bytes as they are,
lib/src/canonicalize.rs:54-63), every symbol lane, empty / tiny / tied records and consecutive batches on both slots."""
import random

import numpy as np
import pytest

import oracle
from circkit_b200 import core

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import circkit_b200
    c = circkit_b200.Context(max_batch_bytes=64 << 20, max_batch_records=1 << 18, table_capacity=1 << 20)
    yield c
    c.close()


def _batch(seqs):
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum([len(s) for s in seqs], out=off[1:])
    arena = np.frombuffer(b"".join(seqs), dtype=np.uint8) if int(off[-1]) else np.zeros(0, dtype=np.uint8)
    return arena, off


def _records(seed, n):
    rng = random.Random(seed)
    alphabets = [b"ACGT"] * 6 + [b"ACGTacgtuU", b"ACGTN-", b"ACGTNRYKMSWBDHV-", b"ACGTacgtNnUuRYx*.~", b"ACGT\n", b"ACGTNn \t\r\n"]
    seqs = [b"", b"A", b"banana", b"A" * 300, b"AC" * 200, b"ACGTNNRYacgu" * 20]
    for _ in range(n):
        alpha = rng.choice(alphabets)
        L = rng.choice([rng.randint(1, 140), rng.randint(120, 600), rng.randint(200, 5000), rng.randint(8000, 20000) if rng.random() < 0.05 else 300])
        seqs.append(bytes(rng.choice(alpha) for _ in range(L)))
    base = [s for s in seqs if len(s) > 150 and set(s) <= set(b"ACGT")]
    for s in base[:60]:                                   # rotated / reverse-complemented duplicates
        r = rng.randrange(len(s))
        t = s[r:] + s[:r]
        seqs.append(oracle.revcomp(t) if rng.random() < 0.5 else t)
    return seqs


@pytest.mark.parametrize("normalize", [True, False], ids=["cli-semantics", "lib-semantics"])
@pytest.mark.parametrize("aligned", [False, True], ids=["same-offsets", "aligned-arena"])
def test_packed_canonicalize_matches_the_oracle(ctx, normalize, aligned):
    seqs = _records(11 + normalize, 1500)
    arena, off = _batch(seqs)
    want = oracle.canonicalize_batch(arena, off, normalize=normalize, threads=8)
    got = ctx.canonicalize_batch_packed(arena, off, normalize=normalize, aligned=aligned, threads=4)
    assert np.array_equal(got["lens"], want["lens"])
    assert np.array_equal(got["start"], want["start"]) and np.array_equal(got["strand"], want["strand"])
    assert np.array_equal(got["hash"], want["hash"])
    gstart = ctx.aligned_starts(off) if aligned else off[:-1]
    for i in range(len(seqs)):
        o, n, g = int(off[i]), int(want["lens"][i]), int(gstart[i])
        assert got["out"][g:g + n].tobytes() == want["out"][o:o + n].tobytes(), (i, seqs[i][:60])


def test_packed_uniq_two_slots_matches_the_serial_consumer(ctx):
    """two consecutive batches through slots 0 and 1 (src/uniq.rs:42-78: one map over the whole input, first record wins)"""
    seqs = _records(5, 2500)
    random.Random(1).shuffle(seqs)
    arena, off = _batch(seqs)
    want = oracle.canonicalize_batch(arena, off, normalize=True, threads=8)
    wh, wfirst = oracle.uniq_consume(want["out"], off, want["lens"])
    cut = len(seqs) // 2
    ctx.uniq_reset()
    parts = []
    for k, (lo, hi) in enumerate([(0, cut), (cut, len(seqs))]):
        o = off[lo: hi + 1] - off[lo]
        a = arena[int(off[lo]): int(off[hi])]
        pb = core.pack2_host(a, o, normalize=True, threads=3)
        ctx.uniq_submit_packed(k, pb, lo, no_bytes=True)
        parts.append((k, pb))
    firsts, hashes = [], []
    for k, pb in parts:
        r = ctx.uniq_wait(k, pb.n, pb.total, want_bytes=False)
        firsts.append(r["first"]); hashes.append(r["hash"])
    assert np.array_equal(np.concatenate(hashes), wh)
    assert np.array_equal(np.concatenate(firsts), wfirst)


def test_packed_and_ascii_paths_agree_on_a_large_batch(ctx):
    rng = np.random.default_rng(3)
    n = 60_000
    lens = rng.integers(200, 700, n)
    off = np.zeros(n + 1, dtype=np.uint64); np.cumsum(lens, out=off[1:])
    arena = rng.choice(np.frombuffer(b"ACGT", np.uint8), int(off[-1])).astype(np.uint8)
    a = ctx.canonicalize_batch(arena, off, normalize=True, aligned=True)
    b = ctx.canonicalize_batch_packed(arena, off, normalize=True, aligned=True)
    for k in ("lens", "start", "strand", "hash"):
        assert np.array_equal(a[k], b[k]), k
    starts = ctx.aligned_starts(off).astype(np.int64)
    idx = np.repeat(starts, lens) + (np.arange(int(off[-1])) - np.repeat(off[:-1].astype(np.int64), lens))
    assert np.array_equal(a["out"][idx], b["out"][idx])


@pytest.mark.parametrize("aligned", [False, True], ids=["same-offsets", "aligned-arena"])
@pytest.mark.parametrize("packed", [False, True], ids=["ascii-in", "packed-in"])
def test_device_compaction_returns_the_survivors_only(ctx, aligned, packed):
    """CK_F_SURVIVORS / ck_uniq_wait_survivors: the first occurrences (src/uniq.rs:47-61), in input order, with their canonical
    bytes -- two batches, so that survivors of the second batch depend on the first."""
    seqs = _records(21, 1800)
    random.Random(2).shuffle(seqs)
    arena, off = _batch(seqs)
    want = oracle.canonicalize_batch(arena, off, normalize=True, threads=8)
    wh, wfirst = oracle.uniq_consume(want["out"], off, want["lens"])
    keep = np.nonzero(wfirst == np.arange(len(seqs), dtype=np.uint64))[0]
    cut = len(seqs) // 3
    ctx.uniq_reset()
    got_idx, got_bytes = [], []
    for k, (lo, hi) in enumerate([(0, cut), (cut, len(seqs))]):
        o = off[lo: hi + 1] - off[lo]
        a = arena[int(off[lo]): int(off[hi])]
        if packed:
            pb = core.pack2_host(a, o, normalize=True, threads=2)
            ctx.uniq_submit_packed(k, pb, lo, aligned=aligned, survivors=True)
        else:
            ctx.uniq_submit(k, a, o, lo, normalize=True, aligned=aligned, survivors=True)
        r = ctx.uniq_wait_survivors(k, hi - lo, int(o[-1]))
        assert np.array_equal(r["first"], wfirst[lo:hi]) and np.array_equal(r["hash"], wh[lo:hi])
        assert np.array_equal(r["lens"], want["lens"][lo:hi])
        assert np.all(np.diff(r["index"].astype(np.int64)) > 0)
        for j, i in enumerate(r["index"]):
            g = lo + int(i)
            n = int(want["lens"][g])
            c0 = int(r["offsets"][j])
            assert c0 % 16 == 0
            got_idx.append(g)
            got_bytes.append(r["bytes"][c0: c0 + n].tobytes())
    assert got_idx == [int(x) for x in keep]
    for g, b in zip(got_idx, got_bytes):
        o0, n = int(off[g]), int(want["lens"][g])
        assert b == want["out"][o0: o0 + n].tobytes(), g


def test_library_drop_ins_from_several_threads(ctx):
    """lib/src/canonicalize.rs:54 is called from T worker threads (src/canonicalize.rs:21-30): the drop-ins must be safe from
    several host threads on one context, and fast enough to be called per record."""
    import threading
    import time
    rng = random.Random(9)
    seqs = [bytes(rng.choice(b"ACGTN") for _ in range(rng.randint(1, 600))) for _ in range(160)] + [b"", b"banana", b"A" * 50]
    want = [(oracle.canonicalize(s), oracle.lmsr(s), oracle.lmsr_index(s)) for s in seqs]
    errors = []

    def work(t):
        try:
            for i in range(t, len(seqs), 4):
                assert ctx.canonicalize(seqs[i]) == want[i][0]
                assert ctx.lmsr(seqs[i]) == want[i][1]
                assert ctx.lmsr_index(seqs[i]) == want[i][2]
        except Exception as e:                              # noqa: BLE001
            errors.append(repr(e))
    th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    [x.start() for x in th]
    [x.join() for x in th]
    assert not errors, errors[:3]
    s = seqs[5]
    ctx.canonicalize(s)
    lat = []
    for _ in range(200):
        t0 = time.perf_counter()
        ctx.canonicalize(s)
        lat.append(time.perf_counter() - t0)
    lat.sort()
    print("single-record ck_canonicalize latency: median %.1f us, p90 %.1f us" % (lat[100] * 1e6, lat[180] * 1e6))
    assert lat[100] < 200e-6
