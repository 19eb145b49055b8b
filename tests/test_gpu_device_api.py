"""GPU parity of the device-resident path bench.py times: synthetic packed batches (ck_synth_*),
ck_dev_canon_packed2 with and without a class promise, the device first-occurrence table -- against
the oracle on the unpacked ASCII of the same records."""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    import circkit_b200
    from circkit_b200 import device as D
    ctx = circkit_b200.Context(max_batch_bytes=0, max_batch_records=0)
    yield ctx, D, torch
    ctx.close()


def _oracle_for(ctx, D, torch, batch, n):
    ascii_ = D.unpack_ascii(ctx, batch, n).cpu().numpy()
    off = batch.offsets[: n + 1].cpu().numpy().astype(np.uint64)
    want = oracle.canonicalize_batch(ascii_, off, normalize=False, threads=8)
    return ascii_, off, want


@pytest.mark.parametrize("name,kind,lo,hi,dup,adv,mask,n", [
    ("c1-like direct mode", 0, 250, 400, 0, 0, 1, 30000),
    ("c2-like four classes", 1, 200, 5000, 300, 0, 1 | 2 | 1 << 10 | 1 << 11, 12000),
    ("c5-like dups", 0, 250, 400, 300, 0, 1, 30000),
    ("c4-like long + adversarial", 1, 5000, 200000, 0, 100, 1 << 11 | 4 | 8, 300),
    ("all classes, no promise", 1, 64, 20000, 200, 20, 0, 3000),
    ("short and tiny", 0, 1, 130, 100, 0, 0, 20000),
])
@pytest.mark.parametrize("aligned", [False, True])
def test_device_path_matches_oracle(env, name, kind, lo, hi, dup, adv, mask, n, aligned):
    ctx, D, torch = env
    b = D.synth_batch(ctx, seed=11, first_index=0, n_records=n, kind=kind, lo=lo, hi=hi, dup_permille=dup,
                      adversarial_permille=adv)
    outs = D.CanonOutputs(n, b.total, b.offsets.device, want_bytes=True, want_hash=True, aligned=aligned)
    ws = D.Workspace(ctx, n)
    D.canon_packed2(ctx, b, outs, ws, class_mask=mask)
    D.check(ctx, ws)
    ascii_, off, want = _oracle_for(ctx, D, torch, b, n)
    if aligned:
        # record i lives at 32 * ((offsets[i] >> 5) + i): gather it back to the compact layout
        arena = outs.out.cpu().numpy()
        starts = ctx.aligned_starts(off).astype(np.int64)
        lens = (off[1:] - off[:-1]).astype(np.int64)
        src = np.repeat(starts - off[:-1].astype(np.int64), lens) + np.arange(int(off[-1]), dtype=np.int64)
        got_out = arena[src]
    else:
        got_out = outs.out[: b.total].cpu().numpy()
    assert np.array_equal(got_out, want["out"]), name
    assert np.array_equal(outs.start[:n].cpu().numpy().astype(np.uint32), want["start"]), name
    assert np.array_equal(outs.strand[:n].cpu().numpy(), want["strand"]), name
    assert np.array_equal(outs.hash[:n].cpu().numpy().astype(np.uint64), want["hash"]), name
    # uniq on device vs the serial consumer
    table = D.DeviceTable(ctx, n)
    slots = torch.empty(n, dtype=torch.int64, device=b.offsets.device)
    first = torch.empty(n, dtype=torch.int64, device=b.offsets.device)
    table.insert(outs.hash, n, slots)
    table.first(slots, n, first)
    _, wfirst = oracle.uniq_consume(want["out"], off, want["lens"])
    assert np.array_equal(first.cpu().numpy().astype(np.uint64), wfirst), name
    if dup:
        assert (wfirst != np.arange(n)).sum() > 0.5 * n * dup / 1000     # duplicates were really injected and found


@pytest.mark.parametrize("want_bytes,want_hash", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("lo,hi,mask_extra", [(250, 400, 0), (100, 700, 0), (3000, 8192, 0), (250, 400, 1 << 1)])
def test_lane_kernel_variants(env, want_bytes, want_hash, lo, hi, mask_extra):
    """Every compile-time variant of the lane-per-record kernel (hash? bytes? work list?), in direct mode (one
    class promised), through the sorted work list (several classes), with records it must leave to the retry
    kernels (n < 128, n <= 240 with a hash) and with XXH3 block scrambles (n > 1024)."""
    ctx, D, torch = env
    n = 6000 if hi <= 1000 else 1500
    b = D.synth_batch(ctx, seed=23, first_index=0, n_records=n, kind=0, lo=lo, hi=hi, dup_permille=100)
    outs = D.CanonOutputs(n, b.total, b.offsets.device, want_bytes=want_bytes, want_hash=want_hash, aligned=True)
    ws = D.Workspace(ctx, n)
    D.canon_packed2(ctx, b, outs, ws, class_mask=D.class_mask_for(lo, hi) | mask_extra)
    D.check(ctx, ws)
    ascii_, off, want = _oracle_for(ctx, D, torch, b, n)
    assert np.array_equal(outs.start[:n].cpu().numpy().astype(np.uint32), want["start"])
    assert np.array_equal(outs.strand[:n].cpu().numpy(), want["strand"])
    if want_hash:
        assert np.array_equal(outs.hash[:n].cpu().numpy().astype(np.uint64), want["hash"])
    if want_bytes:
        arena = outs.out.cpu().numpy()
        starts = ctx.aligned_starts(off).astype(np.int64)
        lens = (off[1:] - off[:-1]).astype(np.int64)
        src = np.repeat(starts - off[:-1].astype(np.int64), lens) + np.arange(int(off[-1]), dtype=np.int64)
        assert np.array_equal(arena[src], want["out"])


def test_packed_arena_holds_every_record_doubled(env):
    """Format of the packed arena (ck_device.cuh): record i at 32-byte granule (offsets[i] >> 6) + 3 i, units
    0 .. max((2 n + 143) >> 4, (n >> 4) + 5) hold S[b mod n] (the record twice, then some)."""
    ctx, D, torch = env
    n = 300
    b = D.synth_batch(ctx, seed=5, first_index=0, n_records=n, kind=0, lo=1, hi=700)
    ascii_ = D.unpack_ascii(ctx, b, n).cpu().numpy()
    off = b.offsets.cpu().numpy().astype(np.int64)
    words = b.packed2.cpu().numpy().view(np.uint32)
    code = {65: 0, 67: 1, 71: 2, 84: 3}
    for i in list(range(40)) + [n - 1]:
        L = int(off[i + 1] - off[i])
        g = (int(off[i]) >> 6) + 3 * i
        units = words[8 * g: 8 * g + max((2 * L + 143) >> 4, (L >> 4) + 5) + 1]
        seq = [code[int(c)] for c in ascii_[off[i]: off[i + 1]]]
        for j, u in enumerate(units):
            want = 0
            for k in range(16):
                want = (want << 2) | seq[(16 * j + k) % L]
            assert int(u) == want, (i, L, j)


def test_packed_arena_single_copy_layout(env):
    """CK_F_SINGLE_COPY (batches of records <= 512 bases): record i at 16-byte granule (offsets[i] >> 6) + 2 i, units
    0 .. (n >> 4) + 4 hold S[b mod n]; the same records give the same results in either layout."""
    ctx, D, torch = env
    n = 400
    b = D.synth_batch(ctx, seed=5, first_index=0, n_records=n, kind=0, lo=1, hi=512, dup_permille=300)
    assert b.single
    ascii_ = D.unpack_ascii(ctx, b, n).cpu().numpy()
    off = b.offsets.cpu().numpy().astype(np.int64)
    words = b.packed2.cpu().numpy().view(np.uint32)
    code = {65: 0, 67: 1, 71: 2, 84: 3}
    for i in list(range(40)) + [n - 1]:
        L = int(off[i + 1] - off[i])
        g = (int(off[i]) >> 6) + 2 * i
        units = words[4 * g: 4 * g + (L >> 4) + 5]
        seq = [code[int(c)] for c in ascii_[off[i]: off[i + 1]]]
        for j, u in enumerate(units):
            want = 0
            for k in range(16):
                want = (want << 2) | seq[(16 * j + k) % L]
            assert int(u) == want, (i, L, j)
    b2 = D.synth_batch(ctx, seed=5, first_index=0, n_records=n, kind=0, lo=1, hi=512, dup_permille=300, single=False)
    assert not b2.single and torch.equal(b.offsets, b2.offsets)
    res = []
    for bb in (b, b2):
        outs = D.CanonOutputs(n, bb.total, bb.offsets.device, want_bytes=True, want_hash=True, aligned=True)
        outs.out.zero_()
        ws = D.Workspace(ctx, n)
        D.canon_packed2(ctx, bb, outs, ws)
        D.check(ctx, ws)
        res.append(outs)
    assert torch.equal(res[0].hash, res[1].hash) and torch.equal(res[0].start, res[1].start) and torch.equal(res[0].strand, res[1].strand)


def test_class_promise_violation_is_reported(env):
    import circkit_b200
    ctx, D, torch = env
    b = D.synth_batch(ctx, seed=3, first_index=0, n_records=1000, kind=0, lo=400, hi=600, dup_permille=0)
    outs = D.CanonOutputs(1000, b.total, b.offsets.device)
    ws = D.Workspace(ctx, 1000)
    D.canon_packed2(ctx, b, outs, ws, class_mask=1)           # promise n <= 512 is false
    with pytest.raises(circkit_b200.CircKitError) as e:
        D.check(ctx, ws)
    assert e.value.code == -4


def test_shards_regenerate_identically(env):
    ctx, D, torch = env
    whole = D.synth_batch(ctx, seed=5, first_index=0, n_records=4000, kind=0, lo=250, hi=400, dup_permille=300)
    part = D.synth_batch(ctx, seed=5, first_index=1000, n_records=2000, kind=0, lo=250, hi=400, dup_permille=300)
    a = D.unpack_ascii(ctx, whole).cpu().numpy()
    off = whole.offsets.cpu().numpy()
    p = D.unpack_ascii(ctx, part).cpu().numpy()
    assert np.array_equal(a[off[1000]: off[3000]], p)


def test_full_size_properties_config1(env):
    """BASELINE config 1 at full size (1M records): size-independent properties instead of the oracle:
    canonical form is idempotent, invariant under rotation + reverse complement (duplicates collapse)."""
    ctx, D, torch = env
    n = 1_000_000
    b = D.synth_batch(ctx, seed=1, first_index=0, n_records=n, kind=0, lo=250, hi=400, dup_permille=300)
    outs = D.CanonOutputs(n, b.total, b.offsets.device, want_bytes=True, want_hash=True)
    ws = D.Workspace(ctx, n)
    D.canon_packed2(ctx, b, outs, ws, class_mask=1)
    D.check(ctx, ws)
    # feed the canonical bytes back through the byte entry point: must be a fixed point with equal hashes
    ws2 = D.Workspace(ctx, n, b.total)
    outs2 = D.CanonOutputs(n, b.total, b.offsets.device, want_bytes=True, want_hash=True)
    lens2 = torch.empty(n, dtype=torch.int32, device=b.offsets.device)
    ctx._check(ctx._lib.ck_dev_canon_bytes(ctx.handle, torch.cuda.current_stream().cuda_stream, outs.out.data_ptr(),
                                           b.offsets.data_ptr(), n, b.total, 0, 0, outs2.out.data_ptr(),
                                           lens2.data_ptr(), outs2.start.data_ptr(), outs2.strand.data_ptr(),
                                           outs2.hash.data_ptr(), ws2.buf.data_ptr(), ws2.bytes))
    D.check(ctx, ws2)
    assert torch.equal(outs.out[: b.total], outs2.out[: b.total])
    assert torch.equal(outs.hash, outs2.hash)
    # duplicates (rotations / reverse complements of earlier originals) hash like their origin:
    table = D.DeviceTable(ctx, n)
    slots = torch.empty(n, dtype=torch.int64, device=b.offsets.device)
    first = torch.empty(n, dtype=torch.int64, device=b.offsets.device)
    table.insert(outs.hash, n, slots)
    table.first(slots, n, first)
    uniq = int((first == torch.arange(n, device=first.device)).sum().item())
    assert abs(uniq - 0.7 * n) < 0.01 * n            # ~30 % injected duplicates, all found, nothing else merged


@pytest.mark.parametrize("world", [2, 8])
def test_owner_partition_matches_owner_of(env, world):
    """ck_dev_owner_partition (send side of the hash-range exchange) against exchange.owner_of: every pair lands
    in its owner's run, runs are contiguous in owner order, pos maps records to their slot."""
    ctx, D, torch = env
    from circkit_b200.exchange import owner_of
    n, base = 200_003, 7_000_000_000
    g = torch.Generator(device="cuda").manual_seed(world)
    h = torch.randint(-2**63, 2**63 - 1, (n,), dtype=torch.int64, device="cuda", generator=g)
    h[:4] = torch.tensor([0, -1, 2**63 - 1, -2**63], dtype=torch.int64)
    part = D.OwnerPartitioner(ctx, n, world)
    pairs, pos, counts = part(h, base, world)
    hs, is_ = pairs[:, 0], pairs[:, 1]
    own = owner_of(h, world)
    assert counts == torch.bincount(own, minlength=world).tolist()
    assert torch.equal(hs[pos.long()], h)
    assert torch.equal(is_[pos.long()], torch.arange(base, base + n, device="cuda"))
    # the owner side over the same rows: first index per key == the table fed with separate arrays
    t1, t2 = D.DeviceTable(ctx, n), D.DeviceTable(ctx, n)
    s1 = torch.empty(n, dtype=torch.int64, device="cuda"); s2 = torch.empty_like(s1)
    f1 = torch.empty_like(s1); f2 = torch.empty_like(s1)
    dup = pairs.clone(); dup[n // 2:, 0] = dup[: n - n // 2, 0]            # force repeated keys
    t1.insert_pairs(dup, n, s1); t1.first(s1, n, f1)
    t2.insert(dup[:, 0].contiguous(), n, s2, index=dup[:, 1].contiguous()); t2.first(s2, n, f2)
    assert torch.equal(f1, f2) and int((f1 != dup[:, 1]).sum()) > 0
    assert torch.equal(torch.sort(pos.long()).values, torch.arange(n, device="cuda"))
    starts = np.concatenate([[0], np.cumsum(counts)])
    o_sorted = owner_of(hs, world).cpu().numpy()
    for r in range(world):
        assert (o_sorted[starts[r]: starts[r + 1]] == r).all()


@pytest.mark.parametrize("normalize", [True, False], ids=["cli-semantics", "lib-semantics"])
@pytest.mark.parametrize("skew", [0, 1, 7, 13])
def test_prepare_any_alignment_and_every_path(env, normalize, skew):
    """k_prepare (normalise + classify + pack) on raw bytes whose base address is not 16-byte aligned: records with and
    without dropped bytes, one and many 512-byte pieces, all three symbol lanes -- lengths and canonical forms against
    the oracle (needletail normalisation rules: src/canonicalize.rs:24)."""
    ctx, D, torch = env
    rng = np.random.default_rng(100 + skew + (50 if normalize else 0))
    alphabets = [b"ACGT", b"ACGTacgtuU", b"ACGTN-", b"ACGTNRYKMSWBDHV-", b"ACGTacgtNnUuRYx*.~", b"ACGT\n", b"ACGTNn \t\r\n"]
    seqs = []
    for n in list(range(0, 40)) + [495, 496, 497, 511, 512, 513, 527, 528, 529, 1023, 1024, 1025, 1500, 4000]:
        for alpha in alphabets:
            seqs.append(bytes(rng.choice(np.frombuffer(alpha, np.uint8), n).astype(np.uint8)))
    for _ in range(600):
        alpha = alphabets[int(rng.integers(len(alphabets)))]
        seqs.append(bytes(rng.choice(np.frombuffer(alpha, np.uint8), int(rng.integers(1, 900))).astype(np.uint8)))
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    off = np.zeros(len(seqs) + 1, dtype=np.int64); np.cumsum(lens, out=off[1:])
    arena = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    n, total = len(seqs), int(off[-1])
    dev = torch.device("cuda", ctx.device)
    buf = torch.zeros(total + 64, dtype=torch.uint8, device=dev)
    raw = buf[skew: skew + total]
    raw.copy_(torch.from_numpy(arena.copy()))
    offsets = torch.from_numpy(off).to(dev)
    ws = D.Workspace(ctx, n, total)
    outs = D.CanonOutputs(n, total, dev, want_bytes=True, want_hash=True, aligned=True)
    lens_out = torch.empty(n, dtype=torch.int32, device=dev)
    D.canon_bytes(ctx, raw, offsets, n, total, outs, lens_out, ws, normalize=normalize)
    D.check(ctx, ws)
    want = oracle.canonicalize_batch(arena, off.astype(np.uint64), normalize=normalize, threads=8)
    got_len = lens_out.cpu().numpy().astype(np.int64)
    assert np.array_equal(got_len, want["lens"].astype(np.int64))
    assert np.array_equal(outs.hash.cpu().numpy().astype(np.uint64), want["hash"])
    out = outs.out.cpu().numpy()
    starts = 32 * ((off[:-1] >> 5) + np.arange(n))
    for i in range(n):
        a, ln = int(off[i]), int(got_len[i])                    # the oracle keeps record i at its raw offset
        assert out[starts[i]: starts[i] + ln].tobytes() == want["out"][a: a + ln].tobytes(), (i, seqs[i][:60])


@pytest.mark.parametrize("world", [2, 8])
def test_padded_owner_partition(env, world):
    """ck_dev_owner_partition_padded: fixed-capacity buckets for the equal-split exchange.  Every record sits in its
    owner's bucket, the unused tail of every bucket carries index ~0, which the table skips (first index ~0), and an
    overflowing bucket (one key repeated) raises the flag instead of writing past the bucket."""
    ctx, D, torch = env
    from circkit_b200.exchange import owner_of, bucket_capacity
    n, base = 300_006, 5_000_000_000
    g = torch.Generator(device="cuda").manual_seed(10 + world)
    h = torch.randint(-2**63, 2**63 - 1, (n,), dtype=torch.int64, device="cuda", generator=g)
    h[n // 2:] = h[: n - n // 2].clone()                                      # repeated keys: first index < own index
    part = D.OwnerPartitioner(ctx, n, world)
    cap = part.bucket_capacity(n)
    assert cap == bucket_capacity(n, world)
    pairs, pos, state = part.padded(h, base, world)
    own = owner_of(h, world)
    assert state[:world].tolist() == torch.bincount(own, minlength=world).tolist() and int(state[world]) == 0
    assert pairs.shape == (world * cap, 2)
    p = pos.long()
    assert torch.equal(pairs[p, 0], h) and torch.equal(pairs[p, 1], torch.arange(base, base + n, device="cuda"))
    assert torch.equal(p // cap, own)
    used = torch.zeros(world * cap, dtype=torch.bool, device="cuda"); used[p] = True
    assert int(used.sum()) == n and bool((pairs[~used, 1] == -1).all())
    for o in range(world):                                                    # records first, padding behind them
        assert bool(used[o * cap: o * cap + int(state[o])].all())
    # owner side: padding is skipped and answers ~0; records get the minimum index of their key
    table = D.DeviceTable(ctx, n)
    m = world * cap
    slots = torch.empty(m, dtype=torch.int64, device="cuda"); first = torch.empty_like(slots)
    table.insert_pairs(pairs, m, slots); table.first(slots, m, first)
    assert bool((first[~used] == -1).all())
    want = torch.arange(base, base + n, device="cuda"); want[n // 2:] = want[: n - n // 2].clone()
    assert torch.equal(first[p], want)
    # overflow: every record carries the same key
    h2 = torch.full((n,), 12345, dtype=torch.int64, device="cuda")
    pairs2, pos2, state2 = part.padded(h2, base, world)
    assert int(state2[world]) == 1 and int(state2[0]) == n
    assert int((pos2.long() == 0xffffffff).sum() + (pos2 == -1).sum()) == n - cap


@pytest.mark.parametrize("want_bytes,want_hash", [(True, True), (True, False), (False, True), (False, False)])
def test_cli_alphabet_lane_kernel(env, want_bytes, want_hash):
    """ck_lane4.cuh: records over {-, A, C, G, N, T} after needletail's normalisation (src/canonicalize.rs:24-27; config 3 of
    BASELINE.json) -- every length around the oct (64-symbol) edges, the 129 / 241 fast-path limits and the 2048 class limit,
    random, IUPAC (normalised to N), and every tie shape (periodic, dimers, reverse-palindromic circles, N runs), all four
    kernel variants, start / strand / bytes / XXH3 against the oracle."""
    ctx, D, torch = env
    import random
    rng = random.Random(4 + 2 * want_bytes + want_hash)
    alpha = b"ACGTN-"

    def rnd(n, al=alpha):
        return bytes(rng.choice(al) for _ in range(n))
    seqs = []
    for n in list(range(120, 330)) + list(range(380, 390)) + [447, 448, 449, 511, 512, 513, 1000, 1023, 1024, 1025, 1984, 2047, 2048, 2049, 2100]:
        seqs.append(rnd(n))
        seqs.append(rnd(n, b"ACGTNRYKMSWBDHVacgtnu"))                   # normalised: IUPAC -> N
        k = rng.randrange(6)
        if k == 0:
            u = rnd(rng.randint(1, 40)); seqs.append((u * (n // len(u) + 1))[:n])           # near-periodic (cut power)
        elif k == 1:
            h = rnd(n // 2); seqs.append(h + h)                                             # dimer
        elif k == 2:
            h = rnd(n // 2); seqs.append(h + oracle.revcomp(h))                             # reverse-palindromic circle
        elif k == 3:
            s = bytearray(rnd(n)); a = rng.randrange(n)
            for j in range(rng.randint(8, n)):
                s[(a + j) % n] = ord("-")                                                   # long run of the smallest symbol
            seqs.append(bytes(s))
        elif k == 4:
            seqs.append(b"N" * n)
        else:
            s = bytearray(rnd(n, b"ACGT")); s[rng.randrange(n)] = ord("N"); seqs.append(bytes(s))   # a single N
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    off = np.zeros(len(seqs) + 1, dtype=np.int64); np.cumsum(lens, out=off[1:])
    arena = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    n, total = len(seqs), int(off[-1])
    dev = torch.device("cuda", ctx.device)
    raw = torch.from_numpy(arena.copy()).to(dev)
    offsets = torch.from_numpy(off).to(dev)
    ws = D.Workspace(ctx, n, total)
    outs = D.CanonOutputs(n, total, dev, want_bytes=want_bytes, want_hash=want_hash, aligned=True)
    lens_out = torch.empty(n, dtype=torch.int32, device=dev)
    ctx._lib.ck_kernel_timing(ctx.handle, 1)
    D.kernel_times(ctx)
    D.canon_bytes(ctx, raw, offsets, n, total, outs, lens_out, ws, normalize=True)
    D.check(ctx, ws)
    times = D.kernel_times(ctx)
    ctx._lib.ck_kernel_timing(ctx.handle, 0)
    assert times["4bit_lane_129_2048"][1] == 1 and times["pack4"][1] == 1, times      # the kernel under test did run
    want = oracle.canonicalize_batch(arena, off.astype(np.uint64), normalize=True, threads=8)
    assert np.array_equal(lens_out.cpu().numpy().astype(np.int64), want["lens"].astype(np.int64))
    bad = np.flatnonzero((outs.start.cpu().numpy().astype(np.int64) != want["start"].astype(np.int64))
                         | (outs.strand.cpu().numpy() != want["strand"]))
    assert bad.size == 0, (bad[:5], [seqs[i][:50] for i in bad[:3]])
    if want_hash:
        assert np.array_equal(outs.hash.cpu().numpy().astype(np.uint64), want["hash"])
    if want_bytes:
        out = outs.out.cpu().numpy()
        starts = 32 * ((off[:-1] >> 5) + np.arange(n))
        for i in range(n):
            a, ln = int(off[i]), int(lens[i])
            assert out[starts[i]: starts[i] + ln].tobytes() == want["out"][a: a + ln].tobytes(), (i, ln, seqs[i][:60])
