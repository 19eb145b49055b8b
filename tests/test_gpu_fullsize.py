"""BASELINE.json configurations at their FULL sizes on the GPU, checked through size-independent properties (the oracle
finishes in seconds only on a prefix, which is compared too):
  * the canonical form is a fixed point: feeding the canonical bytes back gives the same bytes and the same XXH3-64;
  * duplicates injected as random rotations / reverse complements collapse onto their originals (unique fraction);
  * lengths are preserved, every result slot is written (no record is dropped by a class / retry list);
  * equal inputs give equal outputs (config 3: the batch is a tiling of 500 k distinct records).
Config 1 at full size: tests/test_gpu_device_api.py::test_full_size_properties_config1."""
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


def _canon_again(ctx, D, torch, offsets, n, total, out_compact, normalize=False):
    """canonical bytes (compact layout) through the raw-bytes entry: normalise? + classify + pack + canonicalise"""
    ws = D.Workspace(ctx, n, total)
    outs = D.CanonOutputs(n, total, offsets.device, want_bytes=True, want_hash=True, aligned=True)
    lens = torch.empty(n, dtype=torch.int32, device=offsets.device)
    D.canon_bytes(ctx, out_compact, offsets, n, total, outs, lens, ws, normalize=normalize)
    D.check(ctx, ws)
    return outs, lens


def _compact(ctx, torch, outs, offsets, n, total, chunk_bytes=1 << 28):
    """aligned output arena -> compact layout (record i at offsets[i]), on the device, a few hundred MB at a time"""
    dev = offsets.device
    dst = torch.empty(total, dtype=torch.uint8, device=dev)
    shift = 32 * ((offsets[:-1] >> 5) + torch.arange(n, device=dev)) - offsets[:-1]
    base = int(offsets[0].item())
    cuts = torch.searchsorted(offsets, torch.arange(base, base + total, chunk_bytes, device=dev), right=True) - 1
    cuts = cuts.tolist() + [n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        if a == b:
            continue
        lo, hi = int(offsets[a].item()), int(offsets[b].item())
        src = torch.repeat_interleave(shift[a:b], offsets[a + 1:b + 1] - offsets[a:b]) + torch.arange(lo, hi, device=dev)
        dst[lo - base:hi - base] = outs.out[src]
    return dst


def _prefix_vs_oracle(ctx, D, torch, batch, outs, k):
    ascii_ = D.unpack_ascii(ctx, batch, k).cpu().numpy()
    off = batch.offsets[: k + 1].cpu().numpy().astype(np.uint64)
    want = oracle.canonicalize_batch(ascii_, off, normalize=False, threads=8)
    assert np.array_equal(outs.start[:k].cpu().numpy().astype(np.uint32), want["start"])
    assert np.array_equal(outs.strand[:k].cpu().numpy(), want["strand"])
    assert np.array_equal(outs.hash[:k].cpu().numpy().astype(np.uint64), want["hash"])
    got = _compact(ctx, torch, outs, batch.offsets[: k + 1], k, int(off[-1])).cpu().numpy()
    assert np.array_equal(got, want["out"])


def _unique_fraction(ctx, D, torch, outs, n):
    table = D.DeviceTable(ctx, n)
    slots = torch.empty(n, dtype=torch.int64, device=outs.hash.device)
    first = torch.empty(n, dtype=torch.int64, device=outs.hash.device)
    table.insert(outs.hash, n, slots)
    table.first(slots, n, first)
    return int((first == torch.arange(n, device=first.device)).sum().item()) / n


@pytest.mark.parametrize("name,n,kind,lo,hi,dup,adv,seed,prefix", [
    ("config 2: 10 M circRNA-length records, 30 % duplicates", 10_000_000, 1, 200, 5000, 300, 0, 2, 100_000),
    ("config 5 shard: 12.5 M viroid-length records, 30 % duplicates", 12_500_000, 0, 250, 400, 300, 0, 5, 100_000),
    ("config 4: 200 k plasmid-length records, 1 % adversarial", 200_000, 1, 5000, 200_000, 0, 10, 4, 3000),
])
def test_full_size_two_bit_configs(env, name, n, kind, lo, hi, dup, adv, seed, prefix):
    ctx, D, torch = env
    b = D.synth_batch(ctx, seed=seed, first_index=0, n_records=n, kind=kind, lo=lo, hi=hi, dup_permille=dup,
                      adversarial_permille=adv)
    outs = D.CanonOutputs(n, b.total, b.offsets.device, want_bytes=True, want_hash=True, aligned=True)
    outs.start.fill_(-1); outs.strand.fill_(7)
    ws = D.Workspace(ctx, n)
    D.canon_packed2(ctx, b, outs, ws, class_mask=D.class_mask_for(lo, hi))
    D.check(ctx, ws)                                         # no record outside the promised classes, none too long
    lens = b.lens
    assert bool(((outs.start >= 0) & (outs.start.long() < lens)).all()) and bool((outs.strand <= 1).all())
    _prefix_vs_oracle(ctx, D, torch, b, outs, prefix)
    if dup:
        assert abs(_unique_fraction(ctx, D, torch, outs, n) - (1 - dup / 1000)) < 0.01
    # fixed point: canonical bytes in, the same bytes and hashes out (start 0 on the forward strand unless the record is
    # its own reverse complement's rotation)
    canon = _compact(ctx, torch, outs, b.offsets, n, b.total)
    h1 = outs.hash.clone()
    del outs, ws
    outs2, lens2 = _canon_again(ctx, D, torch, b.offsets, n, b.total, canon)
    assert torch.equal(lens2.long(), lens)
    assert torch.equal(outs2.hash, h1)
    assert torch.equal(_compact(ctx, torch, outs2, b.offsets, n, b.total), canon)


def test_full_size_config3_iupac(env):
    """5 M mixed IUPAC / N records (byte-level lanes), CLI semantics: tiling of 500 k distinct records."""
    ctx, D, torch = env
    from circkit_b200 import synth_host
    tile, reps = 500_000, 10
    a_np, o_np = synth_host.make_iupac_records(tile, 250, 400, seed=3)
    dev = torch.device("cuda", ctx.device)
    n = tile * reps
    lens_t = torch.from_numpy((o_np[1:] - o_np[:-1]).astype(np.int64)).to(dev).repeat(reps)
    offsets = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(lens_t, 0, out=offsets[1:])
    total = int(offsets[-1].item())
    raw = torch.from_numpy(a_np).to(dev).repeat(reps)
    outs, lens = _canon_again(ctx, D, torch, offsets, n, total, raw, normalize=True)
    assert torch.equal(lens.long(), lens_t)                  # nothing is dropped by the normalisation here
    tile_bytes = int(o_np[-1])
    canon = _compact(ctx, torch, outs, offsets, n, total)
    # equal inputs, equal outputs (every tile against the first)
    assert torch.equal(outs.hash.view(reps, tile), outs.hash[:tile].expand(reps, tile))
    assert torch.equal(canon.view(reps, tile_bytes), canon[:tile_bytes].expand(reps, tile_bytes))
    # a prefix against the oracle (needletail normalisation + canonical form + XXH3)
    k = 20000
    want = oracle.canonicalize_batch(a_np[: int(o_np[k])], o_np[: k + 1], normalize=True, threads=8)
    assert np.array_equal(canon[: int(o_np[k])].cpu().numpy(), want["out"])
    assert np.array_equal(outs.hash[:k].cpu().numpy().astype(np.uint64), want["hash"])
    # fixed point under both semantics (the canonical form is upper-case {-,A,C,G,N,T})
    h1 = outs.hash.clone()
    del outs
    for normalize in (True, False):
        outs2, lens2 = _canon_again(ctx, D, torch, offsets, n, total, canon, normalize=normalize)
        assert torch.equal(outs2.hash, h1) and torch.equal(_compact(ctx, torch, outs2, offsets, n, total), canon)
        del outs2
