"""World-size-2 (and 4) gloo tests of the multi-GPU uniq exchange (circkit_b200/exchange.py) on CPU.

The plumbing under test is what the N>1 bench path runs on NCCL: owner = hash range, personalised
all-to-all of (hash, global index), min index per key at the owner, reverse all-to-all.  The per-owner
first-occurrence step is played by a torch stand-in of the CUDA table (min index per key); the expected
answer is the oracle's serial consumer over the concatenated shards (src/uniq.rs:42-78).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _min_index_per_key(h: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """stand-in for k_table_insert + k_table_first: for every item, the minimum index among equal keys"""
    if h.numel() == 0:
        return idx.clone()
    uniq, inv = torch.unique(h, return_inverse=True)
    m = torch.full((uniq.numel(),), torch.iinfo(torch.int64).max, dtype=torch.int64)
    m.scatter_reduce_(0, inv, idx, reduce="amin")
    return m[inv]


def _worker(rank, world, port, hashes, expected, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from circkit_b200 import exchange as X
    n = len(hashes) // world
    lo, hi = rank * n, (rank + 1) * n if rank < world - 1 else len(hashes)
    # contiguous shards; the last rank takes the remainder, so bases differ from rank * n only there
    h = torch.from_numpy(hashes[lo:hi].view(np.int64).copy())
    first = X.exchange_first_index(h, lo, _min_index_per_key)
    ok = np.array_equal(first.numpy().astype(np.uint64), expected[lo:hi])
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def _torch_partition(h, base_index, world):
    """stand-in for device.OwnerPartitioner (ck_dev_owner_partition)"""
    from circkit_b200.exchange import owner_of
    owner = owner_of(h, world)
    order = torch.argsort(owner, stable=True)
    gidx = torch.arange(base_index, base_index + h.numel(), dtype=torch.int64)
    pairs = torch.stack([h[order], gidx[order]], dim=1).contiguous()
    pos = torch.empty_like(order)
    pos[order] = torch.arange(h.numel())
    return pairs, pos, torch.bincount(owner, minlength=world).tolist()


def _worker_two_phase(rank, world, port, hashes, expected, q):
    """the shard as three sub-batches: phase 1 (send + insert) of all of them, then phase 2 (query + return), as
    bench.py overlaps them on GPUs; the owner's table is a dict (min index per key, any insertion order)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from circkit_b200 import exchange as X
    n = len(hashes) // world
    lo, hi = rank * n, (rank + 1) * n if rank < world - 1 else len(hashes)
    table = {}

    def insert_pairs(pairs):
        for k, i in pairs.tolist():
            table[k] = min(table.get(k, i), i)
        return pairs[:, 0].clone()

    def query(keys, m):
        return torch.tensor([table[k] for k in keys.tolist()], dtype=torch.int64)

    cuts = [lo, lo + (hi - lo) // 3, lo + 2 * (hi - lo) // 3, hi]
    pend = []
    for a, b in reversed(list(zip(cuts[:-1], cuts[1:]))):          # any order of the sub-batches gives the same answer
        h = torch.from_numpy(hashes[a:b].view(np.int64).copy())
        pend.append((a, b, X.exchange_send(h, a, _torch_partition, insert_pairs)))
    ok = True
    for a, b, p in pend:
        first = X.exchange_finish(p, query)
        ok = ok and np.array_equal(first.numpy().astype(np.uint64), expected[a:b])
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_two_phase_exchange_matches_serial_consumer(world):
    import oracle
    from oracle import synth
    arena, off = synth.make_records(2500, 0, 60, 120, 400, seed=10 + world)
    res = oracle.canonicalize_batch(arena, off, normalize=True, threads=2, want_start=False)
    hashes, expected = oracle.uniq_consume(res["out"], off, res["lens"])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_two_phase, args=(r, world, port, hashes.copy(), expected, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(world))
    assert got == [(r, True) for r in range(world)]


@pytest.mark.parametrize("world", [2, 4])
def test_exchange_matches_serial_consumer(world):
    import oracle
    from oracle import synth
    arena, off = synth.make_records(3001, 0, 60, 120, 400, seed=world)
    res = oracle.canonicalize_batch(arena, off, normalize=True, threads=2, want_start=False)
    hashes, expected = oracle.uniq_consume(res["out"], off, res["lens"])
    # force the edge cases: hash with the top bit set / all ones / zero
    hashes = hashes.copy()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, hashes, expected, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(world))
    assert got == [(r, True) for r in range(world)]


def test_owner_of_covers_unsigned_range():
    from circkit_b200.exchange import owner_of
    h = np.array([0, 1, 2**63 - 1, 2**63, 2**64 - 1, 0x4000000000000000, 0xC000000000000000], dtype=np.uint64)
    t = torch.from_numpy(h.view(np.int64).copy())
    assert owner_of(t, 2).tolist() == [0, 0, 0, 1, 1, 0, 1]
    assert owner_of(t, 4).tolist() == [0, 0, 1, 2, 3, 1, 3]
    assert owner_of(t, 8).tolist() == [0, 0, 3, 4, 7, 2, 6]
    # monotone in the unsigned value: hash-range partition
    r = np.sort(np.random.default_rng(0).integers(0, 2**64, size=1000, dtype=np.uint64))
    o = owner_of(torch.from_numpy(r.view(np.int64).copy()), 8).numpy()
    assert (np.diff(o) >= 0).all() and o.min() == 0 and o.max() == 7


def _first_pairs_skipping_padding(pairs: torch.Tensor) -> torch.Tensor:
    """stand-in for ck_dev_table_insert_pairs + ck_dev_table_first: padding entries (index -1) answer -1"""
    out = torch.full((pairs.shape[0],), -1, dtype=torch.int64)
    real = pairs[:, 1] >= 0
    out[real] = _min_index_per_key(pairs[real, 0].contiguous(), pairs[real, 1].contiguous())
    return out


def _worker_padded(rank, world, port, hashes, expected, overflow_expected, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from circkit_b200 import exchange as X
    n = len(hashes) // world
    lo, hi = rank * n, (rank + 1) * n if rank < world - 1 else len(hashes)
    h = torch.from_numpy(hashes[lo:hi].view(np.int64).copy())
    first, state = X.exchange_first_index_padded(h, lo, _first_pairs_skipping_padding)
    # one flag for the whole job: every rank repeats through the exact path if any bucket anywhere overflowed
    flag = state[world: world + 1].clone().to(torch.int64)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    overflowed = bool(flag.item())
    if overflowed:
        first = X.exchange_first_index(h, lo, _min_index_per_key)
    ok = np.array_equal(first.numpy().astype(np.uint64), expected[lo:hi]) and overflowed == overflow_expected
    ok = ok and int(state[:world].sum()) == hi - lo
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,skewed", [(2, False), (3, False), (2, True)])
def test_padded_exchange_matches_serial_consumer(world, skewed):
    """fixed-capacity buckets + equal-split all-to-all; a skewed key set (most records share one key) overflows a
    bucket, is reported, and the exact path gives the answer"""
    import oracle
    from oracle import synth
    arena, off = synth.make_records(3000, 0, 60, 120, 400, seed=20 + world)
    res = oracle.canonicalize_batch(arena, off, normalize=True, threads=2, want_start=False)
    hashes, expected = oracle.uniq_consume(res["out"], off, res["lens"])
    if skewed:
        hashes = hashes.copy()
        hashes[::2] = hashes[0]                                   # half of the records are copies of record 0
        first_of = {}
        expected = np.array([first_of.setdefault(int(k), i) for i, k in enumerate(hashes)], dtype=np.uint64)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_padded, args=(r, world, port, hashes.copy(), expected, skewed, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(world))
    assert got == [(r, True) for r in range(world)]
