"""Worker of tests/test_gpu_multirank.py (one process per GPU under torch.distributed.run): multi-GPU uniq through the
library's own peer group against (a) the exact NCCL exchange, (b) the single-table rebuild of the global first index on rank 0,
(c) the oracle's serial consumer, for resident shards and for rounds of host batches on both slots."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import circkit_b200
from circkit_b200 import core, device as D, exchange as X


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    R = 200_000
    ctx = circkit_b200.Context(device=dev.index, max_batch_bytes=32 << 20, max_batch_records=1 << 16, table_capacity=4 * R)
    peer = D.PeerGroup(ctx, R, world, rank)
    ok = True
    notes = []

    # ---- (a) + (b): resident shards, three consecutive batches of different sizes (buffers are reused)
    table = D.DeviceTable(ctx, capacity_keys=int(R * 1.3) + 1024, dev=dev)
    part = D.OwnerPartitioner(ctx, R, world, dev)
    for it, n in enumerate([R, R // 3 + rank * 17, R]):
        seed = 5 + it
        # every rank generates ITS shard of the same global record set; shards are contiguous in rank order and, to make the
        # shard sizes differ, padded to R indices per rank (indices rank * R .. rank * R + n)
        b = D.synth_batch(ctx, seed=seed, first_index=rank * R, n_records=n, kind=0, lo=250, hi=400, dup_permille=400)
        outs = D.CanonOutputs(n, b.total, dev, want_bytes=False, want_hash=True, aligned=True)
        ws = D.Workspace(ctx, n)
        D.canon_packed2(ctx, b, outs, ws)
        D.check(ctx, ws)
        h = outs.hash[:n]
        first_peer = torch.empty(n, dtype=torch.int64, device=dev)
        table.clear()
        dist.barrier()
        peer.first_index(h, rank * R, table, first_peer)
        torch.cuda.synchronize()

        def first_fn(hh, idx):
            k = hh.numel()
            s, o = torch.empty(max(k, 1), dtype=torch.int64, device=dev), torch.empty(max(k, 1), dtype=torch.int64, device=dev)
            table.insert(hh, k, s, index=idx, base_index=rank * R if idx is None else 0)
            table.first(s, k, o)
            return o[:k]
        table.clear()
        torch.cuda.synchronize()
        dist.barrier()
        first_exact = X.exchange_first_index(h, rank * R, first_fn, partition_fn=part)
        same = bool(torch.equal(first_peer, first_exact))
        # single-table rebuild on rank 0
        ns = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(ns, torch.tensor([n], dtype=torch.int64, device=dev))
        ns = [int(x.item()) for x in ns]
        hpad = torch.zeros(R, dtype=torch.int64, device=dev); hpad[:n] = h
        fpad = torch.zeros(R, dtype=torch.int64, device=dev); fpad[:n] = first_peer
        hs = [torch.empty(R, dtype=torch.int64, device=dev) for _ in range(world)] if rank == 0 else None
        fs = [torch.empty(R, dtype=torch.int64, device=dev) for _ in range(world)] if rank == 0 else None
        dist.gather(hpad, hs, dst=0); dist.gather(fpad, fs, dst=0)
        if rank == 0:
            big = D.DeviceTable(ctx, capacity_keys=int(R * world * 1.3) + 1024, dev=dev)
            slots = [torch.empty(R, dtype=torch.int64, device=dev) for _ in range(world)]
            for r in range(world):
                big.insert(hs[r], ns[r], slots[r], base_index=r * R)
            want = torch.empty(R, dtype=torch.int64, device=dev)
            for r in range(world):
                big.first(slots[r], ns[r], want)
                same = same and bool(torch.equal(want[:ns[r]], fs[r][:ns[r]]))
            del big
        ok = ok and same
        notes.append("resident batch %d (n=%d): %s" % (it, n, same))

    # ---- all records identical: every pair has ONE owner (what overflowed the fixed-capacity buckets)
    n = 50_000
    h = torch.full((n,), 0x1234_5678_9ABC_DEF0, dtype=torch.int64, device=dev)
    first_peer = torch.empty(n, dtype=torch.int64, device=dev)
    table.clear(); dist.barrier()
    peer.first_index(h, rank * R, table, first_peer)
    torch.cuda.synchronize()
    same = bool((first_peer == 0).all())
    ok = ok and same
    notes.append("one key on every rank: %s" % same)

    # ---- (c) rounds of host batches: batch k of the input goes to rank k % world in round k // world, both slots in flight
    import oracle
    rng = np.random.default_rng(99)
    n_batches, per = 4 * world, 3000
    seqs = []
    for i in range(n_batches * per):
        L = int(rng.integers(130, 900))
        seqs.append(bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), L).astype(np.uint8)))
    for i in range(0, len(seqs), 3):                      # duplicates of earlier records as rotations / reverse complements
        j = int(rng.integers(0, i + 1))
        s = seqs[j]; r = int(rng.integers(len(s)))
        t = s[r:] + s[:r]
        seqs[i] = oracle.revcomp(t) if rng.random() < 0.5 else t
    seqs[7] = b""; seqs[11] = b"ACGTNNRY" * 30
    off = np.zeros(len(seqs) + 1, dtype=np.uint64); np.cumsum([len(s) for s in seqs], out=off[1:])
    arena = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    want = oracle.canonicalize_batch(arena, off, normalize=True, threads=4)
    wh, wfirst = oracle.uniq_consume(want["out"], off, want["lens"])
    ctx.uniq_reset()
    dist.barrier()
    rounds = n_batches // world
    got = {}
    pend = []
    for j in range(rounds + 1):
        if j < rounds:
            k = j * world + rank
            lo, hi = k * per, (k + 1) * per
            o = off[lo: hi + 1] - off[lo]
            a = arena[int(off[lo]): int(off[hi])]
            if j % 2 == 0:
                pb = core.pack2_host(a, o, normalize=True, threads=2)
                ctx.uniq_submit_packed(j & 1, pb, lo, no_bytes=True)
                pend.append((j & 1, lo, pb.n, pb.total, pb))
            else:
                ctx.uniq_submit(j & 1, a, o, lo, normalize=True, no_bytes=True)
                pend.append((j & 1, lo, per, int(o[-1]), None))
        if j >= 1:
            slot, lo, n_, tot, _ = pend[j - 1]
            r = ctx.uniq_wait(slot, n_, tot, want_bytes=False)
            got[lo] = r
    same = True
    for lo, r in got.items():
        same = same and np.array_equal(r["first"], wfirst[lo: lo + per]) and np.array_equal(r["hash"], wh[lo: lo + per])
    ok = ok and same
    notes.append("host rounds on both slots vs the oracle's serial consumer: %s" % same)

    t = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIRANK", "OK" if int(t.item()) else "FAIL", "world", world, "|", "; ".join(notes))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) else 1)


if __name__ == "__main__":
    main()
