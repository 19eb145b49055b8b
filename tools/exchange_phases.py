"""Per-phase CUDA-event timing of the multi-GPU uniq step (config 5 shard per rank) for the three exchanges, and a check
that they give the same first indices.  Run under torchrun:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/exchange_phases.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import circkit_b200
from circkit_b200 import device as D, exchange as X

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
ctx = circkit_b200.Context(max_batch_bytes=0, max_batch_records=0, device=dev.index)
b = D.synth_batch(ctx, seed=5, first_index=rank * R, n_records=R, kind=0, lo=250, hi=400, dup_permille=300)
outs = D.CanonOutputs(R, b.total, dev, want_bytes=True, want_hash=True, aligned=True)
ws = D.Workspace(ctx, R)
table = D.DeviceTable(ctx, int(R * 1.05) + 1024, dev=dev)
part = D.OwnerPartitioner(ctx, R, world, dev)
peer = D.PeerExchange(ctx, R, world, rank, dev)
cap = part.bucket_capacity(R)
m = world * cap
slots = torch.empty(m, dtype=torch.int64, device=dev); fr = torch.empty(m, dtype=torch.int64, device=dev)
p_recv = torch.empty((m, 2), dtype=torch.int64, device=dev); f_back = torch.empty(m, dtype=torch.int64, device=dev)
first = torch.empty(R, dtype=torch.int64, device=dev)
mask = D.class_mask_for(250, 400)
D.canon_packed2(ctx, b, outs, ws, class_mask=mask)
h = outs.hash[:R]
lib = ctx._lib


def first_fn(hh, idx):
    k = hh.numel()
    s, o = torch.empty(max(k, 1), dtype=torch.int64, device=dev), torch.empty(max(k, 1), dtype=torch.int64, device=dev)
    table.insert(hh, k, s, index=idx, base_index=rank * R if idx is None else 0)
    table.first(s, k, o)
    return o[:k]


# ---- agreement of the three paths
table.clear(); f_exact = X.exchange_first_index(h, rank * R, first_fn, partition_fn=part).clone()
table.clear(); st = peer.first_index(h, rank * R, table, first); f_peer = first.clone()
ok = torch.equal(f_exact, f_peer) and int(st[world]) == 0
table.clear()


def fp(pairs):
    table.insert_pairs(pairs, m, slots); table.first(slots, m, fr); return fr


f_pad, st2 = X.exchange_first_index_padded(h, rank * R, fp, padded_fn=part.padded)
ok = ok and torch.equal(f_exact, f_pad) and int(st2[world]) == 0
t = torch.tensor([int(ok)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("world", world, "records/rank", R, "bucket capacity", cap, "| exact == peer == padded on every rank:", bool(t.item()),
          "| unique on rank 0:", int((f_exact == torch.arange(rank * R, rank * R + R, device=dev)).sum()))


def timed(names, fns, label):
    acc = {k: 0.0 for k in names}
    for it in range(8):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        dist.barrier(); torch.cuda.synchronize()
        for k, fn in enumerate(fns):
            ev[k].record(); fn()
        ev[len(fns)].record(); torch.cuda.synchronize()
        if it >= 3:
            for k, name in enumerate(names):
                acc[name] += ev[k].elapsed_time(ev[k + 1]) / 5
    if rank == 0:
        print(label)
        for name in names:
            print("  %-12s %.3f ms" % (name, acc[name]))
        print("  total        %.3f ms" % sum(acc.values()))


state = {}
timed(["canon", "clear", "partition", "a2a fwd", "insert", "first", "a2a back", "gather"],
      [lambda: D.canon_packed2(ctx, b, outs, ws, class_mask=mask), table.clear,
       lambda: state.update(p=part.padded(h, rank * R, world)),
       lambda: dist.all_to_all_single(p_recv, state["p"][0]),
       lambda: table.insert_pairs(p_recv, m, slots), lambda: table.first(slots, m, fr),
       lambda: dist.all_to_all_single(f_back, fr), lambda: first.copy_(f_back[state["p"][1].long()])],
      "padded buckets through NCCL all-to-all")
S = D._stream
timed(["canon", "clear", "scatter>peers", "barrier A", "insert", "first>peers", "barrier B", "gather"],
      [lambda: D.canon_packed2(ctx, b, outs, ws, class_mask=mask), table.clear,
       lambda: ctx._check(lib.ck_dev_owner_scatter_peers(ctx.handle, S(), D._p(h), R, rank * R, world, rank, peer.cap, peer.recv_ptrs,
                                                         D._p(peer.pos), D._p(peer.state))),
       lambda: peer.h_recv.barrier(),
       lambda: table.insert_pairs(peer.recv, peer.m, peer.slots),
       lambda: ctx._check(lib.ck_dev_table_first_peers(ctx.handle, S(), D._p(table.buf), table.bytes, D._p(peer.slots), world, rank,
                                                       peer.cap, peer.ret_ptrs)),
       lambda: peer.h_ret.barrier(),
       lambda: ctx._check(lib.ck_dev_gather_first(ctx.handle, S(), D._p(peer.ret), D._p(peer.pos), R, D._p(first)))],
      "kernels store into the peers' buffers")
dist.barrier()
dist.destroy_process_group()
