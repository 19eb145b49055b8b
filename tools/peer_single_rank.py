"""One rank that is its own peer: runs the fused exchange kernels (k_owner_scatter_peers, k_peer_barrier,
k_table_insert_regions, k_table_first_regions, k_gather_first) of ck_dev_peer_first_index on the config-5 shard, for ncu (a
multi-rank command must not run under ncu).  world = 1: every pair stays on this GPU, the kernels and their access pattern are
the same."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import circkit_b200
from circkit_b200 import device as D

R = int(sys.argv[1]) if len(sys.argv) > 1 else 12_500_000
ctx = circkit_b200.Context(max_batch_bytes=0, max_batch_records=0)
dev = torch.device("cuda", 0)
b = D.synth_batch(ctx, seed=5, first_index=0, n_records=R, kind=0, lo=250, hi=400, dup_permille=300)
outs = D.CanonOutputs(R, b.total, dev, want_bytes=False, want_hash=True, aligned=True)
ws = D.Workspace(ctx, R)
D.canon_packed2(ctx, b, outs, ws, class_mask=D.class_mask_for(250, 400))
table = D.DeviceTable(ctx, int(R * 1.05) + 1024, dev=dev)
ctx.peer_attach([ctx.peer_export(1, 0, R)])
first = torch.empty(R, dtype=torch.int64, device=dev)
for _ in range(3):
    table.clear()
    ctx._check(ctx._lib.ck_dev_peer_first_index(ctx.handle, torch.cuda.current_stream().cuda_stream, outs.hash.data_ptr(), R, 0,
                                                table.buf.data_ptr(), table.bytes, first.data_ptr()))
torch.cuda.synchronize()
slots = torch.empty(R, dtype=torch.int64, device=dev); want = torch.empty(R, dtype=torch.int64, device=dev)
table.clear(); table.insert(outs.hash, R, slots, base_index=0); table.first(slots, R, want)
print("single-rank peer exchange == plain table:", bool(torch.equal(first, want)), "unique", int((first == torch.arange(R, device=dev)).sum()))
