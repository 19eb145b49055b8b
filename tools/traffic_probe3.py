"""DRAM traffic attribution of k_canon_s3 on a config-2 batch.  Run under
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:k_canon_s3 --csv --log-file X python tools/traffic_probe3.py
Launch order (two launches each, the second one counts): for CK_S3_DEBUG in (0, 0x100 no scan prefetch, 0x200 no emit prefetch,
0x700 no L2 prefetch at all): scan only (start/strand), + hash (emit reads, no stores), + hash + bytes."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import circkit_b200
from circkit_b200 import device as D

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 200
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
kind = 1 if hi > 1000 else 0
for dbg in ("0", "0x100", "0x200", "0x700"):
    os.environ["CK_S3_DEBUG"] = dbg
    ctx = circkit_b200.Context(max_batch_bytes=0, max_batch_records=0)
    b = D.synth_batch(ctx, seed=2, first_index=0, n_records=n, kind=kind, lo=lo, hi=hi, dup_permille=300)
    ws = D.Workspace(ctx, n)
    print("debug", dbg, "bases", b.total, "packed bytes", b.total // 4, flush=True)
    for wb, wh in ((False, False), (False, True), (True, True)):
        outs = D.CanonOutputs(n, b.total, b.offsets.device, want_bytes=wb, want_hash=wh, aligned=True)
        for _ in range(2):
            D.canon_packed2(ctx, b, outs, ws, class_mask=D.class_mask_for(lo, hi))
        torch.cuda.synchronize()
        del outs
    del b, ws
    ctx.close()
    torch.cuda.empty_cache()
