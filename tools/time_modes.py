"""Timing probe: lane kernel on records of (nearly) one length, in input order (direct mode) vs through the sorted work list."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import circkit_b200
from circkit_b200 import device as D

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 1450
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1550
ctx = circkit_b200.Context(max_batch_bytes=0, max_batch_records=0)
b = D.synth_batch(ctx, seed=7, first_index=0, n_records=n, kind=0, lo=lo, hi=hi)
outs = D.CanonOutputs(n, b.total, b.offsets.device, want_bytes=True, want_hash=True, aligned=True)
ws = D.Workspace(ctx, n)
for name, mask in (("direct (input order)", D.class_mask_for(lo, hi)), ("sorted list", D.class_mask_for(lo, hi) | D.class_mask_for(3000, 3001))):
    for _ in range(2):
        D.canon_packed2(ctx, b, outs, ws, class_mask=mask)
    torch.cuda.synchronize()
    ctx._lib.ck_kernel_timing(ctx.handle, 1); D.kernel_times(ctx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        D.canon_packed2(ctx, b, outs, ws, class_mask=mask)
    e1.record(); torch.cuda.synchronize()
    kt = D.kernel_times(ctx); ctx._lib.ck_kernel_timing(ctx.handle, 0)
    lane_ms = kt["2bit_lane_128_8192"][0] / max(kt["2bit_lane_128_8192"][1], 1)
    alg = float((8 * ((b.lens + 31) // 32) + b.lens + 24).sum().item())
    print("%-22s step %.3f ms  lane kernel %.3f ms  -> %.1f GB/s algorithmic (%.1f%% of 6551.7)" %
          (name, e0.elapsed_time(e1) / 3, lane_ms, alg / lane_ms / 1e6, 100 * alg / lane_ms / 1e6 / 6551.7))
