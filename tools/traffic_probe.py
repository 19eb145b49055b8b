"""DRAM traffic attribution of the lane kernel on a config-2 batch: run under
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:k_canon_s2 --csv
with (start/strand only), (+hash), (+bytes), (+both): the differences are the emit pass's re-reads and the stores' fills."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import circkit_b200
from circkit_b200 import device as D

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 200
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
ctx = circkit_b200.Context(max_batch_bytes=0, max_batch_records=0)
gran = int(os.environ.get("CK_L2_GRAN", "0"))
if gran:
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    print("cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, %d) ->" % gran, rt.cudaDeviceSetLimit(5, ctypes.c_size_t(gran)))
    v = ctypes.c_size_t(0); rt.cudaDeviceGetLimit(ctypes.byref(v), 5); print("limit now", v.value)
b = D.synth_batch(ctx, seed=2, first_index=0, n_records=n, kind=1, lo=lo, hi=hi, dup_permille=300)
ws = D.Workspace(ctx, n)
print("bases", b.total, "packed bytes", b.total // 4)
for wb, wh in ((False, False), (True, True)):
    outs = D.CanonOutputs(n, b.total, b.offsets.device, want_bytes=wb, want_hash=wh, aligned=True)
    for _ in range(2):
        D.canon_packed2(ctx, b, outs, ws, class_mask=D.class_mask_for(lo, hi))
    torch.cuda.synchronize()
    del outs
