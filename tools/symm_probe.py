"""Does torch symmetric memory rendezvous work on this box?  (peer-mapped buffers for the fused exchange)"""
import os, sys
import torch
import torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem
print(rank, "symm_mem api:", [a for a in dir(symm_mem) if not a.startswith("_")][:40], flush=True)
t = symm_mem.empty(1 << 20, dtype=torch.int64, device=dev)
hdl = symm_mem.rendezvous(t, dist.group.WORLD)
print(rank, "handle:", type(hdl).__name__, [a for a in dir(hdl) if not a.startswith("_")], flush=True)
print(rank, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal_pad_ptrs", [hex(p) for p in hdl.signal_pad_ptrs][:2], flush=True)
t.fill_(rank)
hdl.barrier()
peer = (rank + 1) % world
pt = hdl.get_buffer(peer, (1 << 20,), torch.int64)
pt[rank * 16: rank * 16 + 16] = 1000 + rank          # remote write into the peer's buffer
hdl.barrier()
torch.cuda.synchronize()
src = (rank - 1) % world
print(rank, "after remote write:", t[src * 16: src * 16 + 4].tolist(), t[:2].tolist() if src != 0 else "", flush=True)
dist.barrier()
dist.destroy_process_group()
