"""Timing of k_monomerize on a resident synthetic batch: concatemers (2.3 copies of a 250-400 nt unit, 1 % substitutions),
seed 10, min identity 0.95 -- the reference CLI's defaults on rolling-circle-like reads.  CUDA events, inputs > L2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import circkit_b200
from circkit_b200.monomerize import Monomerizer

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rng = np.random.default_rng(1)
unit_len = rng.integers(250, 401, n)
total_len = (unit_len * 2.3).astype(np.int64)
off = np.zeros(n + 1, dtype=np.int64); np.cumsum(total_len, out=off[1:])
dev = torch.device("cuda", 0)
offsets = torch.from_numpy(off).to(dev)
T = int(off[-1])
# record r, position j -> unit[r][j mod unit_len[r]]: build on the device from per-record random units
rec = torch.repeat_interleave(torch.arange(n, device=dev), torch.from_numpy(total_len).to(dev))
pos = torch.arange(T, device=dev) - offsets[rec]
ul = torch.from_numpy(unit_len).to(dev)[rec]
g = torch.Generator(device=dev).manual_seed(7)
h = (rec * 1000003 + (pos % ul)) * 2654435761 % 4294967296
letters = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
raw = letters[((h >> 13) ^ (h >> 7)) & 3]
mut = torch.rand(T, device=dev, generator=g) < 0.01
raw[mut] = letters[torch.randint(0, 4, (int(mut.sum()),), device=dev, generator=g)]
del rec, pos, ul, h, mut
ctx = circkit_b200.Context(max_batch_bytes=0, max_batch_records=0)
for sensitive in (False, True):
    m = Monomerizer(10, overlap_min_identity=0.95, ctx=ctx)
    for _ in range(3):
        out = m.end_indices_device(raw, offsets, n, sensitive=sensitive)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = m.end_indices_device(raw, offsets, n, sensitive=sensitive)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    found = int((out != -1).sum())
    print("monomerize%s: %d records, %.2f GB raw, %.3f ms/launch, %.1f M records/s, %.0f GB/s of record bytes (%.1f %% of 6551.7), %d monomerized"
          % (" --sensitive" if sensitive else "", n, T / 1e9, ms, n / ms / 1e3, T / ms / 1e6, 100 * T / ms / 1e6 / 6551.7, found))
