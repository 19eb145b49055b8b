"""numpy generator of BASELINE.json-shaped records for the CPU arms of bench.py -- TEST INFRASTRUCTURE.

Same laws as the device generator (ck_synth.cuh): lengths uniform or log-uniform in [lo, hi], iid ACGT,
dup_permille/1000 duplicates = random rotation (+ reverse complement w.p. 1/2) of an earlier original.
An independent draw of the same distribution, not the same bytes.
"""
from __future__ import annotations

import numpy as np

_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGT", b"TGCA"):
    _COMP[_a] = _b


def make_records(n_records: int, kind: int, lo: int, hi: int, dup_permille: int, seed: int):
    rng = np.random.default_rng(seed)
    is_dup = rng.random(n_records) < dup_permille / 1000.0
    is_dup[0] = False
    if kind == 0:
        lens = rng.integers(lo, hi + 1, size=n_records)
    else:
        lens = np.rint(np.exp(rng.uniform(np.log(lo), np.log(hi), size=n_records))).astype(np.int64)
        lens = np.clip(lens, lo, hi)
    orig_idx = np.flatnonzero(~is_dup)
    # origin of a duplicate: a uniformly chosen EARLIER original
    src = np.arange(n_records)
    dup_idx = np.flatnonzero(is_dup)
    if len(dup_idx):
        k = np.searchsorted(orig_idx, dup_idx)            # number of originals before each duplicate (>= 1)
        pick = (rng.random(len(dup_idx)) * k).astype(np.int64)
        src[dup_idx] = orig_idx[pick]
        lens[dup_idx] = lens[src[dup_idx]]
    offsets = np.zeros(n_records + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    arena = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=int(offsets[-1]), dtype=np.uint8)]
    arena = np.ascontiguousarray(arena)
    for i in dup_idx:
        s = arena[int(offsets[src[i]]): int(offsets[src[i] + 1])]
        r = int(rng.integers(0, len(s)))
        t = np.concatenate([s[r:], s[:r]])
        if rng.random() < 0.5:
            t = _COMP[t[::-1]]
        arena[int(offsets[i]): int(offsets[i + 1])] = t
    return arena, offsets

