"""Host-side (numpy) generator of the BASELINE.json config-3 workload: mixed IUPAC / N records as raw bytes.

Configs 1, 2, 4 and 5 are generated on the device straight into the packed arena (csrc/ck_synth.cuh); config 3 is
about raw, un-normalised bytes, so it is drawn here and copied to HBM once, before anything is timed."""
from __future__ import annotations

import numpy as np


def make_iupac_records(n_records: int, lo: int, hi: int, seed: int):
    """BASELINE.json config 3: uniform lengths in [lo, hi], iid ACGT, then each base with
    probability 0.02 replaced by one of N R Y K M S W B D H V; 10 % of the records get a lowercase soft-masked run;
    in 1 % of the records every T becomes U."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(lo, hi + 1, size=n_records)
    offsets = np.zeros(n_records + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    total = int(offsets[-1])
    arena = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=total, dtype=np.uint8)].copy()
    sub = rng.random(total) < 0.02
    codes = np.frombuffer(b"NRYKMSWBDHV", dtype=np.uint8)
    arena[sub] = codes[rng.integers(0, len(codes), size=int(sub.sum()))]
    rec_of = np.repeat(np.arange(n_records), lens)                 # record of every base
    pos = np.arange(total) - np.repeat(offsets[:-1].astype(np.int64), lens)
    # lowercase run [a, a + L) in 10 % of the records
    masked = rng.random(n_records) < 0.10
    a = (rng.random(n_records) * lens).astype(np.int64)
    L = 1 + (rng.random(n_records) * lens * 0.5).astype(np.int64)
    in_run = masked[rec_of] & (pos >= a[rec_of]) & (pos < (a + L)[rec_of])
    arena[in_run] |= 0x20
    rna = rng.random(n_records) < 0.01
    t = rna[rec_of] & ((arena == ord("T")) | (arena == ord("t")))
    arena[t] += 1                                                   # T -> U, t -> u
    return arena, offsets
