"""CPU oracle for `monomerize` -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Plain-Python restatement of the reference's lib/src/monomerize.rs:48-152:
  first_monomer_end_index            :50-99    seed = last seed_len bytes; every occurrence of the seed in seq[..len - seed_len]
                                               in increasing start position (bio 1.3.1 ShiftAnd::find_all yields start
                                               positions left to right, overlapping matches included); overlap = prefix
                                               that ends with that occurrence, compared by Hamming distance with the suffix
                                               of the same length; the first occurrence within the allowed distance wins
  last_monomer_end_index             :100-125  repeat on the monomer until nothing more is found
  last_monomer_end_index_sensitive   :127-141  then one pass over the reverse complement of the monomer (bio dna::revcomp)
  monomerize / monomerize_sensitive  :144-158  the slice
Validation rules of the builder (:20-40): 1 <= seed_len <= 63; overlap_dist and overlap_min_identity exclude each other.

Pinned on every test of lib/src/monomerize.rs:155-554 (tests/golden/monomerize_kats.json, tests/test_oracle_monomerize.py).
"""
from __future__ import annotations

import math
from typing import Optional

_COMP = None


def _comp_table() -> bytes:
    """bio 1.3.1 alphabets::dna complement (the same table the canonicalize oracle uses)"""
    global _COMP
    if _COMP is None:
        import oracle
        _COMP = bytes(oracle.complement_table())
    return _COMP


class Monomerizer:
    def __init__(self, seed_len: int, overlap_dist: Optional[int] = None, overlap_min_identity: Optional[float] = None):
        if overlap_dist is not None and overlap_min_identity is not None:
            raise ValueError("Both overlap_dist and overlap_min_identity are set. They are mutually exclusive")
        if seed_len is None:
            raise ValueError("seed_len must be set")
        if not 1 <= seed_len <= 63:
            raise ValueError("Seed length must be at least 1 and at most 63 but was set to %d." % seed_len)
        self.seed_len, self.overlap_dist, self.overlap_min_identity = seed_len, overlap_dist, overlap_min_identity

    def max_dist(self, overlap_len: int) -> int:
        if self.overlap_min_identity is not None:      # len - floor(len * identity) in f64, lib/src/monomerize.rs:72-78
            return overlap_len - int(math.floor(float(overlap_len) * self.overlap_min_identity))
        return self.overlap_dist or 0

    def first_monomer_end_index(self, seq: bytes) -> Optional[int]:
        s, n = self.seed_len, len(seq)
        if n <= s:
            return None
        seed = seq[n - s:]
        text = seq[: n - s]
        occ = text.find(seed)
        while occ >= 0:
            m = occ + s
            successor, starter = seq[:m], seq[n - m:]
            dist = sum(1 for a, b in zip(starter, successor) if a != b)
            if dist <= self.max_dist(m):
                return n - m
            occ = text.find(seed, occ + 1)
        return None

    def last_monomer_end_index(self, seq: bytes) -> Optional[int]:
        idx = self.first_monomer_end_index(seq)
        while idx is not None:
            nxt = self.first_monomer_end_index(seq[:idx])
            if nxt is None:
                break
            idx = nxt
        return idx

    def last_monomer_end_index_sensitive(self, seq: bytes) -> Optional[int]:
        idx = self.last_monomer_end_index(seq)
        monomer = seq[: len(seq) if idx is None else idx]
        rc = monomer.translate(_comp_table())[::-1]
        k = self.first_monomer_end_index(rc)
        if k is None:
            return idx
        return (len(seq) if idx is None else idx) - (len(monomer) - k)

    def monomerize(self, seq: bytes) -> bytes:
        end = self.last_monomer_end_index(seq)
        return seq if end is None else seq[:end]

    def monomerize_sensitive(self, seq: bytes) -> bytes:
        end = self.last_monomer_end_index_sensitive(seq)
        return seq if end is None else seq[:end]


# ---- the compiled restatement (oracle/ck_oracle.c: ck_o_monomer_end, ck_o_monomerize_batch) -------------------------------
def c_end_index(seq: bytes, seed_len: int, overlap_dist: Optional[int] = None, overlap_min_identity: Optional[float] = None,
                sensitive: bool = False, first_only: bool = False) -> Optional[int]:
    import ctypes as C
    import oracle
    L = oracle.lib()
    L.ck_o_monomer_end.restype = C.c_size_t
    L.ck_o_monomer_end.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, C.c_uint64, C.c_double, C.c_int, C.c_int]
    r = L.ck_o_monomer_end(seq, len(seq), seed_len, int(overlap_dist or 0),
                           -1.0 if overlap_min_identity is None else float(overlap_min_identity), int(sensitive), int(first_only))
    return None if r == C.c_size_t(-1).value else int(r)


def c_end_indices_batch(arena, offsets, seed_len: int, overlap_dist: Optional[int] = None,
                        overlap_min_identity: Optional[float] = None, sensitive: bool = False, threads: int = 1):
    """numpy uint8 arena + uint64 offsets[n + 1] -> uint32[n] end indices (0xffffffff = None), `threads` host threads"""
    import ctypes as C
    import numpy as np
    import oracle
    L = oracle.lib()
    L.ck_o_monomerize_batch.restype = C.c_int
    L.ck_o_monomerize_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_size_t, C.c_uint64, C.c_double, C.c_int, C.c_int,
                                        C.c_void_p]
    arena = np.ascontiguousarray(arena, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    out = np.empty(len(offsets) - 1, dtype=np.uint32)
    L.ck_o_monomerize_batch(arena.ctypes.data, offsets.ctypes.data, len(offsets) - 1, seed_len, int(overlap_dist or 0),
                            -1.0 if overlap_min_identity is None else float(overlap_min_identity), int(sensitive), threads,
                            out.ctypes.data)
    return out
