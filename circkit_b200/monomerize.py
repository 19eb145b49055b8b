"""`monomerize` on the GPU behind the reference's library interface (lib/src/monomerize.rs:7-152).

    m = Monomerizer(seed_len=10, overlap_min_identity=0.95)          # Monomerizer::builder()...build()
    m.monomerize(seq)                      -> bytes                  # :144-150
    m.monomerize_sensitive(seq)            -> bytes                  # :152-158
    m.first_monomer_end_index(seq)         -> int | None             # :50-99
    m.last_monomer_end_index(seq)          -> int | None             # :100-125
    m.last_monomer_end_index_sensitive(seq)-> int | None             # :127-141
    m.end_indices(list_of_seqs, sensitive) -> list[int | None]       # one launch for a batch (the worker closure of
                                                                     # src/monomerize.rs:83-101 without its normalisation)

Same names, argument meaning and errors as the reference's builder (seed_len 1..63; overlap_dist and
overlap_min_identity exclude each other).  Every call runs `k_monomerize` (csrc/ck_monomerize.cuh) through the C ABI
entry `ck_dev_monomerize`; there is no CPU path.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import numpy as np
import torch

from . import _native as N
from .core import Context
from .device import _p, _stream


class Monomerizer:
    def __init__(self, seed_len: int = None, overlap_dist: Optional[int] = None, overlap_min_identity: Optional[float] = None,
                 ctx: Optional[Context] = None):
        if overlap_dist is not None and overlap_min_identity is not None:
            raise ValueError("Both overlap_dist and overlap_min_identity are set. They are mutually exclusive since they may "
                             "produce conflicting filtering results.")
        if seed_len is None:
            raise ValueError("`seed_len` must be initialized")
        if not 1 <= seed_len <= 63:
            raise ValueError("Seed length must be at least 1 and at most 63 but was set to %d." % seed_len)
        self.seed_len, self.overlap_dist, self.overlap_min_identity = seed_len, overlap_dist, overlap_min_identity
        self._ctx, self._own = ctx, False

    def _context(self) -> Context:
        if self._ctx is None:
            self._ctx, self._own = Context(max_batch_bytes=0, max_batch_records=0), True
        return self._ctx

    # ---- batch form
    def end_indices(self, seqs: Iterable[bytes], sensitive: bool = False, first_only: bool = False) -> List[Optional[int]]:
        seqs = [bytes(s) for s in seqs]
        if not seqs:
            return []
        ctx = self._context()
        off = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum([len(s) for s in seqs], out=off[1:])
        dev = torch.device("cuda", ctx.device)
        raw = torch.from_numpy(np.frombuffer(b"".join(seqs) or b"\0", dtype=np.uint8).copy()).to(dev)
        out = self.end_indices_device(raw, torch.from_numpy(off).to(dev), len(seqs), sensitive, first_only)
        return [None if v == N.CK_MONO_NONE else int(v) for v in out.cpu().numpy().astype(np.uint32).tolist()]

    def end_indices_device(self, raw: torch.Tensor, offsets: torch.Tensor, n: int, sensitive: bool = False,
                           first_only: bool = False, lens: Optional[torch.Tensor] = None) -> torch.Tensor:
        """resident batch (uint8 bytes, int64 offsets[n + 1], optional int32 lens[n]: record i is its first lens[i] bytes)
        -> int32[n] end indices (bit pattern 0xffffffff = None)"""
        ctx = self._context()
        out = torch.empty(max(n, 1), dtype=torch.int32, device=raw.device)
        flags = (N.CK_MONO_SENSITIVE if sensitive else 0) | (N.CK_MONO_FIRST_ONLY if first_only else 0)
        ident = -1.0 if self.overlap_min_identity is None else float(self.overlap_min_identity)
        ctx._check(ctx._lib.ck_dev_monomerize(ctx.handle, _stream(), _p(raw), _p(offsets), _p(lens), n, self.seed_len,
                                              int(self.overlap_dist or 0), ident, flags, _p(out)))
        return out[:n]

    def normalize_device(self, raw: torch.Tensor, offsets: torch.Tensor, n: int):
        """needletail normalisation of a resident batch (the CLI's first step, src/monomerize.rs:86-89):
        -> (bytes uint8 like raw, lens int32[n])"""
        ctx = self._context()
        out = torch.empty(raw.numel() + 64, dtype=torch.uint8, device=raw.device)
        lens = torch.empty(max(n, 1), dtype=torch.int32, device=raw.device)
        ctx._check(ctx._lib.ck_dev_normalize(ctx.handle, _stream(), _p(raw), _p(offsets), n, _p(out), _p(lens)))
        return out, lens[:n]

    # ---- the reference's per-sequence methods
    def first_monomer_end_index(self, seq: bytes) -> Optional[int]:
        return self.end_indices([seq], first_only=True)[0]

    def last_monomer_end_index(self, seq: bytes) -> Optional[int]:
        return self.end_indices([seq])[0]

    def last_monomer_end_index_sensitive(self, seq: bytes) -> Optional[int]:
        return self.end_indices([seq], sensitive=True)[0]

    def monomerize(self, seq: bytes) -> bytes:
        end = self.last_monomer_end_index(seq)
        return bytes(seq) if end is None else bytes(seq[:end])

    def monomerize_sensitive(self, seq: bytes) -> bytes:
        end = self.last_monomer_end_index_sensitive(seq)
        return bytes(seq) if end is None else bytes(seq[:end])
