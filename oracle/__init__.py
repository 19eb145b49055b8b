"""CPU oracle for the canonicalize / uniq hot path -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of ``oracle/ck_oracle.c`` (a C restatement of the reference's
``lib/src/canonicalize.rs:5-63``, ``src/canonicalize.rs:21-44``, ``src/uniq.rs:33-78``
and of the third-party arithmetic under them).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package; ``circkit_b200`` never does.

Parity pinning: see ``tests/test_oracle.py`` (reference KATs + fixtures + python-xxhash).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libck_oracle.so")


def build(force: bool = False) -> str:
    """Compile ck_oracle.c with the committed Makefile (gcc only, no reference sources)."""
    src = os.path.join(_HERE, "ck_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True, env={**os.environ, "CC": "gcc"})
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, u64p, u32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
        for name in ("ck_o_lmsr_index", "ck_o_lmsr_index_simple", "ck_o_lmsr_index_2",
                     "ck_o_lmsr_index_faithful_cost"):
            getattr(L, name).restype = C.c_size_t
            getattr(L, name).argtypes = [C.c_char_p, C.c_size_t]
        L.ck_o_lmsr.restype = None
        L.ck_o_lmsr.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
        L.ck_o_revcomp.restype = None
        L.ck_o_revcomp.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
        L.ck_o_canonicalize.restype = C.c_int
        L.ck_o_canonicalize.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
        L.ck_o_canonical_start.restype = None
        L.ck_o_canonical_start.argtypes = [C.c_char_p, C.c_size_t, u32p, u8p]
        L.ck_o_normalize.restype = C.c_size_t
        L.ck_o_normalize.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.POINTER(C.c_int)]
        L.ck_o_xxh3_64.restype = C.c_uint64
        L.ck_o_xxh3_64.argtypes = [C.c_char_p, C.c_size_t]
        L.ck_o_complement_table.restype = None
        L.ck_o_complement_table.argtypes = [C.c_char_p]
        L.ck_o_canonicalize_batch.restype = C.c_int
        L.ck_o_canonicalize_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_int]
        L.ck_o_uniq_consume.restype = C.c_int
        L.ck_o_uniq_consume.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.ck_o_max_threads.restype = C.c_int
        _lib = L
    return _lib


# ---------------------------------------------------------------- single-record API
def lmsr_index(s: bytes) -> int:
    """lib/src/canonicalize.rs:5-36"""
    return lib().ck_o_lmsr_index(s, len(s))


def lmsr_index_simple(s: bytes) -> int:
    """lib/src/canonicalize.rs:154-164"""
    return lib().ck_o_lmsr_index_simple(s, len(s))


def lmsr_index_2(s: bytes) -> int:
    """lib/src/canonicalize.rs:168-214"""
    return lib().ck_o_lmsr_index_2(s, len(s))


def lmsr(s: bytes) -> bytes:
    """lib/src/canonicalize.rs:41-47"""
    out = C.create_string_buffer(len(s))
    lib().ck_o_lmsr(s, len(s), out)
    return out.raw[: len(s)]


def revcomp(s: bytes) -> bytes:
    """bio 1.3.1 alphabets::dna::revcomp (call site lib/src/canonicalize.rs:56)"""
    out = C.create_string_buffer(len(s))
    lib().ck_o_revcomp(s, len(s), out)
    return out.raw[: len(s)]


def complement_table() -> bytes:
    out = C.create_string_buffer(256)
    lib().ck_o_complement_table(out)
    return out.raw[:256]


def canonicalize(s: bytes) -> bytes:
    """lib/src/canonicalize.rs:54-63"""
    out = C.create_string_buffer(len(s))
    lib().ck_o_canonicalize(s, len(s), out)
    return out.raw[: len(s)]


def canonical_start(s: bytes) -> tuple[int, int]:
    st, sd = C.c_uint32(0), C.c_uint8(0)
    lib().ck_o_canonical_start(s, len(s), C.byref(st), C.byref(sd))
    return st.value, sd.value


def normalize(s: bytes) -> bytes:
    """needletail 0.5.1 sequence::normalize(seq, false) (src/canonicalize.rs:24-27)."""
    out = C.create_string_buffer(len(s) or 1)
    ch = C.c_int(0)
    m = lib().ck_o_normalize(s, len(s), out, C.byref(ch))
    return out.raw[:m]


def xxh3_64(s: bytes) -> int:
    """xxhash-rust 0.8.6 xxh3::xxh3_64 (src/uniq.rs:45)."""
    return lib().ck_o_xxh3_64(s, len(s))


def max_threads() -> int:
    return lib().ck_o_max_threads()


# ---------------------------------------------------------------- batch API (numpy)
def canonicalize_batch(arena: np.ndarray, offsets: np.ndarray, *, normalize: bool = False,
                       threads: int = 1, want_start: bool = True, want_hash: bool = True,
                       faithful_cost: bool = False):
    """Worker closure of src/canonicalize.rs:21-30 / src/uniq.rs:33-41 over a batch.

    arena: uint8[total], offsets: uint64[n+1].  Returns dict(out, lens, start, strand, hash).
    Record i's canonical bytes are out[offsets[i] : offsets[i] + lens[i]].
    """
    arena = np.ascontiguousarray(arena, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    out = np.zeros(max(len(arena), 1), dtype=np.uint8)
    lens = np.zeros(max(n, 1), dtype=np.uint32)
    start = np.zeros(max(n, 1), dtype=np.uint32) if want_start else None
    strand = np.zeros(max(n, 1), dtype=np.uint8) if want_start else None
    h = np.zeros(max(n, 1), dtype=np.uint64) if want_hash else None
    p = lambda a: a.ctypes.data if a is not None else None
    rc = lib().ck_o_canonicalize_batch(p(arena), p(offsets), n, 1 if normalize else 0, threads,
                                       p(out), p(lens), p(start), p(strand), p(h), 1 if faithful_cost else 0)
    assert rc == 0
    return dict(out=out[: len(arena)], lens=lens[:n],
                start=None if start is None else start[:n],
                strand=None if strand is None else strand[:n],
                hash=None if h is None else h[:n])


def uniq_consume(canon: np.ndarray, offsets: np.ndarray, lens: np.ndarray | None = None):
    """Consumer closure of src/uniq.rs:42-78 (serial): returns (hash[n], first_index[n])."""
    canon = np.ascontiguousarray(canon, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    h = np.zeros(max(n, 1), dtype=np.uint64)
    first = np.zeros(max(n, 1), dtype=np.uint64)
    lp = None
    if lens is not None:
        lens = np.ascontiguousarray(lens, dtype=np.uint32)
        lp = lens.ctypes.data
    rc = lib().ck_o_uniq_consume(canon.ctypes.data, offsets.ctypes.data, lp, n, h.ctypes.data, first.ctypes.data)
    assert rc == 0
    return h[:n], first[:n]
