"""circkit_b200 -- B200 (sm_100a) implementation of circKit's canonicalize / uniq hot path.

Host-side mirror of the reference interface for this path:

    circkit::canonicalize(&[u8]) -> Vec<u8>           lib/src/lib.rs:3, lib/src/canonicalize.rs:54
    circkit::canonicalize::lmsr(&[u8]) -> Vec<u8>      lib/src/canonicalize.rs:41
    circkit::canonicalize::lmsr_index(&[u8]) -> usize  lib/src/canonicalize.rs:5
    circkit canonicalize / circkit uniq                src/canonicalize.rs:7, src/uniq.rs:15
    circkit::monomerize::Monomerizer                   lib/src/monomerize.rs:7-152 (circkit_b200.monomerize.Monomerizer)

All compute runs in the CUDA library (csrc/, C ABI in include/circkit_b200.h).  There is no CPU
fallback: without the built library or without a GPU the calls raise.
"""
from .core import Context, CircKitError, canonicalize, lmsr, lmsr_index, default_context  # noqa: F401
