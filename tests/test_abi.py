"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/circkit_b200.h declares.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from circkit_b200 import build
    return build.build()


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "circkit_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(ck_[a-z0-9_]+)\s*\(", hdr)))


def test_header_declares_the_boundary():
    syms = _declared_symbols()
    for must in ("ck_init", "ck_destroy", "ck_canon_submit", "ck_canon_wait", "ck_uniq_submit", "ck_uniq_wait",
                 "ck_lmsr_index", "ck_lmsr", "ck_canonicalize", "ck_last_error", "ck_alloc_pinned"):
        assert must in syms


def test_library_exports_every_declared_symbol(built_lib):
    L = ctypes.CDLL(built_lib)
    for s in _declared_symbols():
        assert hasattr(L, s), f"{s} declared in include/circkit_b200.h but not exported"


def test_python_binding_covers_every_declared_symbol(built_lib):
    from circkit_b200 import _native
    assert sorted(_native.SIGNATURES) == _declared_symbols()
    _native.lib()


def test_no_signature_mentions_torch():
    hdr = open(os.path.join(ROOT, "include", "circkit_b200.h")).read()
    assert "torch" not in hdr and "at::" not in hdr


def test_init_fails_loudly_without_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import circkit_b200
    with pytest.raises(circkit_b200.CircKitError) as e:
        circkit_b200.Context()
    assert e.value.code == -1          # CK_ERR_CUDA; no CPU fallback exists


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "circkit_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("ck_oracle_free", ""), f"{f} mentions the oracle"


def test_class_masks_match_the_header():
    """device.class_mask_for / CLASS_NAMES agree with the CK_CLASS_* promise bits of include/circkit_b200.h."""
    import re
    from circkit_b200 import device as D
    hdr = open(os.path.join(ROOT, "include", "circkit_b200.h")).read()
    bits = {m.group(1).lower(): int(m.group(2)) for m in re.finditer(r"#define CK_CLASS_(\w+) \(1u << (\d+)\)", hdr)}
    assert len(bits) == 6
    for name, bit in bits.items():
        assert D.CLASS_NAMES[bit] == name.replace("2bit", "2bit"), (name, bit)
    assert D.class_mask_for(250, 400) == 1 << 0
    assert D.class_mask_for(200, 5000) == (1 << 0) | (1 << 1) | (1 << 10) | (1 << 11)
    assert D.class_mask_for(5000, 200000) == (1 << 11) | (1 << 2) | (1 << 3)
    assert D.class_mask_for(513, 513) == 1 << 1


def test_table_sizing_is_one_and_a_half_slots_per_key():
    """ck_dev_table_bytes: a 64-byte header + 16-byte slots, 1.5 per key (any number, no power-of-two rounding)."""
    from circkit_b200 import _native as N
    lib = N.lib()
    for keys in (0, 1, 1000, 10_500_000 + 1024, 105_000_000):
        b = int(lib.ck_dev_table_bytes(keys))
        slots = (b - 64) // 16
        assert (b - 64) % 16 == 0 and slots >= 1024
        assert slots >= keys + keys // 2 and slots <= max(1024, keys + keys // 2 + 16)
