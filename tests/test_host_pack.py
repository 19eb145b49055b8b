"""ck_pack2_host (include/circkit_b200.h, "packed host input"): the multi-threaded host packer is plain host code, so it is
checked here without a GPU -- against a numpy restatement of the dense layout, needletail's normalisation rules
(src/canonicalize.rs:24-27, via the oracle's table) and the lane rules of k_prepare (ck_kernels.cuh)."""
import numpy as np
import pytest

import oracle
from circkit_b200 import core

ALPHABETS = [b"ACGT", b"ACGTacgtuU", b"ACGTN-", b"ACGTNRYKMSWBDHV-", b"ACGTacgtNnUuRYx*.~", b"ACGT\n", b"ACGTNn \t\r\n"]
SYM16 = set(b"-ABCDGHKMNRSTVWY")
CODE = {65: 0, 67: 1, 71: 2, 84: 3}


def make_batch(seed, n_random=300):
    rng = np.random.default_rng(seed)
    seqs = []
    for n in list(range(0, 40)) + [63, 64, 65, 127, 128, 129, 511, 512, 513, 1500]:
        for alpha in ALPHABETS:
            seqs.append(bytes(rng.choice(np.frombuffer(alpha, np.uint8), n).astype(np.uint8)))
    for _ in range(n_random):
        alpha = ALPHABETS[int(rng.integers(len(ALPHABETS)))]
        seqs.append(bytes(rng.choice(np.frombuffer(alpha, np.uint8), int(rng.integers(1, 900))).astype(np.uint8)))
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum([len(s) for s in seqs], out=off[1:])
    return seqs, np.frombuffer(b"".join(seqs), dtype=np.uint8), off


def expected(seq: bytes, normalize: bool):
    s = oracle.normalize(seq) if normalize else seq
    if all(b in CODE for b in s):
        lane = 2
    elif all(b in SYM16 for b in s):
        lane = 4
    else:
        lane = 8
    return s, lane


@pytest.mark.parametrize("normalize", [False, True], ids=["lib-semantics", "cli-semantics"])
@pytest.mark.parametrize("threads", [1, 3, 8])
def test_packer_matches_the_layout_and_the_normalisation_rules(normalize, threads):
    seqs, arena, off = make_batch(7 + threads)
    pb = core.pack2_host(arena, off, normalize=normalize, threads=threads)
    units = pb.dense.view(np.uint32)
    lane_pos = 0
    for i, seq in enumerate(seqs):
        s, lane = expected(seq, normalize)
        assert int(pb.lens[i]) == len(s), i
        assert int(pb.lane[i]) == lane, (i, seq[:40])
        if lane == 2:
            base = 2 * ((int(off[i]) >> 5) + i)
            for j in range((len(s) + 15) // 16):
                want = 0
                for k in range(16):
                    want = (want << 2) | (CODE[s[16 * j + k]] if 16 * j + k < len(s) else 0)
                assert int(units[base + j]) == want, (i, j)
        else:
            assert int(pb.lane_offsets[i]) == lane_pos
            assert pb.lane_bytes[lane_pos: lane_pos + len(s)].tobytes() == bytes(s), i
            lane_pos += len(s)
    assert pb.lane_total == lane_pos and int(pb.lane_offsets[len(seqs)]) == lane_pos


def test_packer_is_independent_of_the_thread_count():
    seqs, arena, off = make_batch(3, n_random=2000)
    ref = core.pack2_host(arena, off, normalize=True, threads=1)
    for t in (2, 5, 16):
        pb = core.pack2_host(arena, off, normalize=True, threads=t)
        assert np.array_equal(pb.lens, ref.lens) and np.array_equal(pb.lane, ref.lane)
        assert np.array_equal(pb.lane_offsets, ref.lane_offsets) and pb.lane_total == ref.lane_total
        assert np.array_equal(pb.lane_bytes[: pb.lane_total], ref.lane_bytes[: ref.lane_total])
        # dense words of the 2-bit records (the rest of the buffer is unspecified)
        u, r = pb.dense.view(np.uint32), ref.dense.view(np.uint32)
        for i in np.nonzero(ref.lane == 2)[0]:
            b = 2 * ((int(off[i]) >> 5) + int(i))
            k = (int(ref.lens[i]) + 15) // 16
            assert np.array_equal(u[b: b + k], r[b: b + k])


def test_packer_edge_cases():
    empty = core.pack2_host(np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert empty.n == 0 and empty.lane_total == 0
    only_empty = core.pack2_host(np.zeros(0, np.uint8), np.zeros(4, np.uint64), normalize=True)
    assert list(only_empty.lens) == [0, 0, 0] and list(only_empty.lane) == [2, 2, 2]
    with pytest.raises(core.CircKitError):
        core.pack2_host(np.frombuffer(b"ACGT", np.uint8), np.array([0, 3, 2, 4], dtype=np.uint64))   # offsets go backwards
