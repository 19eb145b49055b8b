"""Hash-range exchange for multi-GPU uniq.

Each rank canonicalises + hashes its contiguous shard; the owner of a key is the rank its top hash
bits select; ranks exchange (hash64, global_index) with one personalised all-to-all, the owner keeps the
minimum index per key (first occurrence in input order, src/uniq.rs:47-48), and a reverse all-to-all
returns first_index to the record's home rank.  The result is independent of the rank count because
min over global indices is associative and commutative.

Written against torch.distributed only (NCCL on GPUs, gloo in the CPU tests); the per-owner
first-occurrence step is passed in (`first_fn`) -- on GPUs it is the CUDA table of libcirckit_b200.
"""
from __future__ import annotations

from typing import Callable

import torch
import torch.distributed as dist


def owner_of(hash64: torch.Tensor, world: int) -> torch.Tensor:
    """Owner rank = floor(hash * world / 2^64) computed on the top 32 bits (hash carried as int64 bits)."""
    top = (hash64 >> 32) & 0xFFFFFFFF                      # arithmetic shift + mask == logical shift
    return (top * world) >> 32


def _all_to_all(send: torch.Tensor, send_counts: list[int], recv_counts: list[int], group=None) -> torch.Tensor:
    recv = torch.empty((sum(recv_counts),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
    backend = dist.get_backend(group)
    if backend == "nccl":
        dist.all_to_all_single(recv, send, recv_counts, send_counts, group=group)
        return recv
    # gloo has no all-to-all: personalised exchange with point-to-point ops
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    s_off = [0]
    for c in send_counts:
        s_off.append(s_off[-1] + c)
    r_off = [0]
    for c in recv_counts:
        r_off.append(r_off[-1] + c)
    recv[r_off[rank]: r_off[rank + 1]] = send[s_off[rank]: s_off[rank + 1]]
    reqs = []
    for p in range(world):
        if p == rank:
            continue
        if send_counts[p]:
            reqs.append(dist.isend(send[s_off[p]: s_off[p + 1]].contiguous(), p, group=group))
    bufs = {}
    for p in range(world):
        if p == rank or not recv_counts[p]:
            continue
        bufs[p] = torch.empty((recv_counts[p],) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        reqs.append(dist.irecv(bufs[p], p, group=group))
    for r in reqs:
        r.wait()
    for p, b in bufs.items():
        recv[r_off[p]: r_off[p + 1]] = b
    return recv


def _exchange_counts(send_counts_t: torch.Tensor, group=None) -> list[int]:
    """how many items every peer sends to this rank"""
    world = dist.get_world_size(group)
    if dist.get_backend(group) == "nccl":
        recv_counts_t = torch.empty_like(send_counts_t)
        dist.all_to_all_single(recv_counts_t, send_counts_t, group=group)
    else:
        gathered = [torch.empty_like(send_counts_t) for _ in range(world)]
        dist.all_gather(gathered, send_counts_t, group=group)
        recv_counts_t = torch.stack(gathered)[:, dist.get_rank(group)].contiguous()
    return [int(x) for x in recv_counts_t.tolist()]


def exchange_first_index(hash64: torch.Tensor, base_index: int,
                         first_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], group=None,
                         partition_fn=None, first_pairs_fn=None) -> torch.Tensor:
    """first_index[i] (global) for every local record i.

    hash64     int64[n]  XXH3-64 of the local records' canonical forms (bit pattern)
    base_index           global input index of local record 0 (shards are contiguous, rank order)
    first_fn(h, idx) ->  int64[m]: for the items this rank owns, the minimum global index per key
                         (idx None in the single-rank case: the index of item j is base_index + j)
    partition_fn(h, base_index, world) -> (pairs int64[n, 2], pos, send_counts): bucket (hash, index) by owner on the
                         device (device.OwnerPartitioner); default: the same with torch ops (CPU tests)
    first_pairs_fn(pairs) -> int64[m]: first_fn over the received (hash, index) rows without splitting them
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n = hash64.numel()
    if world == 1:
        return first_fn(hash64, None)                      # idx None: global index = base_index + position
    if partition_fn is not None:
        pairs, pos, send_counts = partition_fn(hash64, base_index, world)
        send_counts_t = torch.tensor(send_counts, dtype=torch.int64, device=hash64.device)
    else:
        owner = owner_of(hash64, world)
        order = torch.argsort(owner, stable=True)          # bucket by owner (order inside a bucket is free)
        send_counts_t = torch.bincount(owner, minlength=world)
        gidx = torch.arange(base_index, base_index + n, dtype=torch.int64, device=hash64.device)
        pairs = torch.stack([hash64[order], gidx[order]], dim=1).contiguous()
        pos = torch.empty_like(order)
        pos[order] = torch.arange(n, dtype=order.dtype, device=order.device)
        send_counts = [int(x) for x in send_counts_t.tolist()]
    recv_counts = _exchange_counts(send_counts_t, group)
    p_recv = _all_to_all(pairs, send_counts, recv_counts, group)          # one exchange: 16 bytes per record
    if first_pairs_fn is not None:
        f_recv = first_pairs_fn(p_recv)                    # owner side: min global index per key
    else:
        f_recv = first_fn(p_recv[:, 0].contiguous(), p_recv[:, 1].contiguous())
    f_back = _all_to_all(f_recv.contiguous(), recv_counts, send_counts, group)
    return f_back[pos]                                     # back to input order


def bucket_capacity(n: int, world: int) -> int:
    """entries per fixed-capacity bucket: the mean n / world plus 8 standard deviations of a uniform split
    (the same figure as device.exchange_bucket_capacity)"""
    mean = n / world
    return int(mean + 8.0 * (mean ** 0.5) + 64) + 1


def _padded_partition_torch(hash64: torch.Tensor, base_index: int, world: int):
    """device.OwnerPartitioner.padded with torch ops (CPU tests): fixed-capacity buckets, index -1 = no record"""
    n = hash64.numel()
    cap = bucket_capacity(n, world)
    owner = owner_of(hash64, world)
    order = torch.argsort(owner, stable=True)
    counts = torch.bincount(owner, minlength=world)
    starts = torch.cumsum(counts, 0) - counts
    rank_in_bucket = torch.arange(n, dtype=torch.int64, device=hash64.device) - starts[owner[order]]
    pairs = torch.zeros((world * cap, 2), dtype=torch.int64, device=hash64.device)
    pairs[:, 1] = -1
    ok = rank_in_bucket < cap
    dst = owner[order] * cap + rank_in_bucket
    gidx = torch.arange(base_index, base_index + n, dtype=torch.int64, device=hash64.device)
    pairs[dst[ok], 0] = hash64[order][ok]
    pairs[dst[ok], 1] = gidx[order][ok]
    pos = torch.full((n,), -1, dtype=torch.int64, device=hash64.device)
    pos[order[ok]] = dst[ok]
    state = torch.cat([counts.to(torch.int32), (~ok).any().to(torch.int32).reshape(1)])
    return pairs, pos, state


def _all_to_all_equal(send: torch.Tensor, group=None) -> torch.Tensor:
    world = dist.get_world_size(group)
    if dist.get_backend(group) == "nccl":
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=group)
        return recv
    each = send.shape[0] // world
    return _all_to_all(send, [each] * world, [each] * world, group)


def exchange_first_index_padded(hash64: torch.Tensor, base_index: int, first_pairs_fn, group=None, padded_fn=None):
    """exchange_first_index without counts and without host synchronisation: every rank sends `world` buckets of one
    fixed capacity (bucket_capacity: mean + 8 sigma of a uniform split, < 1 % padding at a million records per
    bucket) with an equal-split all-to-all; unused entries carry index -1, which first_pairs_fn must answer with -1.

    -> (first_index int64[n], state int32[world + 1]).  state[world] != 0 means that a bucket overflowed (a heavily
    duplicated key set) and first_index is NOT valid: the caller repeats the batch through exchange_first_index.
    The check is left to the caller so that it can be made once per batch, after everything has been queued."""
    world = dist.get_world_size(group)
    pairs, pos, state = (padded_fn or _padded_partition_torch)(hash64, base_index, world)
    p_recv = _all_to_all_equal(pairs, group)
    f_recv = first_pairs_fn(p_recv)
    f_back = _all_to_all_equal(f_recv.contiguous(), group)
    return f_back[pos.long()], state


# ---- the same exchange in two phases, so that a rank can overlap it with the canonicalisation of its next sub-batch ----
class PendingExchange:
    """State between exchange_send and exchange_finish: where every local record went and what this rank received."""
    __slots__ = ("pos", "send_counts", "recv_counts", "handle", "m")

    def __init__(self, pos, send_counts, recv_counts, handle, m):
        self.pos, self.send_counts, self.recv_counts, self.handle, self.m = pos, send_counts, recv_counts, handle, m


def exchange_send(hash64: torch.Tensor, base_index: int, partition_fn, insert_pairs_fn, group=None) -> PendingExchange:
    """Phase 1: bucket the local (hash, global index) pairs by owner, deliver them with one all-to-all and insert what
    arrives into this rank's first-occurrence table (insert_pairs_fn(pairs) -> handle for the later query).  The table
    keeps the minimum index per key, so phase 1 of any number of sub-batches may run in any order on any rank."""
    world = dist.get_world_size(group)
    pairs, pos, send_counts = partition_fn(hash64, base_index, world)
    send_counts_t = torch.tensor(send_counts, dtype=torch.int64, device=hash64.device)
    recv_counts = _exchange_counts(send_counts_t, group)
    p_recv = _all_to_all(pairs, send_counts, recv_counts, group)
    return PendingExchange(pos, send_counts, recv_counts, insert_pairs_fn(p_recv), p_recv.shape[0])


def exchange_finish(p: PendingExchange, first_query_fn, group=None) -> torch.Tensor:
    """Phase 2, after phase 1 of EVERY sub-batch has completed on this rank (all-to-alls are collective, so by then
    every pair owned by this rank has been inserted): first_query_fn(handle, m) -> int64[m] first index of the received
    pairs; a reverse all-to-all returns them; the result is first_index of the local records in input order."""
    f_recv = first_query_fn(p.handle, p.m)
    f_back = _all_to_all(f_recv.contiguous(), p.recv_counts, p.send_counts, group)
    return f_back[p.pos]
