"""Device-resident front end: torch tensors as HBM buffers + current CUDA stream, kernels via the C ABI.

torch is plumbing here (allocation, streams, torch.distributed); every kernel that touches record data
is launched by libcirckit_b200.so.  64-bit unsigned buffers are carried as torch.int64 (same bits).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _native as N
from .core import Context

CLASS_NAMES = ["2bit_le_512", "2bit_le_2048", "2bit_le_65536", "2bit_le_425984", "4bit_le_2048", "4bit_le_212992",
               "byte_le_1024", "byte_le_106496", "huge", "empty", "2bit_le_4096", "2bit_le_8192", "2bit_lane_128_8192",
               "table_insert", "table_first", "2bit_seg_8193_425984", "4bit_lane_129_2048", "pack4", "prepare"]
# length range [lo, hi] of each 2-bit class (class_mask bit = index in CLASS_NAMES)
BYTE_CLASSES = ("4bit_le_2048", "4bit_le_212992", "byte_le_1024", "byte_le_106496")
CLASS_RANGE = {"2bit_le_512": (1, 512), "2bit_le_2048": (513, 2048), "2bit_le_4096": (2049, 4096), "2bit_le_8192": (4097, 8192),
               "2bit_le_65536": (8193, 65536), "2bit_le_425984": (65537, 425984), "2bit_lane_128_8192": (1, 8192), "2bit_seg_8193_425984": (8193, 425984)}


def class_mask_for(lo: int, hi: int) -> int:
    """class_mask promise for 2-bit records with lengths in [lo, hi]."""
    m = 0
    for name, (a, b) in CLASS_RANGE.items():
        if "lane" not in name and "seg" not in name and lo <= b and hi >= a:
            m |= 1 << CLASS_NAMES.index(name)
    return m


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


class DeviceBatch:
    """A packed 2-bit batch resident in HBM: offsets[int64, n+1], packed2[int64 words]."""

    def __init__(self, offsets: torch.Tensor, packed2: torch.Tensor, n_records: int, total: int, single: bool = False):
        self.offsets, self.packed2, self.n, self.total = offsets, packed2, n_records, total
        self.single = single        # CK_F_SINGLE_COPY: the single-copy arena layout (batches of records <= 512 bases)

    @property
    def layout_flag(self) -> int:
        return N.CK_F_SINGLE_COPY if self.single else 0

    @property
    def lens(self) -> torch.Tensor:
        return self.offsets[1:] - self.offsets[:-1]


def synth_batch(ctx: Context, *, seed: int, first_index: int, n_records: int, kind: int, lo: int, hi: int,
                dup_permille: int = 0, adversarial_permille: int = 0, device=None, single: bool | None = None) -> DeviceBatch:
    """BASELINE.json synthetic records generated straight into the packed arena.  single: the single-copy layout
    (default: chosen like the host-buffer entries do -- batches whose records are all <= 512 bases)."""
    if single is None:
        single = hi <= 512
    lf = N.CK_F_SINGLE_COPY if single else 0
    dev = device or torch.device("cuda", ctx.device)
    offsets = torch.empty(n_records + 1, dtype=torch.int64, device=dev)
    total = C.c_uint64(0)
    ctx._check(ctx._lib.ck_synth_offsets(ctx.handle, _stream(), seed, first_index, n_records, kind, lo, hi,
                                         dup_permille, _p(offsets), C.byref(total)))
    words = int(ctx._lib.ck_packed2_words(total.value, n_records, lf))
    packed2 = torch.empty(words, dtype=torch.int64, device=dev)
    ctx._check(ctx._lib.ck_synth_packed2(ctx.handle, _stream(), seed, first_index, n_records, _p(offsets),
                                         dup_permille, adversarial_permille, _p(packed2), lf))
    return DeviceBatch(offsets, packed2, n_records, total.value, single)


def unpack_ascii(ctx: Context, b: DeviceBatch, n_records: int | None = None) -> torch.Tensor:
    """ASCII arena of the first n_records records (for the CPU baseline / parity checks)."""
    n = b.n if n_records is None else n_records
    total = int(b.offsets[n].item())
    out = torch.empty(max(total, 1), dtype=torch.uint8, device=b.offsets.device)
    ctx._check(ctx._lib.ck_dev_unpack2(ctx.handle, _stream(), _p(b.packed2), _p(b.offsets), n, _p(out), b.layout_flag))
    return out[:total]


class CanonOutputs:
    def __init__(self, n: int, total: int, dev, want_bytes=True, want_hash=True, aligned=False):
        self.aligned = aligned
        nbytes = 32 * ((total >> 5) + n) + 32 if aligned else max(total, 1) + 16
        self.out = torch.empty(nbytes, dtype=torch.uint8, device=dev) if want_bytes else None
        self.start = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        self.strand = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
        self.hash = torch.empty(max(n, 1), dtype=torch.int64, device=dev) if want_hash else None


class Workspace:
    def __init__(self, ctx: Context, n_records: int, total_bytes: int = 0, dev=None):
        self.bytes = int(ctx._lib.ck_dev_workspace_bytes(n_records, total_bytes))
        self.buf = torch.empty(self.bytes, dtype=torch.uint8, device=dev or torch.device("cuda", ctx.device))


def canon_packed2(ctx: Context, b: DeviceBatch, outs: CanonOutputs, ws: Workspace, class_mask: int = 0):
    """k_classify + LMSR/canonical/XXH3 kernels over a resident batch (no copies, no sync)."""
    flags = (N.CK_F_ALIGNED_OUT if outs.aligned else 0) | (0 if outs.out is not None else N.CK_F_NO_BYTES) | b.layout_flag
    ctx._check(ctx._lib.ck_dev_canon_packed2(ctx.handle, _stream(), _p(b.packed2), _p(b.offsets), b.n, flags, class_mask,
                                             _p(outs.out), _p(outs.start), _p(outs.strand), _p(outs.hash),
                                             _p(ws.buf), ws.bytes))


def canon_bytes(ctx: Context, raw: torch.Tensor, offsets: torch.Tensor, n: int, total: int, outs: CanonOutputs,
                lens_out: torch.Tensor, ws: Workspace, normalize: bool, class_mask: int = 0):
    """normalise + classify + pack + canonicalise a resident batch of raw record bytes (every symbol lane): the
    worker closure of src/canonicalize.rs:21-30 on HBM-resident input (no copies, no sync)."""
    flags = ((N.CK_F_NORMALIZE if normalize else 0) | (N.CK_F_ALIGNED_OUT if outs.aligned else 0)
             | (0 if outs.out is not None else N.CK_F_NO_BYTES))
    ctx._check(ctx._lib.ck_dev_canon_bytes(ctx.handle, _stream(), _p(raw), _p(offsets), n, total, flags, class_mask,
                                           _p(outs.out), _p(lens_out), _p(outs.start), _p(outs.strand), _p(outs.hash),
                                           _p(ws.buf), ws.bytes))


def check(ctx: Context, ws: Workspace):
    ctx._check(ctx._lib.ck_dev_check(ctx.handle, _stream(), _p(ws.buf)))


class DeviceTable:
    """First-occurrence table in HBM (src/uniq.rs:27,47-48 on device)."""

    def __init__(self, ctx: Context, capacity_keys: int, dev=None):
        self.ctx = ctx
        self.bytes = int(ctx._lib.ck_dev_table_bytes(capacity_keys))
        self.buf = torch.empty(self.bytes, dtype=torch.uint8, device=dev or torch.device("cuda", ctx.device))
        self.clear()

    def clear(self):
        self.ctx._check(self.ctx._lib.ck_dev_table_clear(self.ctx.handle, _stream(), _p(self.buf), self.bytes))

    def insert(self, hash64: torch.Tensor, n: int, slot_scratch: torch.Tensor, index: torch.Tensor | None = None,
               base_index: int = 0):
        self.ctx._check(self.ctx._lib.ck_dev_table_insert(self.ctx.handle, _stream(), _p(self.buf), self.bytes,
                                                          _p(hash64), _p(index), base_index, n, _p(slot_scratch)))

    def insert_pairs(self, pairs: torch.Tensor, n: int, slot_scratch: torch.Tensor):
        """insert (hash64, index) rows of an int64[n, 2] tensor (what the hash-range exchange delivers)"""
        self.ctx._check(self.ctx._lib.ck_dev_table_insert_pairs(self.ctx.handle, _stream(), _p(self.buf), self.bytes,
                                                                _p(pairs), n, _p(slot_scratch)))

    def first(self, slot_scratch: torch.Tensor, n: int, out_first: torch.Tensor):
        self.ctx._check(self.ctx._lib.ck_dev_table_first(self.ctx.handle, _stream(), _p(self.buf), self.bytes,
                                                         _p(slot_scratch), n, _p(out_first)))


class OwnerPartitioner:
    """Buckets local (hash64, global index) pairs by owning rank with two CUDA kernels (the send side of the
    hash-range exchange, exchange.exchange_first_index)."""

    def __init__(self, ctx: Context, n: int, world: int, dev=None):
        dev = dev or torch.device("cuda", ctx.device)
        self.ctx, self.world = ctx, world
        self.pairs = torch.empty((max(n, 1), 2), dtype=torch.int64, device=dev)
        self.pos = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        self.counts_dev = torch.zeros(2 * world, dtype=torch.int32, device=dev)
        self.counts_host = (C.c_uint32 * world)()
        self._padded, self._state = None, None

    def __call__(self, hash64: torch.Tensor, base_index: int, world: int):
        n = hash64.numel()
        assert world == self.world and n <= self.pairs.shape[0]
        self.ctx._check(self.ctx._lib.ck_dev_owner_partition(self.ctx.handle, _stream(), _p(hash64), n, base_index, world,
                                                             _p(self.pairs), _p(self.pos), _p(self.counts_dev),
                                                             self.counts_host))
        return self.pairs[:n], self.pos[:n], [int(c) for c in self.counts_host]

    def bucket_capacity(self, n: int) -> int:
        """entries per fixed-capacity bucket: the mean n / world plus 8 standard deviations of a uniform split"""
        return exchange_bucket_capacity(n, self.world)

    def padded(self, hash64: torch.Tensor, base_index: int, world: int):
        """fixed-capacity buckets for the equal-split exchange (no counts, no host synchronisation):
        -> (pairs int64[world * cap, 2], pos int32[n], state int32[world + 1]: per-owner counts, then the overflow flag)"""
        n = hash64.numel()
        cap = self.bucket_capacity(n)
        assert world == self.world and n <= self.pos.shape[0]
        if self._padded is None or self._padded.shape[0] != world * cap:
            self._padded = torch.empty((world * cap, 2), dtype=torch.int64, device=self.pos.device)
            self._state = torch.zeros(world + 1, dtype=torch.int32, device=self.pos.device)
        self.ctx._check(self.ctx._lib.ck_dev_owner_partition_padded(self.ctx.handle, _stream(), _p(hash64), n, base_index, world,
                                                                    cap, _p(self._padded), _p(self.pos), _p(self._state)))
        return self._padded, self.pos[:n], self._state


class PeerExchange:
    """Multi-GPU uniq's hash-range exchange fused into the kernels around it: the partition kernel stores every
    (hash, index) pair straight into its owner's receive buffer over NVLink, the owner's query kernel stores every answer
    straight into the asking rank's return buffer; two device-side barriers, no collective on the data path, no host
    synchronisation.  torch symmetric memory maps the buffers of all ranks into every process and provides the barrier.

    first_index(hash64, base_index, table, out) -> state int32[world + 1]; state[world] != 0: a bucket overflowed, `out`
    is not valid and the batch must be repeated through exchange.exchange_first_index (checked by the caller, once,
    after everything has been queued).  Every rank must call it for every batch (the barriers are collective), and every
    rank must construct it with the same n_max (the bucket capacity, hence the buffer layout, derives from it)."""

    def __init__(self, ctx: Context, n_max: int, world: int, rank: int, dev=None, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        dev = dev or torch.device("cuda", ctx.device)
        self.ctx, self.world, self.rank = ctx, world, rank
        self.cap = exchange_bucket_capacity(n_max, world)
        m = world * self.cap
        self.m = m
        self.recv = symm_mem.empty((m, 2), dtype=torch.int64, device=dev)
        self.ret = symm_mem.empty(m, dtype=torch.int64, device=dev)
        group = group or dist.group.WORLD
        self.h_recv = symm_mem.rendezvous(self.recv, group)
        self.h_ret = symm_mem.rendezvous(self.ret, group)
        self.recv_ptrs = (C.c_uint64 * world)(*[int(p) for p in self.h_recv.buffer_ptrs])
        self.ret_ptrs = (C.c_uint64 * world)(*[int(p) for p in self.h_ret.buffer_ptrs])
        self.pos = torch.empty(max(n_max, 1), dtype=torch.int32, device=dev)
        self.state = torch.zeros(world + 1, dtype=torch.int32, device=dev)
        self.slots = torch.empty(m, dtype=torch.int64, device=dev)

    def first_index(self, hash64: torch.Tensor, base_index: int, table: "DeviceTable", out: torch.Tensor) -> torch.Tensor:
        n = hash64.numel()
        assert n <= self.pos.shape[0] and out.numel() >= n
        lib, h = self.ctx._lib, self.ctx.handle
        self.ctx._check(lib.ck_dev_owner_scatter_peers(h, _stream(), _p(hash64), n, base_index, self.world, self.rank, self.cap,
                                                       self.recv_ptrs, _p(self.pos), _p(self.state)))
        self.h_recv.barrier()                       # every rank's pairs (and padding) have landed here
        table.insert_pairs(self.recv, self.m, self.slots)
        self.ctx._check(lib.ck_dev_table_first_peers(h, _stream(), _p(table.buf), table.bytes, _p(self.slots), self.world,
                                                     self.rank, self.cap, self.ret_ptrs))
        self.h_ret.barrier()                        # every owner's answers have landed here
        self.ctx._check(lib.ck_dev_gather_first(h, _stream(), _p(self.ret), _p(self.pos), n, _p(out)))
        return self.state


class PeerGroup:
    """The same fused exchange through the library's own peer group (ck_peer_export / ck_peer_attach, CUDA IPC): exact
    per-owner counts travel with the first barrier, so there is no padding, no overflow and no fallback, and nothing here
    needs torch beyond the process group that carries the 64-byte handles.  Every rank constructs it with the same n_max
    and calls first_index() for every batch (the barriers are collective)."""

    def __init__(self, ctx: Context, n_max: int, world: int, rank: int, group=None):
        import torch.distributed as dist
        self.ctx, self.world, self.rank = ctx, world, rank
        mine = ctx.peer_export(world, rank, n_max)
        handles = [None] * world
        dist.all_gather_object(handles, mine, group=group)
        ctx.peer_attach(handles)

    def first_index(self, hash64: torch.Tensor, base_index: int, table: "DeviceTable", out: torch.Tensor):
        n = hash64.numel()
        self.ctx._check(self.ctx._lib.ck_dev_peer_first_index(self.ctx.handle, _stream(), _p(hash64), n, base_index, _p(table.buf),
                                                              table.bytes, _p(out)))


def exchange_bucket_capacity(n: int, world: int) -> int:
    mean = n / world
    return int(mean + 8.0 * (mean ** 0.5) + 64) + 1


def kernel_times(ctx: Context):
    """{class name: (total ms, launches)} since the last call (needs ck_kernel_timing(1))."""
    n = len(CLASS_NAMES)
    ms = (C.c_double * n)()
    cnt = (C.c_uint32 * n)()
    ctx._check(ctx._lib.ck_kernel_times(ctx.handle, ms, cnt, n))
    return {CLASS_NAMES[i]: (ms[i], cnt[i]) for i in range(n)}


def algorithmic_bytes(lens: torch.Tensor, uniq: bool, lo: int = 1, hi: int = 1 << 62) -> int:
    """SURVEY §8d, 2-bit lane: 8*ceil(n/32) packed read + n ASCII write + 16 (offset read, start/strand write);
    uniq adds 8 (hash write) + 32 (one 16-byte table slot read + write)."""
    sel = lens[(lens >= lo) & (lens <= hi)]
    b = (8 * ((sel + 31) // 32) + sel + 16).sum()
    if uniq:
        b = b + 40 * sel.numel()
    return int(b.item())
