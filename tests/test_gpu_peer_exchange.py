"""The exchange fused into the partition / query kernels (device.PeerExchange: ck_dev_owner_scatter_peers,
ck_dev_table_first_peers, ck_dev_gather_first) in a one-rank NCCL group: the peer is the rank itself, so every store goes
through the same peer-pointer path as on an NVLink node.  The multi-rank agreement (exact == padded == peer on every
rank) is checked by tools/exchange_phases.py under torchrun; its 2- and 8-GPU logs are in profiles/."""
import socket

import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.fixture(scope="module")
def env():
    import torch
    import torch.distributed as dist
    import circkit_b200
    from circkit_b200 import device as D
    torch.cuda.set_device(0)
    created = False
    if not dist.is_initialized():
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % _free_port(), world_size=1, rank=0,
                                device_id=torch.device("cuda", 0))
        created = True
    ctx = circkit_b200.Context(max_batch_bytes=0, max_batch_records=0)
    yield ctx, D, torch
    ctx.close()
    if created:
        dist.destroy_process_group()


def test_peer_exchange_one_rank_matches_table(env):
    ctx, D, torch = env
    n, base = 400_000, 9_000_000_000
    g = torch.Generator(device="cuda").manual_seed(3)
    h = torch.randint(-2**63, 2**63 - 1, (n,), dtype=torch.int64, device="cuda", generator=g)
    h[7] = -1                                                                 # the key that equals the table's empty marker
    h[n // 2:] = h[: n // 2].clone()                                          # every key twice
    peer = D.PeerExchange(ctx, n, 1, 0)
    table = D.DeviceTable(ctx, n)
    out = torch.empty(n, dtype=torch.int64, device="cuda")
    state = peer.first_index(h, base, table, out)
    assert state.tolist() == [n, 0]
    ref_t = D.DeviceTable(ctx, n)
    slots = torch.empty(n, dtype=torch.int64, device="cuda"); want = torch.empty_like(slots)
    ref_t.insert(h, n, slots, base_index=base); ref_t.first(slots, n, want)
    assert torch.equal(out, want)
    assert torch.equal(out[n // 2:], torch.arange(base, base + n // 2, device="cuda"))
    # a second batch through the same buffers (the barriers order reuse), smaller than the first
    table.clear()
    state = peer.first_index(h[: n // 3], base, table, out)
    assert state.tolist() == [n // 3, 0]
    assert torch.equal(out[: n // 3], torch.arange(base, base + n // 3, device="cuda"))


def test_peer_exchange_overflow_is_flagged(env):
    ctx, D, torch = env
    n = 100_000
    peer = D.PeerExchange(ctx, n // 4, 1, 0)           # buckets sized for a quarter of what arrives
    peer.pos = torch.empty(n, dtype=torch.int32, device="cuda")
    table = D.DeviceTable(ctx, n)
    out = torch.empty(n, dtype=torch.int64, device="cuda")
    h = torch.arange(n, dtype=torch.int64, device="cuda")
    state = peer.first_index(h, 0, table, out)
    assert int(state[1]) == 1                           # flagged; nothing was written past the bucket
    assert int((out == -1).sum()) == n - peer.cap
