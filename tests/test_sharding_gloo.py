"""Host-side multi-rank logic on CPU: two gloo processes exchange shard halos with the same code the
NCCL path uses and agree on the blanking-carry plan."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gr_adsb_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, L, sps, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        geo = sharding.shard_geometry(rank, world, L, sps)
        buf = torch.full((geo["n_in"], 2), -1.0)
        # own shard = global stream index in both components
        own = torch.arange(rank * L, (rank + 1) * L, dtype=torch.float32)
        buf[geo["left"]: geo["left"] + L, 0] = own
        buf[geo["left"]: geo["left"] + L, 1] = -own
        sharding.exchange_halos(buf, geo, dist)
        first = rank * L - geo["left"]
        expect = torch.arange(first, first + geo["n_in"], dtype=torch.float32)
        ok_halo = bool(torch.equal(buf[:, 0], expect) and torch.equal(buf[:, 1], -expect))
        # carry plan: every rank contributes (first, last, used) and all must derive the same plan
        info = torch.tensor([rank * L + 50, rank * L + L - 40, sharding.NONE], dtype=torch.int64)
        got = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(got, info)
        infos = [tuple(int(v) for v in t.tolist()) for t in got]
        e_in, need = sharding.carry_plan(infos, geo["blank"])
        out.put((rank, ok_halo, e_in, need))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("sps", [2, 8])
def test_two_rank_halo_exchange_and_carry(sps):
    world, L = 2, 4096
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, L, sps, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)
    assert res[0][2] == res[1][2] and res[0][3] == res[1][3]
    blank = 63 * sps
    # rank 1's first candidate (L+50) lies within 63 symbols of rank 0's last hit (L-40) -> must re-resolve
    assert res[0][2] == [sharding.NONE, L - 40 + blank]
    assert res[0][3] == [False, (L + 50) <= (L - 40 + blank)]


def test_geometry_tiles_the_stream():
    for sps in (2, 4, 8):
        for world in (1, 2, 4, 8):
            L = 1 << 16
            prev_end = 0
            for r in range(world):
                geo = sharding.shard_geometry(r, world, L, sps)
                own_lo = geo["base"] + geo["halo_left"]
                own_hi = geo["base"] + geo["n_out"]
                assert own_lo == prev_end                      # owned window ranges tile [0, world*L)
                prev_end = own_hi
                assert geo["lead"] + geo["n_in"] >= geo["n_out"] + 8 * sps - 1
                if r + 1 < world:                               # right halo long enough to demodulate
                    assert geo["lead"] + geo["n_in"] - geo["n_out"] >= 120 * sps
            assert prev_end == world * L


def test_carry_plan_cascade():
    N = sharding.NONE
    blank = 126
    # rank 1 fully blanked -> passes the state through to rank 2
    infos = [(10, 1000, N), (1050, 1050, N), (1100, 1500, N)]
    e_in, need = sharding.carry_plan(infos, blank)
    assert e_in == [N, 1126, 1176] and need == [False, True, True]
    # after rank 1 re-resolved (no hits now) its state passes through: rank 2 sees 1126
    infos = [(10, 1000, N), (1050, N, 1126), (1100, 1500, 1176)]
    e_in, need = sharding.carry_plan(infos, blank)
    assert e_in == [N, 1126, 1126] and need == [False, False, True]
    infos = [(10, 1000, N), (1050, N, 1126), (1100, 1500, 1126)]
    e_in, need = sharding.carry_plan(infos, blank)
    assert need == [False, False, False]


# ---- peer mode (adsb_shard_*): host-side plumbing under gloo ------------------------------------------------------
class _FakeCtx(object):
    """Stands in for gr_adsb_b200.Context where there is no GPU: records what peer_attach asks of it."""

    def __init__(self, rank):
        self.rank = rank
        self.attached = None

    def shard_alloc(self, nbytes):
        return 0x10000 * (self.rank + 1), bytes([self.rank]) * 64          # (sample pointer, 64-byte IPC handle)

    def shard_attach(self, rank, world, handles):
        self.attached = (rank, world, list(handles))

    def shard_attach_ptrs(self, rank, world, ptrs):
        self.attached = (rank, world, list(ptrs))


def _peer_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ctx = _FakeCtx(rank)
        ptr = sharding.peer_attach(ctx, dist, rank, world, 1 << 20)
        out.put((rank, ptr, ctx.attached))
    finally:
        dist.destroy_process_group()


def test_two_rank_peer_attach_exchanges_every_handle():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ptr, attached in res:
        assert ptr == 0x10000 * (rank + 1)
        assert attached == (rank, world, [bytes([0]) * 64, bytes([1]) * 64])      # every rank's handle, rank order


def test_peer_windows_tile_the_whole_buffer_call():
    """adsb_shard_step's ownership: the ranks' fall ranges tile [0, n_out) of the single whole-stream call exactly
    (n_out = stream length; the last 8*sps-1 items are look-ahead only, framer.py:72-77)."""
    for sps in (2, 4, 8):
        for world in (1, 2, 3, 8):
            L = 1 << 16
            prev = 0
            for r in range(world):
                lo, hi = sharding.peer_window(r, world, L, sps)
                assert lo == prev and hi > lo
                prev = hi
            assert prev == world * L                    # == n_out of one work() call over world*L items of stream
