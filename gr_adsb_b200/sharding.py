"""Sharding one long sample stream across GPUs.

The stream is cut into `world` contiguous shards; shard r owns the pulses whose falling edge lies at
global window index [r*L, (r+1)*L)  (window index = stream index + 8*sps-1, the GNU Radio history of
the single whole-buffer work() call this reproduces).  Detection and demodulation only look at a
bounded neighbourhood, so each rank needs

  left halo   256 window items  (>= 100-sample noise window + pulse-run look-back, framer.py:156-159)
  right halo  120*sps items     (preamble + 112 bits,                          demod.py:75-76)

from its neighbours -- a few KiB exchanged with NCCL send/recv -- and one sequential quantity: the
blanking state `prev_eob_idx` (framer.py:121-123,165) entering the shard.  Every rank first resolves
its candidates assuming none, then the ranks all-gather (first candidate, last hit) and re-run the
two tiny resolve kernels wherever the true incoming state reaches their first candidate.
"""
LEFT_HALO = 256


# ---- peer mode (the product path on NVLink machines): adsb_shard_* in include/adsb_b200.h ------------------------
# Nothing is copied between GPUs.  Every rank keeps exactly its own `shard_items` samples in a symmetric allocation
# that the other ranks map through CUDA IPC; the scan kernel's edge tiles read the neighbours' samples in place and
# the blanking carry travels through peer-written slots.  torch.distributed is only the plumbing that hands the
# 64-byte IPC handles around once at set-up.
def peer_attach(ctx, dist, rank, world, sample_bytes):
    """Allocate this rank's symmetric shard, exchange the IPC handles over `dist` (any backend) and map the peers.
    Returns the device pointer of the sample area."""
    ptr, handle = ctx.shard_alloc(sample_bytes)
    if world == 1:
        ctx.shard_attach_ptrs(0, 1, [ptr])
        return ptr
    handles = [None] * world
    dist.all_gather_object(handles, handle)
    ctx.shard_attach(rank, world, handles)
    return ptr


def peer_window(rank, world, shard_items, sps):
    """Geometry of adsb_shard_step (mirrors shard_issue in csrc/adsb_api.cu): the GLOBAL window indices [own_lo, own_hi)
    of the pulse falls rank `rank` owns; global window index = stream index + 8*sps - 1."""
    h1 = 8 * sps - 1
    lo = 0 if rank == 0 else rank * shard_items + h1
    hi = (rank + 1) * shard_items + (0 if rank == world - 1 else h1)
    return lo, hi



def shard_geometry(rank, world, shard_items, sps):
    """Arguments of adsb_burst_work_dev for shard `rank` holding `shard_items` of the stream."""
    h1 = 8 * sps - 1
    first, last = rank == 0, rank == world - 1
    left = 0 if first else LEFT_HALO + h1          # stream items borrowed from the left neighbour
    right = 0 if last else 120 * sps - h1          # ... and from the right neighbour
    lead = h1 if first else 0
    halo_left = 0 if first else LEFT_HALO
    n_in = left + shard_items + right
    n_out = halo_left + shard_items
    base = 0 if first else rank * shard_items - LEFT_HALO     # local window index 0 -> global window index
    assert lead + n_in >= n_out + h1
    return {"rank": rank, "world": world, "left": left, "right": right, "lead": lead, "halo_left": halo_left,
            "n_in": n_in, "n_out": n_out, "base": base, "shard_items": shard_items, "blank": 63 * sps}


def exchange_halos(buf, geo, dist):
    """buf: [n_in, ...] device tensor, own shard at [left, left+shard_items).  Fills both halos from the
    neighbours with batched NCCL (or gloo) point-to-point transfers."""
    r, w, left, n = geo["rank"], geo["world"], geo["left"], geo["shard_items"]
    if w == 1:
        return
    ops = []
    h1_left = None
    if r + 1 < w:
        # my tail is the right neighbour's left halo; its head is my right halo
        nxt = shard_geometry(r + 1, w, n, (geo["blank"] // 63))
        ops.append(dist.P2POp(dist.isend, buf[left + n - nxt["left"]: left + n], r + 1))
        ops.append(dist.P2POp(dist.irecv, buf[left + n: left + n + geo["right"]], r + 1))
    if r > 0:
        prv = shard_geometry(r - 1, w, n, (geo["blank"] // 63))
        ops.append(dist.P2POp(dist.isend, buf[left: left + prv["right"]], r - 1))
        ops.append(dist.P2POp(dist.irecv, buf[0:left], r - 1))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    return h1_left


NONE = -(1 << 62)


def carry_plan(infos, blank):
    """Pure host logic of the blanking carry.  infos[r] = (first_g, last_g, used_g) in GLOBAL window
    coordinates: first candidate (NONE if the shard has none), last hit of the shard's latest resolve
    (NONE if none), and the incoming state that resolve assumed (NONE = nothing).  Returns
    (e_in[r] the true incoming state per rank given what is known, need[r] whether rank r must
    re-resolve).  Ranks apply their fix-ups and call again until no rank needs one."""
    e_in, need = [], []
    e = NONE
    for first_g, last_g, used_g in infos:
        e_in.append(e)
        changed = e != used_g
        reach = first_g != NONE and first_g <= max(e, used_g)
        need.append(bool(changed and reach))
        e_out = (last_g + blank) if last_g != NONE else e
        # a stale last hit cannot exceed the true one's successor chain start, but if this rank is about
        # to re-resolve its last hit may move: later ranks are re-planned on the next round
        e = e_out
    return e_in, need
