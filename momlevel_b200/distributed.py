"""Sharding the steric path over the GPUs of one box (one process per GPU).

Every water column and every time step is independent given the reference state
(src/momlevel/steric.py:150-166 has no horizontal coupling), so the time axis -- or the
member axis of an ensemble -- is cut into contiguous blocks, one per rank, with no halo and no
collective on the data path.  The only exchange is the gather of the per-step global mass
series ``M(t)`` (<= a few hundred doubles) before the ``ln`` formula of steric.py:136-142.

``torch.distributed`` is the plumbing: NCCL over NVLink on GPUs, gloo in the CPU tests.
"""

import numpy as np
import torch
import torch.distributed as dist

__all__ = ["shard_range", "shard_sizes", "assign_members", "assign_member_blocks", "gather_series",
           "global_sea_level", "steric_global_sharded", "finish_global_series", "steric_local_members", "steric_local_pieces",
           "bind_host_to_device", "local_world_size"]


def bind_host_to_device(device_index):
    """Pin this process to the CPU cores next to its GPU (NVML's ideal affinity), before host buffers exist.

    With one process per GPU the pinned staging buffers of ``ml_steric_local_host`` should live on
    the NUMA node the GPU's PCIe root hangs off: pages are placed where the allocating thread runs,
    and a copy that crosses the socket interconnect shares that link with every other rank.
    Returns the number of cores bound to, or 0 when NVML or the affinity call is unavailable.
    """
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(device_index)
        busid = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(busid.encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:  # noqa: BLE001 -- no NVML, masked device, restricted cpuset: stay unbound
        return 0


def shard_sizes(n, world, light_first=False):
    """Sizes of ``world`` contiguous blocks covering ``n`` items; the first ``n % world`` get one more -- the LAST
    ones with ``light_first``: the rank that owns step 0 of a global series also sums the reference volume
    (reference.py:74-77), so it is the one to take a short block when the steps do not divide evenly."""
    base, extra = divmod(int(n), int(world))
    sizes = [base + (1 if r < extra else 0) for r in range(world)]
    return sizes[::-1] if light_first else sizes


def shard_range(n, world, rank, light_first=False):
    """``(start, stop)`` of this rank's block, e.g. 365 steps on 8 ranks -> 46,46,46,46,46,45,45,45
    (45,45,45,46,46,46,46,46 with ``light_first``)."""
    sizes = shard_sizes(n, world, light_first)
    start = sum(sizes[:rank])
    return start, start + sizes[rank]


def assign_members(n_members, world, rank):
    """Ensemble members owned by ``rank`` (contiguous blocks; 30 members on 8 ranks -> 4,4,4,4,4,4,3,3)."""
    start, stop = shard_range(n_members, world, rank)
    return list(range(start, stop))


def assign_member_blocks(n_members, nt, world, rank, block=12):
    """This rank's share of an ensemble cut along member AND time: ``[(member, t_start, t_stop), ...]``.

    Whole members per rank leave 30 members on 8 ranks at 4,4,4,4,4,4,3,3 -- the two light ranks idle for a
    quarter of the job.  A member's time axis shards as freely as the members do once the rank has that
    member's reference state (its step 0, one extra step to load: steric.py:105-107 takes it from
    ``dset.isel(time=0)``), so the flattened list of (member, ``block``-step block) pairs is cut into
    contiguous shares instead: 30 x 120 months on 8 ranks -> 38,38,38,38,37,37,37,37 blocks of 12 steps.
    Consecutive blocks of one member are merged into one piece.
    """
    per = -(-int(nt) // int(block))
    lo, hi = shard_range(int(n_members) * per, world, rank)
    pieces = []
    i = lo
    while i < hi:
        m, b = divmod(i, per)
        b_hi = min(per, b + (hi - i))
        pieces.append((m, b * block, min(int(nt), b_hi * block)))
        i += b_hi - b
    return pieces


def gather_series(local, n_total, group=None, extra=None, light_first=False):
    """All-gather the ranks' contiguous shards of a 1-D series into the full series on every rank.

    ``local`` is this rank's block (length ``shard_sizes(n_total, world)[rank]``), a tensor on
    the device the process group communicates with.  Blocks are padded to a common length so
    a single ``all_gather_into_tensor`` moves them.  ``extra`` (a small 1-D tensor of the same
    length on every rank, or ``None`` everywhere) rides in the same message: the scalars of the
    reference state, which only the rank that owns step 0 has; returns ``(series, extras[world, k])`` then.
    ``light_first``: the blocks were cut with ``shard_range(..., light_first=True)``.
    """
    if not dist.is_available() or not dist.is_initialized():
        assert local.numel() == n_total
        return local.clone() if extra is None else (local.clone(), extra.clone().view(1, -1))
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(n_total, world, light_first)
    assert local.numel() == sizes[rank], f"rank {rank} holds {local.numel()} values, expected {sizes[rank]}"
    width = max(sizes)
    k = 0 if extra is None else int(extra.numel())
    send = torch.zeros(width + k, dtype=local.dtype, device=local.device)
    send[: local.numel()] = local
    if k:
        send[width:] = extra.to(send.dtype)
    recv = torch.empty(world * (width + k), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.view(world, width + k)
    series = torch.cat([recv[r, : sizes[r]] for r in range(world)])
    return series if extra is None else (series, recv[:, width:].clone())


def global_sea_level(masso, volo, rhoga, area_sum):
    """steric.py:136-142: ``(volo / sum(areacello)) * ln(rhoga / (masso / volo))`` on the host."""
    masso = np.asarray(masso, dtype=np.float64)
    reference_height = np.float64(volo) / np.float64(area_sum)
    return reference_height * np.log(np.float64(rhoga) / (masso / np.float64(volo))), reference_height


def steric_global_sharded(T_local, S_local, v_ref, p_level, volo, rhoga, area_sum, n_total, eos="Wright", group=None,
                          masso_local=None, ref_sums=None, light_first=False):
    """Global steric series with the time axis sharded over ranks.

    Each rank passes its own contiguous block of time steps ``T_local, S_local`` (resident on
    its GPU) plus the shared reference volume -- or, when its block was streamed through windows,
    the per-step masses ``masso_local`` it has already computed; returns
    ``(eta[n_total], reference_height)`` on every rank.

    ``volo, rhoga`` are the scalars of the reference state (reference.py:74-80).  Only the rank whose
    block holds step 0 can compute them without loading that step again; it passes
    ``ref_sums = tensor([volo, masso_ref])`` (what ``core.reference_state`` returns) and every other rank
    ``volo = rhoga = None``: the two numbers then travel in the one all-gather of the series.
    """
    from . import core

    if masso_local is None:
        masso_local = core.steric_global(T_local, S_local, v_ref, p_level, eos=eos)
    if volo is not None and rhoga is not None:
        masso = gather_series(masso_local, n_total, group=group, light_first=light_first)
        return global_sea_level(masso.cpu().numpy(), volo, rhoga, area_sum)
    mine = ref_sums if ref_sums is not None else torch.zeros(2, dtype=masso_local.dtype, device=masso_local.device)
    masso, extras = gather_series(masso_local, n_total, group=group, extra=mine.to(masso_local.device),
                                  light_first=light_first)
    return finish_global_series(masso, extras, area_sum)


def finish_global_series(masso, extras, area_sum):
    """The host end of :func:`steric_global_sharded`: one read-back of the gathered series and the scalars that rode
    with it, then the ``ln`` formula (steric.py:136-142).  ``extras`` is ``[world, 2]``: ``{volo, masso_ref}`` from the
    rank that owns step 0, zeros from the others."""
    n_total = masso.numel()
    host = torch.cat([masso, extras.reshape(-1)]).cpu().numpy()  # one read-back
    masso_h, extras_h = host[:n_total], host[n_total:].reshape(-1, 2)
    owner = int(np.argmax(extras_h[:, 0] != 0.0))  # the rank that holds step 0 (volo > 0)
    volo, masso_ref = float(extras_h[owner, 0]), float(extras_h[owner, 1])
    return global_sea_level(masso_h, volo, masso_ref / volo, area_sum)


_STREAMS = {}  # device index -> side streams, kept so that the allocator's per-stream pools stay warm


def _member_streams(dev, n):
    pool = _STREAMS.setdefault(dev.index if dev.index is not None else torch.cuda.current_device(), [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=dev))
    return pool[:n]


def steric_local_members(members, z_i, deptho, p_level, rhozero=1035.0, eos="Wright", n_streams=4):
    """Local steric height of this rank's ensemble members, one ``steric()`` problem each.

    The reference API is strictly 4-D (``util.py:762-770``), so an ensemble is a loop over members;
    every member carries its own reference state (its step 0).  ``members`` is a sequence of
    ``(T, S, v_ref)`` device tensors.  Members are issued round-robin on ``n_streams`` CUDA streams:
    a member of a 1-degree grid is only a few waves of CTAs per launch, and with several streams the
    tail of one member's launch is filled by the head of the next member's.  Results are ordered
    after the calling stream.  Returns a list of ``(eta, rho_ref, sums)``.
    """
    from . import core

    members = list(members)
    if not members:
        return []
    dev = members[0][0].device
    caller = torch.cuda.current_stream(dev)
    streams = _member_streams(dev, max(1, min(int(n_streams), len(members))))
    # outputs live on the caller's stream: allocated before the fork, used after the join
    outs = [core.selfref_outputs(T, S) for T, S, _ in members]
    fork = torch.cuda.Event()
    fork.record(caller)
    for i, (T, S, V) in enumerate(members):
        st = streams[i % len(streams)]
        if i < len(streams):
            st.wait_event(fork)
        with torch.cuda.stream(st):  # the reduction scratch is allocated (and recycled) on this stream
            core.steric_local_selfref(T, S, V, z_i, deptho, p_level, rhozero=rhozero, eos=eos, out=outs[i])
    for st in streams:
        join = torch.cuda.Event()
        join.record(st)
        caller.wait_event(join)
    return outs


def steric_local_pieces(pieces, z_i, deptho, p_level, rhozero=1035.0, eos="Wright", n_streams=4, outs=None):
    """Local steric height of this rank's share of an ensemble cut by :func:`assign_member_blocks`.

    ``pieces`` is a sequence of ``(T, S, v_ref, ref)``: the time block of one member resident on this GPU,
    the member's reference volume, and ``ref = None`` when the block starts at the member's step 0 (the
    fused self-reference pass) or ``ref = (T0, S0)``, the member's step-0 slabs, when it does not: the
    reference density is then evaluated from them first (reference.py:60-71) and the block integrated
    against it.  Issued round-robin on ``n_streams`` streams like :func:`steric_local_members`.
    Returns a list of ``(eta, rho_ref, sums)``; ``outs`` may hand those tensors in (``core.selfref_outputs``).
    """
    from . import core

    pieces = list(pieces)
    if not pieces:
        return []
    dev = pieces[0][0].device
    caller = torch.cuda.current_stream(dev)
    streams = _member_streams(dev, max(1, min(int(n_streams), len(pieces))))
    if outs is None:
        outs = [core.selfref_outputs(T, S) for T, S, _, _ in pieces]  # on the caller's stream, before the fork
    fork = torch.cuda.Event()
    fork.record(caller)
    for i, (T, S, V, ref) in enumerate(pieces):
        st = streams[i % len(streams)]
        if i < len(streams):
            st.wait_event(fork)
        with torch.cuda.stream(st):
            if ref is None:
                core.steric_local_selfref(T, S, V, z_i, deptho, p_level, rhozero=rhozero, eos=eos, out=outs[i])
            else:
                eta, rho, sums = outs[i]
                core.reference_state(ref[0], ref[1], V, p_level, eos=eos, out=(rho, sums))
                core.steric_local(T, S, rho, V, z_i, deptho, p_level, rhozero=rhozero, eos=eos, eta_out=eta)
    for st in streams:
        join = torch.cuda.Event()
        join.record(st)
        caller.wait_event(join)
    return outs


def local_world_size():
    """Ranks sharing this host (``LOCAL_WORLD_SIZE`` under torchrun; the world size of a one-node job otherwise)."""
    import os

    for key in ("LOCAL_WORLD_SIZE", "WORLD_SIZE"):
        v = os.environ.get(key)
        if v and v.isdigit() and int(v) > 0:
            return int(v)
    return 1
