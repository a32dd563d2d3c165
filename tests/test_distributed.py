"""World-size-2 gloo tests (CPU) of the multi-GPU host logic, plus a GPU test of the sharded path.

The kernels cannot run here, so the CPU tests feed the gather with per-rank mass series
computed by the oracle and check that the sharded result equals the unsharded one.
"""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from momlevel_b200 import distributed as mld
from oracle import steric as osteric
from oracle import testdata


def test_shard_sizes():
    assert mld.shard_sizes(365, 8) == [46, 46, 46, 46, 46, 45, 45, 45]
    assert mld.shard_sizes(30, 8) == [4, 4, 4, 4, 4, 4, 3, 3]  # config 3
    assert mld.shard_sizes(3, 8) == [1, 1, 1, 0, 0, 0, 0, 0]
    assert mld.shard_sizes(365, 8, light_first=True) == [45, 45, 45, 46, 46, 46, 46, 46]  # rank 0 also sums the volume
    assert mld.shard_range(365, 8, 0, light_first=True) == (0, 45) and mld.shard_range(365, 8, 7, light_first=True) == (319, 365)
    for n, w in [(365, 8), (12, 5), (7, 7), (1, 4), (0, 3)]:
        blocks = [mld.shard_range(n, w, r) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
    assert mld.assign_members(30, 8, 6) == [24, 25, 26]


def test_member_blocks_balance_and_cover():
    """Config 3 cut by (member, 12-step block): contiguous shares that differ by at most one block, cover every block
    once, and keep the blocks of one member on a rank together (whole members would leave 4,4,4,4,4,4,3,3)."""
    for n_members, nt, world in [(30, 120, 8), (30, 120, 4), (30, 120, 1), (7, 25, 3), (2, 5, 5)]:
        per = -(-nt // 12)
        seen = []
        loads = []
        for r in range(world):
            pieces = mld.assign_member_blocks(n_members, nt, world, r)
            loads.append(sum(-(-(t1 - t0) // 12) for _, t0, t1 in pieces))
            for m, t0, t1 in pieces:
                assert 0 <= t0 < t1 <= nt and t0 % 12 == 0
                seen += [(m, b) for b in range(t0 // 12, -(-t1 // 12))]
        assert sorted(seen) == [(m, b) for m in range(n_members) for b in range(per)]
        assert max(loads) - min(loads) <= 1
    assert [sum(-(-(t1 - t0) // 12) for _, t0, t1 in mld.assign_member_blocks(30, 120, 8, r)) for r in range(8)] == \
        [38, 38, 38, 38, 37, 37, 37, 37]


def test_gather_series_without_process_group():
    x = torch.arange(5, dtype=torch.float64)
    assert torch.equal(mld.gather_series(x, 5), x)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nt, out_dir, light_first=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d = testdata.generate_test_data(ntimes=nt)
        ref = osteric.reference_state(d["thetao"], d["so"], d["volcello"], d["areacello"], d["z_l"])
        lo, hi = mld.shard_range(nt, world, rank, light_first)
        # this rank's time block only (what its GPU would hold); the reference state is step 0,
        # which every rank regenerates itself instead of receiving it
        _, _, masso_local = osteric.steric_global(d["thetao"][lo:hi], d["so"][lo:hi], d["z_l"], ref)
        masso = mld.gather_series(torch.from_numpy(np.ascontiguousarray(masso_local)), nt, light_first=light_first)
        eta, href = mld.global_sea_level(masso.numpy(), ref["volo"], ref["rhoga"], np.nansum(ref["areacello"]))
        # the same with the scalars of the reference state riding in the gather: only the rank that owns step 0 has them
        sums = torch.tensor([ref["volo"], ref["masso"]], dtype=torch.float64) if lo == 0 else None
        eta2, href2 = mld.steric_global_sharded(None, None, None, None, None, None, np.nansum(ref["areacello"]), nt,
                                                masso_local=torch.from_numpy(np.ascontiguousarray(masso_local)),
                                                ref_sums=sums, light_first=light_first)
        assert np.array_equal(eta2, eta) and href2 == href
        np.save(os.path.join(out_dir, f"eta_{rank}.npy"), eta)
        np.save(os.path.join(out_dir, f"href_{rank}.npy"), href)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("nt,light_first", [(5, False), (12, False), (5, True)])
def test_time_sharded_global_series_gloo(tmp_path, nt, light_first):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), nt, str(tmp_path), light_first), nprocs=world, join=True)
    d = testdata.generate_test_data(ntimes=nt)
    ref = osteric.reference_state(d["thetao"], d["so"], d["volcello"], d["areacello"], d["z_l"])
    want, href, _ = osteric.steric_global(d["thetao"], d["so"], d["z_l"], ref)
    for rank in range(world):
        got = np.load(tmp_path / f"eta_{rank}.npy")
        assert got.shape == (nt,)
        np.testing.assert_array_equal(got, want)  # every rank holds the full, identical series
        assert float(np.load(tmp_path / f"href_{rank}.npy")) == href


@pytest.mark.gpu
def test_sharded_global_matches_unsharded_on_gpu():
    """Emulate 3 ranks on one GPU: per-block kernels + the same assembly as the gather."""
    import momlevel_b200 as ml
    from momlevel_b200 import core, synth

    ds = synth.make_dataset(7, 12, 16, 64, seed=9, device="cuda", dtype=torch.float32)
    res, reference = ml.steric(ds, domain="global")
    pres = ds["z_l"].values * 1e4 + 101325.0
    V = reference["volcello"].data
    parts = []
    for r in range(3):
        lo, hi = mld.shard_range(7, 3, r)
        parts.append(core.steric_global(ds["thetao"].data[lo:hi].contiguous(), ds["so"].data[lo:hi].contiguous(), V, pres))
    masso = torch.cat(parts).cpu().numpy()
    eta, href = mld.global_sea_level(masso, float(reference["volo"]), float(reference["rhoga"]),
                                     float(reference["areacello"].sum()))
    assert np.allclose(eta, res["steric"].values, rtol=0, atol=1e-12)
    assert href == pytest.approx(float(res["reference_height"]), rel=1e-15)


@pytest.mark.gpu
@pytest.mark.parametrize("n_streams", [1, 2, 3])
def test_members_on_streams_equal_sequential_calls(n_streams):
    """Config 3's member loop: issuing members on several CUDA streams changes nothing but the timing."""
    from momlevel_b200 import core, synth

    shape = (25, 10, 16, 64)  # three 12-step chunks (one short) per member: self-reference + local launches
    grid = synth.make_grid(*shape[1:], seed=4, device="cuda")
    pres = grid["z_l"] * 1.0e4 + 101325.0
    fields = [synth.make_fields(grid, shape[0], seed=100 + m, dtype=torch.float32) for m in range(5)]
    want = [core.steric_local_selfref(T, S, V, grid["z_i"], grid["deptho"], pres) for T, S, V in fields]
    torch.cuda.synchronize()
    for _ in range(3):  # repeated: a missing stream dependency would show as a flaky mismatch
        got = mld.steric_local_members(fields, grid["z_i"], grid["deptho"], pres, n_streams=n_streams)
        for (e1, r1, s1), (e2, r2, s2) in zip(got, want):
            assert torch.equal(torch.nan_to_num(e1, nan=-1.0), torch.nan_to_num(e2, nan=-1.0))
            assert torch.equal(torch.nan_to_num(r1, nan=-1.0), torch.nan_to_num(r2, nan=-1.0))
            assert torch.equal(s1, s2)
