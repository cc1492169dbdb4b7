"""Host-side multi-rank logic on CPU: world_size-2 gloo processes (no GPU needed)."""

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, nf, ret):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from oracle import erfa_np as E
    from oracle import skymodel_np as S
    from pyradiosky_b200 import dist as rdist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(123)  # same catalogue on every rank
        lon = rng.uniform(0, 2 * np.pi, n)
        lat = np.arcsin(rng.uniform(-1, 1, n))
        st = np.zeros((4, nf, n))
        st[0] = rng.uniform(1, 2, (nf, n))
        st[1] = 0.1 * rng.normal(size=(nf, n))
        kwargs = dict(lon=lon, lat=lat, stokes=st, spectral_type="full", freq_array=np.linspace(1e8, 2e8, nf))
        mine = rdist.shard_kwargs(kwargs, world, rank)
        lo, hi = rdist.shard_bounds(n, world, rank)
        assert mine["lon"].shape[0] == hi - lo and mine["stokes"].shape == (4, nf, hi - lo)
        lat_s, lon_s = -np.deg2rad(30.72), np.deg2rad(21.43)
        astrom = E.astrom_context(2457082.5, 0.1, lon_s, lat_s, 1073.0)
        # the oracle stands in for the kernels: this test is about the sharding / exchange logic only
        alt_az, _, above = S.update_positions(mine["lon"], mine["lat"], astrom)
        ravg = S.calc_average_rotation_matrix(astrom)
        local = S.coherency_calc(mine["stokes"], mine["lon"], mine["lat"], alt_az, above, ravg, vectorized=True)
        counts, col0, total = rdist.exchange_counts(local.shape[-1])
        assert len(counts) == world and counts[rank] == local.shape[-1] and col0 == sum(counts[:rank])
        full = rdist.gather_columns(torch.as_tensor(local), counts)
        assert tuple(full.shape) == (2, 2, nf, total)
        # reference: the unsharded computation
        alt_az, _, above = S.update_positions(lon, lat, astrom)
        ref = S.coherency_calc(st, lon, lat, alt_az, above, ravg, vectorized=True)
        ok = full.shape == ref.shape and np.array_equal(full.numpy(), ref)
        ok = ok and np.array_equal(full.numpy()[..., col0: col0 + counts[rank]], local)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [1001, 2])
def test_sharded_equals_unsharded_gloo(n):
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), n, 3, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}


def test_shard_bounds():
    from pyradiosky_b200.dist import shard_bounds, shard_kwargs

    for n in (0, 1, 7, 8, 1000, 1_000_000, 12 * 1024 * 1024):
        for world in (1, 2, 3, 8):
            blocks = [shard_bounds(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            assert max(hi - lo for lo, hi in blocks) == -(-n // world) if n else True
    assert shard_bounds(1_000_000, 8, 3) == (375_000, 500_000)  # BASELINE config 4: 8 x 125,000
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)
    kw = dict(nside=4, stokes=np.zeros((4, 1, 192)), spectral_type="flat")
    assert shard_kwargs(kw, 4, 2)["hpx_first"] == 96 and shard_kwargs(kw, 4, 2)["stokes"].shape[-1] == 48


def test_exchange_counts_without_process_group():
    from pyradiosky_b200.dist import exchange_counts, gather_columns

    assert exchange_counts(17) == ([17], 0, 17)
    t = torch.zeros((2, 2, 3, 5))
    assert gather_columns(t, [5]) is t
