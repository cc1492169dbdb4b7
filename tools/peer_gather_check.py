#!/usr/bin/env python
"""Peer-memory assembly of the sharded coherency array (run under torchrun, one rank per GPU).

Every rank's K2 kernel writes its compacted columns straight into the gathered (2, 2, Nf, sum N_above) array in
the destination GPU's HBM over NVLink (pyradiosky_b200.dist.PeerGather); compared, bit for bit, with the
unsharded computation and, for time, with the NCCL route (local coherency_calc + all_gather + re-pack).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_gather_check.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps=5):
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = min(best, float(t.item()))
    return best, out


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from pyradiosky_b200 import EarthLocation, SkyModel, Time
    from pyradiosky_b200 import dist as rdist

    n, nf = int(os.environ.get("GC_N", 400_000)), int(os.environ.get("GC_NF", 256))
    rng = np.random.default_rng(4)  # the same catalogue on every rank
    lon = rng.uniform(0, 2 * np.pi, n)
    lat = np.arcsin(rng.uniform(-1, 1, n))
    i = 10 ** rng.uniform(-2, 1, n)
    st = np.zeros((4, 1, n))
    st[0, 0], st[1, 0], st[2, 0] = i, 0.05 * i, -0.03 * i
    kwargs = dict(lon=lon, lat=lat, stokes=st, spectral_type="spectral_index",
                  spectral_index=np.clip(rng.normal(-0.8, 0.2, n), -2, 1), reference_frequency=np.full(n, 200e6))
    freqs = np.linspace(100e6, 200e6, nf)
    loc = EarthLocation.from_dms((-30, 43, 17.5), (21, 25, 41.9), 1073.0)
    t = Time.from_utc("2015-03-01 00:00:00")
    mine = SkyModel(**rdist.shard_kwargs(kwargs, world, rank))
    mine.update_positions(t, loc)

    res = {"world": world, "n": n, "nf": nf}
    by_streams = {}
    for ns in (1, 2, 4):  # concurrent destinations
        pgs = rdist.PeerGather(nf, n, streams=ns)
        by_streams[ns], _ = timed(lambda: pgs.coherency_calc(mine, freqs=freqs))
        del pgs
        torch.cuda.empty_cache()
    res["peer_write_all_ranks_ms_by_streams"] = by_streams
    pg = rdist.PeerGather(nf, n)
    # all ranks get the array
    ms_all, (view, counts) = timed(lambda: pg.coherency_calc(mine, freqs=freqs))
    whole = SkyModel(**kwargs)
    whole.update_positions(t, loc)
    ref = whole.coherency_calc(freqs=freqs, pad=False)
    ok_all = tuple(view.shape) == tuple(ref.shape) and bool(torch.equal(view, ref))
    # only rank 0 gets it
    pg.buf.zero_()
    ms_root, (view0, _) = timed(lambda: pg.coherency_calc(mine, freqs=freqs, dst=0))
    ok_root = bool(torch.equal(view0, ref)) if rank == 0 else True
    # the kernels' own stores over NVLink (route="store") must give the same array
    pg.buf.zero_()
    ms_store, (view_s, _) = timed(lambda: pg.coherency_calc(mine, freqs=freqs, route="store"))
    ok_all = ok_all and bool(torch.equal(view_s, ref))
    res["peer_store_route_all_ranks_ms"] = ms_store
    # NCCL route: local result, all_gather of padded blocks, re-pack
    def nccl_route():
        local_out = mine.coherency_calc(freqs=freqs, pad=False)
        c, _, _ = rdist.exchange_counts(local_out.shape[-1])
        return rdist.gather_columns(local_out, c)

    ms_nccl, full = timed(nccl_route)
    ok_nccl = bool(torch.equal(full, ref))
    ms_local, _ = timed(lambda: mine.coherency_calc(freqs=freqs, pad=False))
    flags = torch.tensor([int(ok_all), int(ok_root), int(ok_nccl)], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    ok = bool(flags.min().item())
    if rank == 0:
        gb = ref.numel() * 16 / 1e9
        res.update({"counts": counts, "gathered_GB": gb, "equal_to_unsharded": ok,
                    "local_compute_only_ms": ms_local,
                    "peer_write_all_ranks_ms": ms_all, "peer_write_all_GBps_received_per_rank": gb * (world - 1) / world / (ms_all * 1e-3),
                    "peer_write_to_rank0_ms": ms_root, "peer_write_root_GBps_received": gb * (world - 1) / world / (ms_root * 1e-3),
                    "nccl_all_gather_route_ms": ms_nccl})
        print(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
