#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, one rank per GPU): component-sharded coherency_calc on every rank,
count exchange, optional NCCL gather of the compacted blocks; verifies the gathered array against the
unsharded computation and times the gather (it is optional and far slower than compute).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/gather_check.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


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
    st[0, 0] = i
    st[1, 0] = 0.05 * i
    st[2, 0] = -0.03 * i
    kwargs = dict(lon=lon, lat=lat, stokes=st, spectral_type="spectral_index",
                  spectral_index=np.clip(rng.normal(-0.8, 0.2, n), -2, 1), reference_frequency=np.full(n, 200e6))
    freqs = np.linspace(100e6, 200e6, nf)
    loc = EarthLocation.from_dms((-30, 43, 17.5), (21, 25, 41.9), 1073.0)
    t = Time.from_utc("2015-03-01 00:00:00")

    mine = SkyModel(**rdist.shard_kwargs(kwargs, world, rank))
    mine.update_positions(t, loc)
    local_out = mine.coherency_calc(freqs=freqs, pad=False)
    counts, col0, total = rdist.exchange_counts(local_out.shape[-1])
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    full = rdist.gather_columns(local_out, counts)
    e1.record()
    torch.cuda.synchronize()
    gather_ms = e0.elapsed_time(e1)
    ok = True
    if rank == 0:
        whole = SkyModel(**kwargs)
        whole.update_positions(t, loc)
        ref = whole.coherency_calc(freqs=freqs, pad=False)
        ok = tuple(full.shape) == tuple(ref.shape) and bool(torch.equal(full, ref))
        # the other assembly route: every shard's kernel writes its own column range of one wide buffer
        wide = torch.zeros_like(ref)
        mine.coherency_calc(freqs=freqs, out=wide, col0=col0)
        ok = ok and bool(torch.equal(wide[..., col0: col0 + counts[0]], ref[..., col0: col0 + counts[0]]))
        print(json.dumps({"world": world, "n": n, "nf": nf, "counts": counts, "total": total, "equal_to_unsharded": ok,
                          "gather_ms": gather_ms, "gathered_GB": full.numel() * 16 / 1e9,
                          "gather_GBps_per_rank_recv": full.numel() * 16 * (world - 1) / world / (gather_ms * 1e-3) / 1e9}))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
