#!/usr/bin/env python
"""All-gather of the horizon-masked coherency over NVLink (dist.PeerGather): copy-engine route by streams x column slabs,
against the store route, on the c4 split of bench.py (torchrun, one rank per GPU)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import bench

    class A:
        pass

    ctx = bench.Ctx(A())
    from pyradiosky_b200.dist import PeerGather, exchange_counts

    sky, cat = bench.make_sky(ctx, "c4")
    freqs = cat["freqs"][::4][:256]
    sky.update_positions(ctx.time_at(3), ctx.loc)
    _, n_above = sky.compact_above_horizon()
    counts, col0, total = exchange_counts(n_above, ctx.dev)
    nbytes = 64.0 * len(freqs) * total
    recv = nbytes * (ctx.world - 1) / ctx.world
    (nmax,), _ = ctx.reduce([sky.Ncomponents], [0.0])
    res = {"world": ctx.world, "bytes": nbytes, "rows": []}

    def timed(fn):
        fn()
        best = 1e30
        for _ in range(3):
            ctx.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ctx.stream)
            fn()
            e1.record(ctx.stream)
            ctx.barrier()
            (ms,), _ = ctx.reduce([e0.elapsed_time(e1)], [0.0])
            best = min(best, ms)
        return best

    for streams in (1, 2, 4, 7, 14):
        pg = PeerGather(len(freqs), total + int(nmax), device=ctx.dev, streams=streams)
        for slabs in (1, 2, 4, 8):
            ms = timed(lambda: pg.coherency_calc(sky, freqs=freqs, route="copy", slabs=slabs))
            res["rows"].append({"route": "copy", "streams": streams, "slabs": slabs, "ms": ms, "GBps_recv": recv / ms / 1e6})
        if streams in (2, 4):
            ms = timed(lambda: pg.coherency_calc(sky, freqs=freqs, route="store"))
            res["rows"].append({"route": "store", "streams": streams, "ms": ms, "GBps_recv": recv / ms / 1e6})
        del pg
        ctx.free()
    if ctx.rank == 0:
        for r in res["rows"]:
            print(json.dumps(r))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
