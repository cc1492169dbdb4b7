#!/usr/bin/env python
"""BASELINE config 5: LST sweep, N sources x T time steps of positions in one launch (K1 sweep).

Reports source.time position evaluations per second and the fraction of the HBM roofline at
17 B / source.time (alt, az, mask; 41 B with lmn).  Not the bench.py metric (that one is coherency
evaluations); numbers go to profiles/ and DESIGN.md.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--t", type=int, default=1000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--lmn", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=20000)
    args = ap.parse_args()
    import torch

    import __graft_entry__ as entry

    entry.build()
    import bench  # its cpu_baseline leg is the only place outside tests/ that runs the oracle
    from pyradiosky_b200 import EarthLocation, SkyModel, Time, astrometry

    # under torchrun: strong scaling, the n sources are sharded over the ranks (contiguous blocks), every rank
    # sweeps all T contexts; time = max over ranks (device events), value = n * T / that time
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    rng = np.random.default_rng(5)
    n_total, nt = args.n, args.t
    lon = rng.uniform(0, 2 * np.pi, n_total)
    lat = np.arcsin(rng.uniform(-1, 1, n_total))
    if world > 1:
        from pyradiosky_b200.dist import shard_bounds

        lo, hi = shard_bounds(n_total, world, rank)
        lon, lat = lon[lo:hi], lat[lo:hi]
    n = lon.size
    st = np.zeros((4, 1, n))
    st[0] = 1.0
    sky = SkyModel(ra=lon, dec=lat, frame="icrs", stokes=st, spectral_type="flat")
    loc = EarthLocation.from_dms((-30, 43, 17.5), (21, 25, 41.9), 1073.0)
    t0 = Time.from_utc("2015-03-01 00:00:00")
    th = time.perf_counter()
    astroms = astrometry.astrom_contexts([t0 + k * 86.164 / 86400.0 for k in range(nt)], loc)
    host_ctx_s = time.perf_counter() - th
    best = 1e30
    kev = []
    sky._timing_events = kev
    for r in range(args.reps + 1):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        alt_az, lmn, above = sky.positions_sweep(None, loc, astroms=astroms, want_lmn=args.lmn)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            ms = float(tmax.item())
        if r > 0:
            best = min(best, ms)
        if r < args.reps:
            del alt_az, lmn, above
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
        if rank != 0:
            return
    bytes_per = 41 if args.lmn else 17
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    # spot check against the oracle
    k = nt // 3
    ref_alt_az, _ = bench.positions_cpu_baseline(lon[:5000], lat[:5000], astroms[k:k + 1])
    got = alt_az[k][:, :5000].cpu().numpy()
    err_uas = np.abs(got[0] - ref_alt_az[0]).max() / (np.pi / 180 / 3600 * 1e-6)
    # CPU oracle on a sample, one core
    ns = min(n, args.cpu_sample)
    _, cpu_rate = bench.positions_cpu_baseline(lon[:ns], lat[:ns], astroms[:5])
    kernel_ms = min(a.elapsed_time(b) for a, b in kev[1:])
    out = {
        "metric": "position_source_time_evals_per_s", "value": n_total * nt / (best * 1e-3), "ms": best, "n": n_total,
        "t": nt, "n_gpus": world, "scaling": "strong", "sources_per_gpu": n,
        "kernel_ms": kernel_ms, "kernel_value": n * nt / (kernel_ms * 1e-3),
        "kernel_frac_of_hbm_peak": n * nt * (41 if args.lmn else 17) / (kernel_ms * 1e-3) / 1e9 / peak,
        "bytes_per_eval": bytes_per, "achieved_gbs": n_total * nt * bytes_per / (best * 1e-3) / 1e9,
        "frac_of_hbm_peak_per_gpu": n_total * nt * bytes_per / (best * 1e-3) / 1e9 / peak / world,
        "host_context_build_s": host_ctx_s,
        "max_alt_err_uas_vs_oracle": err_uas,
        "cpu_oracle_1core_evals_per_s": cpu_rate,
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
