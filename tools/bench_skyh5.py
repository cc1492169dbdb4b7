"""File -> HBM throughput of the SkyH5 loader, against the PCIe copy it is bounded by.

Writes a synthetic 'full'-spectrum point catalogue with ``SkyModel.write_skyh5`` (uncompressed and
gzip), then times ``SkyModel.from_skyh5`` (file in the page cache -> pinned slabs -> HBM) and, beside
it, (a) one pinned-host -> device copy of the same number of bytes (the PCIe bound) and (b) the
two-step route a host library gives: read the whole cube into pageable host memory, then copy.

    python tools/bench_skyh5.py [--ncomp 1000000] [--nfreq 64] [--dir /tmp] > gpurun_out/skyh5_bench.json
"""

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps=3):
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        del out
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ncomp", type=int, default=1_000_000)
    ap.add_argument("--nfreq", type=int, default=64)
    ap.add_argument("--dir", default="/tmp")
    ap.add_argument("--threads", type=int, default=0)
    args = ap.parse_args()
    import __graft_entry__ as entry

    entry.build()
    from pyradiosky_b200 import SkyModel, skyh5

    n, nf = args.ncomp, args.nfreq
    g = torch.Generator(device="cuda").manual_seed(1)
    stokes = torch.randn((4, nf, n), dtype=torch.float64, device="cuda", generator=g)
    stokes[1:] *= 0.05
    # quantise so that gzip has something to find (real catalogues carry ~6 significant digits)
    stokes = torch.round(stokes * 1e6) / 1e6
    ra = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * (2 * np.pi)
    dec = torch.asin(torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1)
    sky = SkyModel(name=np.char.add("s", np.arange(n).astype(str)), ra=ra, dec=dec, stokes=stokes, spectral_type="full",
                   freq_array=np.linspace(100e6, 200e6, nf))
    nbytes = stokes.numel() * 8
    res = {"ncomp": n, "nfreq": nf, "stokes_bytes": nbytes, "host_threads": os.cpu_count()}

    host = torch.empty(stokes.shape, dtype=torch.float64).pin_memory()
    host.copy_(stokes)
    dev = torch.empty_like(stokes)
    t = timed(lambda: dev.copy_(host, non_blocking=True))
    res["pinned_h2d_copy"] = {"s": t, "GBps": nbytes / t / 1e9}

    for label, comp in (("plain", None), ("gzip", "gzip")):
        path = os.path.join(args.dir, f"bench_{label}.skyh5")
        t0 = time.perf_counter()
        sky.write_skyh5(path, clobber=True, data_compression=comp)
        tw = time.perf_counter() - t0
        fsize = os.path.getsize(path)
        t = timed(lambda: SkyModel.from_skyh5(path, nthreads=args.threads))
        entry_ = {"file_bytes": fsize, "write_s": tw, "write_GBps": nbytes / tw / 1e9, "from_skyh5_s": t,
                  "from_skyh5_GBps_of_array": nbytes / t / 1e9}
        with skyh5.H5File(path) as f:
            t = timed(lambda: f.read_device("/Data/stokes", "cuda", nthreads=args.threads))
            entry_["read_device_stokes_s"] = t
            entry_["read_device_GBps_of_array"] = nbytes / t / 1e9
            entry_["frac_of_pinned_copy"] = res["pinned_h2d_copy"]["s"] / t
            t = timed(lambda: torch.as_tensor(f.read("/Data/stokes", nthreads=args.threads)).to("cuda"))
            entry_["host_read_then_copy_s"] = t
            entry_["host_read_then_copy_GBps"] = nbytes / t / 1e9
            got = f.read_device("/Data/stokes", "cuda", nthreads=args.threads)
            entry_["bit_equal"] = bool(torch.equal(got, stokes))
        res[label] = entry_
        os.remove(path)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
