#!/usr/bin/env python
"""Host-side ceilings of the end-to-end path: pinned device->host copy rate (one GPU per process; run under torchrun for the
concurrent figure) and the host's rate of expanding Hermitian-packed slabs (rsk_unpack_coherency_host) by thread count."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import __graft_entry__ as entry

    rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(rank)
    if rank == 0:
        entry.build()
    from pyradiosky_b200 import _lib

    lib = _lib.load()
    res = {"rank": rank, "cpus": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}
    nbytes = 2 << 30
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    for name, fn in (("d2h", lambda: host.copy_(dev, non_blocking=True)), ("h2d", lambda: dev.copy_(host, non_blocking=True))):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[name + "_GBps"] = 4 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
    del dev, host
    # expand rate: 64 frequencies x 150k columns per slab (what a C2 step unpacks 16 times)
    nf, ncols = 64, 150_000
    ld = -(-ncols // 32) * 32
    packed = torch.rand((4, nf, ld), dtype=torch.float64).pin_memory()
    for out_kind in ("pinned", "pageable"):
        out = torch.empty(4 * nf * ncols, dtype=torch.complex128)
        if out_kind == "pinned":
            out = out.pin_memory()
        out.zero_()
        for nt in (1, 2, 4, 8, 12, 16, 24, 32):
            if nt > 2 * res["affinity"]:
                break
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                lib.rsk_unpack_coherency_host(packed.data_ptr(), nf, ncols, ld, out.data_ptr(), nf, 0, ncols, nt)
                best = min(best, time.perf_counter() - t0)
            res[f"expand_{out_kind}_{nt}thr_GBps_out"] = 64.0 * nf * ncols / best / 1e9
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
