#!/usr/bin/env python
"""Host-side ceilings, second look: (a) expansion rate into a transparent-huge-page destination, (b) D2H copies and
expansion running at the same time without any dependency between them (what the box sustains when both compete for
host DRAM), (c) expansion from a small cache-resident staging buffer (what the reads cost)."""
import json
import mmap
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import __graft_entry__ as entry

    entry.build()
    from pyradiosky_b200 import _lib

    lib = _lib.load()
    res = {}
    nf, ncols = 64, 150_000
    ld = -(-ncols // 32) * 32
    packed = torch.rand((4, nf, ld), dtype=torch.float64).pin_memory()
    nout = 4 * nf * ncols

    def expand_rate(out_ptr, nt, src=packed, reps=3, nf_=nf):
        best = 1e9
        for _ in range(reps):
            t0 = time.perf_counter()
            lib.rsk_unpack_coherency_host(src.data_ptr(), nf_, ncols, ld, out_ptr, nf_, 0, ncols, nt)
            best = min(best, time.perf_counter() - t0)
        return 64.0 * nf_ * ncols / best / 1e9

    # (a) THP destination
    size = nout * 16
    m = mmap.mmap(-1, size + (2 << 20))
    m.madvise(mmap.MADV_HUGEPAGE)
    arr = np.frombuffer(m, dtype=np.uint8)
    off = (-arr.ctypes.data) % (2 << 20)
    thp = torch.from_numpy(arr[off:off + size]).view(torch.complex128)
    thp.zero_()
    plain = torch.empty(nout, dtype=torch.complex128)
    plain.zero_()
    pinned = torch.empty(nout, dtype=torch.complex128).pin_memory()
    pinned.zero_()
    for name, t in (("thp", thp), ("malloc", plain), ("pinned", pinned)):
        for nt in (8, 12, 15, 16):
            res[f"expand_{name}_{nt}thr"] = expand_rate(t.data_ptr(), nt)
    try:
        print(open("/proc/meminfo").read().split("AnonHugePages")[1].split("\n")[0], file=sys.stderr)
    except Exception:
        pass
    # (c) small staging: 2 channels = 9.6 MB of packed input, reused
    small = torch.rand((4, 2, ld), dtype=torch.float64).pin_memory()
    for nt in (2, 8, 15):
        res[f"expand_pinned_small_input_{nt}thr"] = expand_rate(pinned.data_ptr(), min(nt, 2), src=small, reps=20, nf_=2)
    # (b) concurrent D2H + expansion
    dev = torch.rand((4, nf, ld), dtype=torch.float64, device="cuda")
    hstage = [torch.empty((4, nf, ld), dtype=torch.float64).pin_memory() for _ in range(2)]
    stop = threading.Event()
    copied = [0]

    def copier():
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            while not stop.is_set():
                for h in hstage:
                    h.copy_(dev, non_blocking=True)
                s.synchronize()
                copied[0] += 2
    for nt in (8, 12, 15):
        stop.clear()
        copied[0] = 0
        th = threading.Thread(target=copier)
        th.start()
        time.sleep(0.05)
        c0, t0 = copied[0], time.perf_counter()
        nrep = 12
        for _ in range(nrep):
            lib.rsk_unpack_coherency_host(packed.data_ptr(), nf, ncols, ld, pinned.data_ptr(), nf, 0, ncols, nt)
        dt = time.perf_counter() - t0
        c1 = copied[0]
        stop.set()
        th.join()
        res[f"concurrent_{nt}thr"] = {"expand_GBps_out": nrep * 64.0 * nf * ncols / dt / 1e9,
                                      "d2h_GBps": (c1 - c0) * dev.numel() * 8 / dt / 1e9}
    # plain-route analogue: D2H of full rows alone, into pinned memory, while nothing else runs
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(6):
        hstage[0].copy_(dev, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    res["d2h_alone_GBps"] = 6 * dev.numel() * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
