#!/usr/bin/env python
"""One eighth of the headline catalogue (125k x 1024) on one GPU: device time per step against the host's time to
queue a step, with and without the position overlap -- is the 8-GPU strong-scaling tail the host or the device?"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import __graft_entry__ as entry

    entry.build()
    import bench

    n = int(os.environ.get("N", 125_000))
    cat = bench.synth_catalog("n1m", n_override=n)
    from pyradiosky_b200 import EarthLocation, SkyModel, Time

    loc = EarthLocation.from_dms((-30, 43, 17.5), (21, 25, 41.9), 1073.0)
    freqs = np.linspace(100e6, 200e6, 1024)
    out = {}
    for overlap in (False, True):
        sky = SkyModel(**bench.model_kwargs(cat))
        if overlap:
            sky.enable_position_overlap()
        ld = -(-n // 16) * 16
        buf = torch.empty((2, 2, 1024, ld), dtype=torch.complex128, device="cuda")
        def step(k):
            sky.update_positions(Time(*bench.step_time(k)), loc)
            sky.coherency_calc(freqs=freqs, out=buf, sync=False)

        for k in range(8):
            step(k)
        torch.cuda.synchronize()
        steps = 40
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev = []
        sky._timing_events = ev
        e0.record()
        t0 = time.perf_counter()
        for k in range(8, 8 + steps):
            step(k)
        host = time.perf_counter() - t0
        e1.record()
        sky._timing_events = None
        torch.cuda.synchronize()
        out["overlap" if overlap else "serial"] = {
            "device_ms_per_step": e0.elapsed_time(e1) / steps, "host_queue_ms_per_step": 1e3 * host / steps,
            "k2_ms": sum(a.elapsed_time(b) for a, b in ev) / max(1, len(ev))}
        del sky, buf
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
