#!/usr/bin/env python
"""K2 launch shape for small per-GPU catalogues (the strong-scaling shards of bench.py at 8 GPUs): kernel time by
frequencies per block."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import __graft_entry__ as entry

    entry.build()
    import bench
    from pyradiosky_b200 import EarthLocation, SkyModel, Time

    loc = EarthLocation.from_dms((-30, 43, 17.5), (21, 25, 41.9), 1073.0)
    for workload, n in (("n1m", 125_000), ("c2", 37_500), ("c4", 125_000), ("n1m", 1_000_000)):
        cat = bench.synth_catalog(workload, 2, 0, n)
        sky = SkyModel(**bench.model_kwargs(cat))
        freqs = cat["freqs"]
        out = torch.empty((2, 2, len(freqs), -(-n // 16) * 16), dtype=torch.complex128, device="cuda")
        for fc in (16, 32, 48, 64, 128):
            os.environ["RSK_K2_FCHUNK"] = str(fc)
            ev = []
            for k in range(8):
                sky.update_positions(Time(*bench.step_time(k)), loc)
                sky._timing_events = ev if k >= 3 else None
                sky.coherency_calc(freqs=freqs, out=out, sync=False)
            sky._timing_events = None
            torch.cuda.synchronize()
            ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
            print(json.dumps({"workload": workload, "n": n, "fchunk": fc, "kernel_ms": ms}), flush=True)
        os.environ.pop("RSK_K2_FCHUNK")
        del sky, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
