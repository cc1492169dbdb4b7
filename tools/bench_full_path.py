#!/usr/bin/env python
"""coherency_calc() on a stored 'full' Stokes cube (the reference-shaped per-time-step call): gathered reads from the
(4, Nf, N) cube against reads from the component-major copy (RskSpectrum.stokes_cm), over horizon-mask densities and
orderings.  96 B per evaluation algorithmic (32 read + 64 written)."""
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
    from pyradiosky_b200 import EarthLocation, SkyModel, Time

    peak = 6547.8
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    n, nf = int(os.environ.get("N", 300_000)), int(os.environ.get("NF", 1024))
    rng = np.random.default_rng(3)
    loc = EarthLocation.from_dms((-30, 43, 17.5), (21, 25, 41.9), 1073.0)
    t0 = Time.from_utc("2015-03-01 00:00:00")
    i = 10 ** rng.uniform(-2, 1, n)
    st = np.zeros((4, 1, n))
    st[0, 0], st[1, 0], st[2, 0], st[3, 0] = i, 0.05 * i, -0.02 * i, 0.01 * i
    base = SkyModel(lon=rng.uniform(0, 2 * np.pi, n), lat=np.arcsin(rng.uniform(-1, 1, n)), stokes=st,
                    spectral_type="spectral_index", spectral_index=np.clip(rng.normal(-0.8, 0.2, n), -2, 1),
                    reference_frequency=np.full(n, 200e6))
    sky = base.at_frequencies(np.linspace(100e6, 200e6, nf), inplace=False, run_check=False)
    del base
    sky.update_positions(t0, loc)
    real_mask = sky.above_horizon.clone()
    res = {"n": n, "nf": nf, "peak": peak, "rows": []}
    out = torch.empty((2, 2, nf, -(-n // 16) * 16), dtype=torch.complex128, device="cuda")

    def run(label, mask):
        sky.above_horizon = mask
        row = {"mask": label, "density": float(mask.float().mean())}
        for cm in ("0", "1"):
            os.environ["RSK_FULL_CM"] = cm
            events = []
            sky._timing_events = events
            for _ in range(6):
                sky.coherency_calc(out=out, sync=False)
            sky._timing_events = None
            torch.cuda.synchronize()
            ms = min(a.elapsed_time(b) for a, b in events[1:])
            evals = int(mask.sum()) * nf
            row["cm" + cm] = {"ms": ms, "frac_of_96B": 96.0 * evals / (ms * 1e-3) / 1e9 / peak}
        res["rows"].append(row)
        print(json.dumps(row), flush=True)

    run("real horizon, random catalogue order", real_mask)
    for d in (0.05, 0.25, 0.5, 0.75, 0.95, 1.0):
        run(f"random {d}", torch.rand(n, device="cuda") < d)
    runs = (torch.arange(n, device="cuda") // 4096) % 2 == 0
    run("contiguous runs of 4096", runs)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
