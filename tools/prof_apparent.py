"""One apparent_coherency_calc (Jones arrays, two antennas) at BASELINE config 2 size, for an ncu capture."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

entry.build()
from pyradiosky_b200 import EarthLocation, SkyModel, Time  # noqa: E402

rng = np.random.default_rng(7)
n, nf = 300_000, 1024
i = 10 ** rng.uniform(-2, 1, n)
st = np.zeros((4, 1, n))
st[0, 0], st[1, 0], st[2, 0], st[3, 0] = i, 0.05 * i, -0.02 * i, 0.01 * i
sky = SkyModel(lon=rng.uniform(0, 2 * np.pi, n), lat=np.arcsin(rng.uniform(-1, 1, n)), frame="fk5", stokes=st,
               spectral_type="spectral_index", spectral_index=np.clip(rng.normal(-0.8, 0.2, n), -2, 1),
               reference_frequency=np.full(n, 200e6))
sky.update_positions(Time.from_utc("2015-03-01 00:00:00"), EarthLocation.from_dms((-30, 43, 17.5), (21, 25, 41.9), 1073.0))
na = int(sky.above_horizon.sum().item())
freqs = np.linspace(100e6, 200e6, nf)
j1 = torch.randn((2, 2, nf, na), dtype=torch.complex128, device="cuda")
j2 = torch.randn((2, 2, nf, na), dtype=torch.complex128, device="cuda")
out = sky.apparent_coherency_calc(jones=j1, jones2=j2, freqs=freqs)
torch.cuda.synchronize()
print("apparent", tuple(out.shape))
if "--time" in sys.argv:  # throughput of the two Jones-array variants (not under ncu)
    peak = 6547.8
    for label, kw, nbytes in (("one antenna", dict(jones=j1), 128), ("two antennas", dict(jones=j1, jones2=j2), 192)):
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = sky.apparent_coherency_calc(freqs=freqs, **kw)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(f"{label}: {best:.3f} ms  {nbytes * na * nf / best / 1e6:.0f} GB/s  frac {nbytes * na * nf / best / 1e6 / peak:.3f}")
