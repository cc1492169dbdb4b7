"""update_positions on 16M point components and on a HEALPix nside-1024 map, for ncu captures of K1."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

entry.build()
from pyradiosky_b200 import EarthLocation, SkyModel, Time  # noqa: E402

loc = EarthLocation.from_dms((-30, 43, 17.5), (21, 25, 41.9), 1073.0)
t0 = Time.from_utc("2015-03-01 00:00:00")
n = 16_000_000
lon = torch.rand(n, dtype=torch.float64, device="cuda") * (2 * np.pi)
lat = torch.asin(torch.rand(n, dtype=torch.float64, device="cuda") * 2 - 1)
st = torch.zeros((4, 1, n), dtype=torch.float64, device="cuda")
st[0] = 1.0
st[1] = 0.1
sky = SkyModel(lon=lon, lat=lat, stokes=st, spectral_type="flat")
sky.update_positions(t0, loc)
sky.update_positions(t0 + 1e-3, loc)
del sky, lon, lat, st
nside = 1024
st = torch.ones((4, 1, 12 * nside * nside), dtype=torch.float64, device="cuda")
st[1:] = 0
for order in ("ring", "nested"):
    hp = SkyModel(nside=nside, hpx_order=order, stokes=st, spectral_type="flat", component_type="healpix", stokes_unit="K")
    hp.update_positions(t0, loc)
    hp.update_positions(t0 + 1e-3, loc)
torch.cuda.synchronize()
print("ok")
