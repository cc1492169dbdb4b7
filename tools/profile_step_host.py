"""Host-side cost of one small time step (BASELINE config 1 sizes): cProfile over update_positions +
coherency_calc for 1000 flat-spectrum sources, where the GPU work is a few microseconds and Python is the step."""
import cProfile
import os
import pstats
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

entry.build()
from pyradiosky_b200 import EarthLocation, SkyModel, Time  # noqa: E402

rng = np.random.default_rng(1)
n = 1000
st = np.zeros((4, 1, n))
st[0, 0] = 10 ** rng.uniform(-2, 1, n)
sky = SkyModel(ra=rng.uniform(0, 2 * np.pi, n), dec=np.arcsin(rng.uniform(-1, 1, n)), stokes=st, spectral_type="flat")
loc = EarthLocation.from_dms((-30, 43, 17.5), (21, 25, 41.9), 1073.0)
t0 = Time.from_utc("2015-03-01 00:00:00")


def steps(k0, count):
    for k in range(count):
        sky.update_positions(t0 + 1e-4 * (k0 + k), loc)
        sky.coherency_calc()


steps(0, 50)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
steps(50, 500)
torch.cuda.synchronize()
pr.disable()
ps = pstats.Stats(pr).sort_stats("cumulative")
ps.print_stats(28)
