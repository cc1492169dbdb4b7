"""Where an end-to-end step spends its time (BASELINE config 2 sizes): model construction from pinned host
arrays, update_positions, and coherency_calc_to_host by route.  Host-side wall clock with a device sync
after each phase."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as entry  # noqa: E402

entry.build()
from pyradiosky_b200 import EarthLocation, SkyModel, Time  # noqa: E402

n = int(os.environ.get("N", 300_000))
nf = 1024
rng = np.random.default_rng(2)
i = 10 ** rng.uniform(-2, 1, n)
st = np.zeros((4, 1, n))
st[0, 0], st[1, 0], st[2, 0], st[3, 0] = i, 0.05 * i, -0.02 * i, 0.01 * i
host = dict(lon=rng.uniform(0, 2 * np.pi, n), lat=np.arcsin(rng.uniform(-1, 1, n)), stokes=st,
            spectral_index=np.clip(rng.normal(-0.8, 0.2, n), -2, 1), reference_frequency=np.full(n, 200e6))
pinned = {k: torch.as_tensor(v).pin_memory() for k, v in host.items()}
freqs = np.linspace(100e6, 200e6, nf)
loc = EarthLocation.from_dms((-30, 43, 17.5), (21, 25, 41.9), 1073.0)
out = torch.empty(4 * nf * n, dtype=torch.complex128).pin_memory()
threads = int(os.environ.get("THREADS", max(1, (3 * (os.cpu_count() or 1)) // 4)))
res = {"n": n, "nf": nf, "cpu_count": os.cpu_count(), "unpack_threads": threads}


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


for route in ("auto", True, False, "auto", True, False):
    t0 = sync()
    sky = SkyModel(frame="fk5", spectral_type="spectral_index", **{k: v.to("cuda", non_blocking=True) for k, v in pinned.items()})
    t1 = sync()
    sky.update_positions(Time(2457082.5, 0.01 * len(res)), loc)
    t2 = sync()
    view = sky.coherency_calc_to_host(freqs, out, slab=64, packed=route, nthreads=threads)
    t3 = sync()
    key = {"auto": "mixed", True: "packed", False: "plain"}[route]
    res.setdefault(key, []).append({"construct_ms": (t1 - t0) * 1e3, "positions_ms": (t2 - t1) * 1e3, "to_host_ms": (t3 - t2) * 1e3,
                                    "evals": view.shape[2] * view.shape[3],
                                    "counts": getattr(sky, "_e2e_route_counts", None) if route == "auto" else None})
print(json.dumps(res, indent=1))
