#!/usr/bin/env python
"""Benchmark of the per-component evaluation hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload n1m|c1|c2|c3|c4|full] [--impl reference]

One *step* = one time step of the hot path over one synthetic catalogue:
``update_positions(time_k, HERA)`` (K1 positions + horizon mask), horizon compaction (K1b) and the fused
``coherency_calc(freqs=...)`` (K2: spectrum + polarisation rotation + Stokes->coherency), i.e. what the
reference does with ``at_frequencies`` once and ``update_positions`` + ``coherency_calc`` per time step.
Every step uses a new time, so nothing is cached between steps.

Headline workload (default): the north-star target shape, 1M polarised components x 1024 frequencies, power law
(``n1m``).  Under ``--gpus N`` the SAME catalogue is sharded by components over the N ranks (strong scaling, no
data-path collective); the catalogue is generated in 8 seeded blocks so it is identical for every N.

metric  = Az/El coherency component.frequency evaluations per second (one eval = one 2x2 complex128
          coherency written for one above-horizon component at one frequency), whole job over all ranks.
value   = inputs already resident in HBM, CUDA-event timed, max over ranks.
roofline= the dominant kernel (k_eval, coherency epilogue) against the measured HBM copy bandwidth; every number
          needed to recompute ``frac`` is in the record (algorithmic bytes per launch, kernel ms, peak).
configs = the same measurement for every BASELINE.json config: c1, c2, c3 (HEALPix nside 1024 x 256, un-chunked),
          c4 (1M subband sources -> 1024 channels), c5 (LST sweep, 1M sources x 1000 contexts: positions only, its
          own unit) and ``full`` (the reference-shaped ``coherency_calc()`` on a stored Stokes cube); c2..c5 and
          ``full`` are sharded over the ranks like the headline (strong), c1 runs on rank 0.
e2e     = the same metric through the public SkyModel API with HOST buffers: pinned host inputs are
          uploaded and the full coherency array is delivered into host memory inside the timed region, every step;
          ``e2e.roofline`` holds the ceilings measured in the same run (concurrent pinned D2H rate of the N ranks, host
          expansion rate) and the fraction of them the step reaches.  ``e2e.visibility`` is the fused consumer
          (``visibility_calc``: nothing per-component leaves the GPU), reported beside it, not instead of it.
gather  = (N > 1) assembly of the horizon-masked output of the c4 split on one / on every GPU: the coherency kernels
          storing straight into peer HBM over NVLink (``dist.PeerGather``) and the NCCL all-gather route, timed apart
          from the step as north_star asks.
cpu_baseline = the CPU oracle (restated reference algorithm; astropy is absent so the reference itself
          cannot run) on a bounded sample of the same workload, 1 core, rank 0, N=1 only.

``--impl reference`` times that CPU oracle with all host cores (component-sharded processes) on the same config.
"""

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "azel_coherency_component_freq_evals_per_s"
UNIT = "evals/s"
GLEAM_BANDS = np.array([76, 84, 92, 99, 107, 115, 122, 130, 143, 151, 158, 166, 174, 181, 189, 197, 204,
                        212, 220, 227]) * 1e6
SIDEREAL_DAY = 0.99726958
NBLOCKS = 8  # every catalogue is generated in 8 seeded blocks: identical for 1, 2, 4 and 8 ranks

WORKLOADS = {
    # name: (description, n components (whole job), nf, spectral_type, frame)
    "c1": ("BASELINE configs[0]: 1,000 flat-spectrum Stokes-I point sources, one time at HERA", 1000, 1, "flat", "icrs"),
    "c2": ("BASELINE configs[1]: GLEAM-like 300k point sources, full Stokes, power-law spectral index, "
           "1024 frequencies, fk5", 300_000, 1024, "spectral_index", "fk5"),
    "c3": ("BASELINE configs[2]: HEALPix nside=1024 ring map (12.6M pixels), full Stokes, power law, "
           "256 frequencies, one launch per time step (no frequency chunking); N > 1: pixels dealt out in blocks of 4096",
           12 * 1024 * 1024, 256, "spectral_index", "icrs"),
    "c4": ("BASELINE configs[3]: 1M polarized sources, 20 GLEAM subbands cubic-interpolated to 1024 channels",
           1_000_000, 1024, "subband", "icrs"),
    "n1m": ("north_star target: 1M polarized sources x 1024 frequencies, power law", 1_000_000, 1024,
            "spectral_index", "icrs"),
    "full": ("reference-shaped call: at_frequencies() once, then update_positions + coherency_calc() on the stored "
             "(4, 1024, 300k) Stokes cube per time step", 300_000, 1024, "spectral_index", "fk5"),
}
C5 = ("BASELINE configs[4]: LST sweep, 1M sources x 1000 time steps (per-time apco context), positions + horizon mask",
      1_000_000, 1000)


def shard_bounds(n, world, rank):
    per = -(-int(n) // world)
    lo = min(int(n), rank * per)
    return lo, min(int(n), lo + per)


def _synth_block(workload, nb, seed, block):
    """Block ``block`` (nb components) of a synthetic catalogue per SURVEY.md section 8(d)."""
    _, _, nf, stype, _ = WORKLOADS[workload]
    rng = np.random.default_rng([seed, block])
    cat = {}
    if workload != "c3":
        cat["lon"] = rng.uniform(0, 2 * np.pi, nb)
        cat["lat"] = np.arcsin(rng.uniform(-1, 1, nb))
    i = 10 ** rng.uniform(-2, 1, nb)
    st = np.zeros((4, 1, nb))
    st[0, 0] = i
    if workload != "c1":
        p = rng.uniform(0, 0.1, nb)
        psi = rng.uniform(0, np.pi, nb)
        st[1, 0] = p * i * np.cos(2 * psi)
        st[2, 0] = p * i * np.sin(2 * psi)
        st[3, 0] = 0.01 * i * rng.normal(size=nb)
    if stype == "spectral_index":
        cat["spectral_index"] = np.clip(rng.normal(-0.8, 0.2, nb), -2, 1)
        cat["reference_frequency"] = np.full(nb, 200e6)
    elif stype == "subband":
        alpha = np.clip(rng.normal(-0.8, 0.2, nb), -2, 1)
        st = st * (GLEAM_BANDS[None, :, None] / 200e6) ** alpha[None, None, :]
        st = st * (1 + 0.02 * rng.normal(size=st.shape))
    if os.environ.get("RSK_BENCH_UNPOLARIZED"):  # tuning aid: no component takes the rotation branch
        st[1:] = 0.0
    cat["stokes"] = st
    return cat


def synth_catalog(workload, seed=2, lo=None, hi=None, n_override=None):
    """Components [lo, hi) of the workload's synthetic catalogue as a dict of host numpy arrays.  The catalogue is
    built from NBLOCKS independently seeded blocks, so any rank count that divides NBLOCKS sees the same catalogue."""
    _, n, nf, stype, frame = WORKLOADS[workload]
    if n_override:
        n = n_override
    lo = 0 if lo is None else lo
    hi = n if hi is None else hi
    bsize = -(-n // NBLOCKS)
    parts = []
    for b in range(lo // bsize, -(-hi // bsize) if hi > lo else lo // bsize):
        b0, b1 = b * bsize, min(n, (b + 1) * bsize)
        blk = _synth_block(workload, b1 - b0, seed, b)
        s0, s1 = max(lo, b0) - b0, min(hi, b1) - b0
        parts.append({k: v[..., s0:s1] for k, v in blk.items()})
    cat = {"n": hi - lo, "n_total": n, "lo": lo, "nf": nf, "spectral_type": stype, "frame": frame}
    if parts:
        for k in parts[0]:
            cat[k] = np.ascontiguousarray(np.concatenate([p[k] for p in parts], axis=-1))
    else:
        cat["stokes"] = np.zeros((4, len(GLEAM_BANDS) if stype == "subband" else 1, 0))
    if workload == "c3":
        cat["nside"] = 1024  # pixel centres are generated on the device; the oracle makes its own
    if stype == "subband":
        cat["freq_array"] = GLEAM_BANDS.copy()
        cat["freqs"] = np.linspace(80e6, 220e6, nf)
    elif stype == "spectral_index":
        cat["freqs"] = np.linspace(100e6, 200e6, nf)
    else:
        cat["freqs"] = np.array([150e6]) if nf == 1 else np.linspace(100e6, 200e6, nf)
    return cat


def model_kwargs(cat):
    kw = dict(frame=cat["frame"], stokes=cat["stokes"], spectral_type=cat["spectral_type"])
    if "nside" in cat:
        kw.update(nside=cat["nside"], hpx_order="ring", component_type="healpix", stokes_unit="K", hpx_first=cat["lo"])
    else:
        kw.update(lon=cat["lon"], lat=cat["lat"])
    if "spectral_index" in cat:
        kw["spectral_index"] = cat["spectral_index"]
        kw["reference_frequency"] = cat["reference_frequency"]
    if "freq_array" in cat:
        kw["freq_array"] = cat["freq_array"]
    return kw


def step_time(k):
    """UTC two-part JD of time step k: 2015-03-01 00:00 UTC + k * (sidereal day / 97)."""
    return 2457082.5, (k * SIDEREAL_DAY / 97.0) % 1.0


HERA = dict(lat_dms=(-30, 43, 17.5), lon_dms=(21, 25, 41.9), height=1073.0)


def config_record(workload, world, headline=True):
    """The ``config`` object: the same keys and values for the GPU arm and the reference arm."""
    desc, n, nf, stype, frame = WORKLOADS[workload]
    return {
        "workload": f"{workload}: {desc}", "components": n, "components_per_gpu": -(-n // world), "nfreqs": nf,
        "spectral_type": stype, "frame": frame, "site": "HERA",
        "step": "update_positions + coherency_calc at a new time (GPU arm: enable_position_overlap(), then per step "
                "update_positions + coherency_calc(freqs, out=buf, sync=False) when the full-width buffer fits, else "
                "update_positions + coherency_calc(freqs))",
        "sharding": "the same catalogue split into contiguous component blocks, one per rank, no data-path collective",
        "l2": f"output per step {0.5 * n * nf * 64 / world / 1e9:.2f} GB per GPU >> 126 MB L2, no flush needed",
    }


# ---------------------------------------------------------------------------------------------------
# CPU oracle timing (cpu_baseline leg and --impl reference)
# ---------------------------------------------------------------------------------------------------
def oracle_step(cat, k, sl=slice(None)):
    """One step of the reference algorithm on the CPU oracle for components ``sl``. Returns evals."""
    from oracle import erfa_np, skymodel_np

    lat = -np.deg2rad(30 + 43 / 60 + 17.5 / 3600)
    lon = np.deg2rad(21 + 25 / 60 + 41.9 / 3600)
    jd1, jd2 = step_time(k)
    astrom = erfa_np.astrom_context(jd1, jd2, lon, lat, 1073.0)
    m = None
    if cat["frame"] != "icrs":
        m = erfa_np.frame_to_icrs(cat["frame"])
    if "nside" in cat:
        from oracle import healpix_np

        npix = 12 * cat["nside"] ** 2
        pix = (np.arange(cat["n"]) * (npix // cat["n"]))[sl]  # a sample strides over the whole sphere
        clon, clat = healpix_np.healpix_to_lonlat(cat["nside"], pix)
    else:
        clon, clat = cat["lon"][sl], cat["lat"][sl]
    st = cat["stokes"][:, :, sl]
    alt_az, _, above = skymodel_np.update_positions(clon, clat, astrom, m)
    full = skymodel_np.at_frequencies(
        st, cat["spectral_type"], cat["freqs"], freq_array=cat.get("freq_array"),
        reference_frequency=None if "reference_frequency" not in cat else cat["reference_frequency"][sl],
        spectral_index=None if "spectral_index" not in cat else cat["spectral_index"][sl])
    ravg = skymodel_np.calc_average_rotation_matrix(astrom)
    out = skymodel_np.coherency_calc(full, clon, clat, alt_az, above, ravg,
                                     polarized=skymodel_np.polarized_indices(st))
    return int(out.shape[2]) * int(out.shape[3])


def _oracle_worker(args):
    workload, seed, n_sample, k, lo, hi = args
    cat = _WORKER_CACHE.get((workload, seed, n_sample))
    if cat is None:
        cat = synth_catalog(workload, seed, n_override=n_sample)
        _WORKER_CACHE[(workload, seed, n_sample)] = cat
    return oracle_step(cat, k, slice(lo, hi))


_WORKER_CACHE = {}


def time_oracle(workload, n_sample, steps, warmup, procs):
    """Oracle throughput (evals/s) on an n_sample-component sample of the workload with ``procs`` processes."""
    if procs <= 1:
        cat = synth_catalog(workload, 2, n_override=n_sample)
        for k in range(warmup):
            oracle_step(cat, k)
        t0 = time.perf_counter()
        evals = sum(oracle_step(cat, warmup + k) for k in range(steps))
        dt = time.perf_counter() - t0
        return evals / dt, dt / steps
    import multiprocessing as mp

    bounds = np.linspace(0, n_sample, procs + 1).astype(int)
    with mp.get_context("fork").Pool(procs) as pool:
        def run(k):
            return sum(pool.map(_oracle_worker, [(workload, 2, n_sample, k, int(bounds[i]), int(bounds[i + 1]))
                                                 for i in range(procs)]))
        for k in range(warmup):
            run(k)
        t0 = time.perf_counter()
        evals = sum(run(warmup + k) for k in range(steps))
        dt = time.perf_counter() - t0
    return evals / dt, dt / steps


def run_reference(args):
    """--impl reference: the reference's CPU path for the same metric/config, restated (the oracle),
    component-sharded over all host cores.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    desc, n, nf, stype, frame = WORKLOADS[args.workload]
    procs = max(1, min(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1), 64))
    n_sample = min(n, args.ref_sample_per_core * procs)
    value, s_per_step = time_oracle(args.workload, n_sample, args.steps, args.warmup, procs)
    sample = (f"{n_sample} of {n} components x {nf} frequencies per step, sharded over {procs} processes; "
              "restated oracle (numpy/scipy, astropy absent), per-source Python loops kept; the rate per evaluation does "
              "not depend on the sample size")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_record(args.workload, max(1, world)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def positions_cpu_baseline(lon, lat, astroms):
    """cpu_baseline leg of the LST-sweep record (BASELINE config 5): the oracle's update_positions on one core.
    Returns (alt_az of the first context, source.time evaluations per second).  Lives here because bench.py is the
    one place outside tests/ that may execute oracle/ (as checker and as the CPU baseline, never as the thing
    measured)."""
    from oracle import skymodel_np

    t0 = time.perf_counter()
    first = None
    for a in astroms:
        alt_az, _, _ = skymodel_np.update_positions(lon, lat, a)
        if first is None:
            first = alt_az
    return first, len(astroms) * lon.size / (time.perf_counter() - t0)


# ---------------------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU with NVML while the timed region runs."""

    BAD = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown"}
    NOTE = {0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x1: "gpu_idle", 0x2: "applications_clocks_setting"}

    def __init__(self, index, period=0.02):
        self.index, self.period = index, period
        self.samples, self.reasons = [], 0
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if visible:
                ids = [v for v in visible.split(",") if v.strip() != ""]
                if index < len(ids) and ids[index].strip().isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                self.reasons |= int(get(self.handle))
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        names = [n for bit, n in {**self.BAD, **self.NOTE}.items() if self.reasons & bit and n != "gpu_idle"]
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": names,
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def k2_algorithmic_bytes(cat, n_above, stored_full=False):
    """Algorithmic HBM bytes of one fused coherency launch (DESIGN.md 'Kernels'):
    64 B written per eval + per above-horizon component, read once: idx 8, pol flag 1, rotation coefficients 24,
    Stokes 32 * nf_in (spline coefficients for subband: 32 * nc), (alpha, f_ref) 16; + the frequency axis.
    Stored-'full' cubes: 32 B read + 64 B written per eval."""
    nf = cat["nf"]
    if stored_full:
        return 96 * nf * n_above + (8 + 1 + 24) * n_above
    nf_in = cat["stokes"].shape[1]
    per_comp = 8 + 1 + 24 + 32 * nf_in + (16 if cat["spectral_type"] == "spectral_index" else 0)
    return 64 * nf * n_above + per_comp * n_above + 8 * nf


def traffic_record(workload):
    """``roofline.traffic``: dram__bytes_read + dram__bytes_write per launch of the dominant kernel from the committed
    ``ncu --set full`` capture of this workload (profiles/k_eval_traffic.json) -- reported only while the kernel sources
    are the ones the capture was taken with; otherwise null, never a stale constant."""
    import hashlib

    path = os.path.join(ROOT, "profiles", "k_eval_traffic.json")
    try:
        entry = json.load(open(path)).get(workload)
        h = hashlib.sha1()
        for f in ("rsk_kernels.cu", "rsk_device.cuh"):
            h.update(open(os.path.join(ROOT, "pyradiosky_b200", "csrc", f), "rb").read())
    except (OSError, ValueError):
        entry = None
    if not entry:
        return {"traffic": None, "traffic_source": "no ncu capture committed for this workload"}
    if entry.get("kernel_source_sha1") != h.hexdigest():
        return {"traffic": None, "traffic_source": f"capture {entry.get('capture')} predates the current kernel sources"}
    return {"traffic": entry["traffic_bytes"], "traffic_source": f"{entry.get('capture')}; kernel sources unchanged since "
            f"(sha1 {h.hexdigest()[:12]}); per launch of one GPU holding the whole catalogue"}


class Ctx:
    """Per-process state of the GPU arm."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        import __graft_entry__ as entry

        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            torch.cuda.set_device(self.local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        else:
            torch.cuda.set_device(0)
        if self.rank == 0:
            entry.build()  # no-op when the prebuilt librsk.so is current
        if self.world > 1:
            dist.barrier()  # nobody loads the library while rank 0 might be rebuilding it
        import pyradiosky_b200
        from pyradiosky_b200 import _lib

        _lib.load()
        self.rsk, self._lib = pyradiosky_b200, _lib
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.loc = pyradiosky_b200.EarthLocation.from_dms(HERA["lat_dms"], HERA["lon_dms"], HERA["height"])
        self.stream = torch.cuda.current_stream()
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            self.peak, self.peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            self.peak, self.peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        self.cpus = self._pin_cpus()

    def _pin_cpus(self):
        """At N > 1 give every rank its own share of the host cores, taken from the cores next to its GPU when the
        topology says which those are (/sys/bus/pci/devices/<gpu>/local_cpulist), so that the staging buffers (first
        touched by this process) and the unpack threads stay on the GPU's NUMA node."""
        try:
            allowed = sorted(os.sched_getaffinity(0))
        except AttributeError:
            return os.cpu_count() or 1
        if self.world == 1:
            return len(allowed)
        local = None
        try:
            prop = self.torch.cuda.get_device_properties(self.dev)
            bus = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
            txt = open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read().strip()
            cpus = set()
            for part in txt.split(","):
                if "-" in part:
                    a, b = part.split("-")
                    cpus.update(range(int(a), int(b) + 1))
                elif part:
                    cpus.add(int(part))
            local = sorted(cpus & set(allowed))
        except Exception:
            local = None
        # ranks that share a core pool split it evenly; the pool is the GPU-local list when it is a proper subset
        pool = local if local and len(local) < len(allowed) else allowed
        key = self.torch.tensor([hash(tuple(pool)) % (1 << 31)], dtype=self.torch.int64, device=self.dev)
        keys = [self.torch.empty_like(key) for _ in range(self.world)]
        self.dist.all_gather(keys, key)
        same = [r for r in range(self.world) if int(keys[r].item()) == int(key.item())]
        share = len(pool) // max(1, len(same))
        if share >= 1:
            j = same.index(self.rank)
            mine = pool[j * share:(j + 1) * share]
            try:
                os.sched_setaffinity(0, mine)
                return len(mine)
            except OSError:
                pass
        return max(1, len(allowed) // self.world)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, maxes, sums):
        """MAX over ranks of ``maxes`` and SUM over ranks of ``sums`` (lists of floats)."""
        torch = self.torch
        m = torch.tensor(list(maxes), dtype=torch.float64, device=self.dev)
        s = torch.tensor(list(sums), dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(m, op=self.dist.ReduceOp.MAX)
            self.dist.all_reduce(s, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in m.cpu()], [float(v) for v in s.cpu()]

    def free(self):
        import gc

        gc.collect()
        self.torch.cuda.empty_cache()

    def time_at(self, k):
        return self.rsk.Time(*step_time(k))


def measure_coherency(ctx, sky, cat, freqs, steps, warmup, stored_full=False):
    """Device-resident timing of ``steps`` time steps.  Returns a dict of per-rank numbers.

    mode "async": the step is queued without any device->host read -- the output goes to a preallocated buffer wide
    enough for every component and the above-horizon count stays on the device (coherency_calc(..., out=buf,
    sync=False)); counts are read once after the timed region.  mode "sync" (when that buffer would not fit): the
    reference-shaped call, which returns an exactly sized array and reads the count back every step."""
    torch = ctx.torch
    n, nf = sky.Ncomponents, (sky.Nfreqs if freqs is None else len(freqs))
    ld_full = -(-n // 16) * 16
    free, _ = torch.cuda.mem_get_info(ctx.dev)
    can_async = 64 * nf * ld_full < 0.8 * free
    out_buf = torch.empty((2, 2, nf, ld_full), dtype=torch.complex128, device=ctx.dev) if can_async else None
    half_buf = None
    if not can_async:
        # the result alone is more than half of HBM (c3 on one GPU: 103 GB): one buffer with room for 53 % of the
        # components, reused by every step (the count is read back per step, the reference-shaped call)
        ld_half = -(-int(0.53 * n) // 16) * 16
        if 64 * nf * ld_half < 0.9 * free:
            half_buf = torch.empty((2, 2, nf, ld_half), dtype=torch.complex128, device=ctx.dev)

    if can_async and not os.environ.get("RSK_BENCH_NO_OVERLAP"):
        # positions + compaction of step k + 1 run on a side stream beside the coherency kernel of step k
        sky.enable_position_overlap()

    def run_steps(first, nsteps, events):
        counts = []
        sky._timing_events = events
        for k in range(first, first + nsteps):
            sky.update_positions(ctx.time_at(k), ctx.loc)
            if can_async:
                counts.append(sky.coherency_calc(freqs=freqs, out=out_buf, sync=False)[1])
            else:
                o = sky.coherency_calc(freqs=freqs, out=half_buf)
                counts.append(o.shape[3])
                del o
        sky._timing_events = None
        return counts

    run_steps(0, warmup, None)
    ctx.barrier()
    events = []
    launches0 = ctx._lib.LAUNCHES[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ctx.stream)
    counts = run_steps(warmup, steps, events)
    e1.record(ctx.stream)
    ctx.barrier()
    launches = ctx._lib.LAUNCHES[0] - launches0
    if can_async:
        counts = [int(c) for c in torch.stack(counts).flatten().cpu()]
    ms = e0.elapsed_time(e1)
    k2 = sum(a.elapsed_time(b) for a, b in events) / max(1, len(events))
    k2_bytes = sum(k2_algorithmic_bytes(cat, c, stored_full) for c in counts) / max(1, len(counts))
    del out_buf, half_buf
    return dict(ms=ms, counts=counts, evals=sum(counts) * nf, k2_ms=k2, k2_bytes=k2_bytes, launches=launches,
                mode="async" if can_async else "sync")


def coherency_record(ctx, r, steps, kernel, unit=UNIT):
    """Reduce one measure_coherency result over the ranks into a config record."""
    achieved = r["k2_bytes"] / (r["k2_ms"] * 1e-3) / 1e9 if r["k2_ms"] > 0 else 0.0
    (ms, k2_ms, neg_ach), (evals, n_above, k2_bytes) = ctx.reduce(
        [r["ms"], r["k2_ms"], -achieved], [r["evals"], sum(r["counts"]) / max(1, len(r["counts"])), r["k2_bytes"]])
    ach = -neg_ach  # the slowest rank's rate
    return {
        "value": evals / (ms * 1e-3), "unit": unit, "ms_per_step": ms / steps, "steps": steps, "n_above": n_above,
        "scaling": "strong", "mode": r["mode"],
        "roofline": {"bound": "hbm", "kernel": kernel, "kernel_ms": k2_ms, "algorithmic_bytes": k2_bytes / ctx.world,
                     "achieved": ach, "peak": ctx.peak, "unit": "GB/s", "frac": ach / ctx.peak,
                     "kernel_share_of_step": k2_ms / (ms / steps),
                     "per": "GPU: bytes of one launch of one rank (mean over ranks) / slowest rank's mean kernel time"},
    }


HPX_SHARD_BLOCK = 4096  # pixels per block of the block-cyclic HEALPix split


def make_sky(ctx, workload, seed=2):
    _, n, nf, stype, frame = WORKLOADS[workload]
    if workload == "c3" and ctx.world > 1:
        # A ring-ordered map split into contiguous ranges is a split by latitude: at HERA the southern ranks would hold
        # three times as many above-horizon pixels as the northern ones.  Deal the pixels out in blocks instead (rank r
        # takes blocks r, r + N, ...): every rank sees every latitude.  hpx_inds names the pixels explicitly.
        cat = synth_catalog(workload, seed)
        blocks = np.arange(n // HPX_SHARD_BLOCK)
        mine = blocks[blocks % ctx.world == ctx.rank]
        pix = (mine[:, None] * HPX_SHARD_BLOCK + np.arange(HPX_SHARD_BLOCK)[None, :]).ravel()
        for key in ("stokes", "spectral_index", "reference_frequency"):
            cat[key] = np.ascontiguousarray(cat[key][..., pix])
        cat["n"] = pix.size
        kw = model_kwargs(cat)
        kw.pop("hpx_first")
        kw["hpx_inds"] = pix
        return ctx.rsk.SkyModel(**kw, device=ctx.dev), cat
    lo, hi = shard_bounds(n, ctx.world, ctx.rank)
    cat = synth_catalog(workload, seed, lo, hi)
    sky = ctx.rsk.SkyModel(**model_kwargs(cat), device=ctx.dev)
    return sky, cat


def run_config(ctx, name, steps, warmup):
    """One entry of the ``configs`` record."""
    torch = ctx.torch
    if name == "c1":
        # too small to shard: rank 0 runs it, eagerly and as a CUDA graph (launch-bound either way)
        rec = None
        if ctx.rank == 0:
            cat = synth_catalog("c1", 2)
            sky = ctx.rsk.SkyModel(**model_kwargs(cat), device=ctx.dev)
            w = ctx.world
            ctx.world = 1
            try:
                r = measure_coherency(ctx, sky, cat, cat["freqs"], max(steps, 20), warmup)
                rec = coherency_record(ctx, r, max(steps, 20), "k_eval<flat, coherency>")
                graph = sky.capture_step(freqs=cat["freqs"])
                for k in range(warmup):
                    graph.run(ctx.time_at(k), ctx.loc)
                torch.cuda.synchronize()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                cnts = []
                g0.record(ctx.stream)
                for k in range(warmup, warmup + 50):
                    cnts.append(graph.run(ctx.time_at(k), ctx.loc)[1].clone())
                g1.record(ctx.stream)
                torch.cuda.synchronize()
                gms = g0.elapsed_time(g1)
                rec["graph"] = {"value": int(torch.stack(cnts).sum().item()) / (gms * 1e-3), "unit": UNIT,
                                "ms_per_step": gms / 50, "what": "SkyModel.capture_step: the step replayed as one CUDA graph"}
                rec["scaling"] = "rank 0 only (1000 components)"
                rec["roofline"]["note"] = "launch-latency scale: 64 KB written per step"
            finally:
                ctx.world = w
        ctx.barrier()
        return rec
    if name == "c5":
        return run_sweep(ctx, steps, warmup)
    sky, cat = make_sky(ctx, name)
    freqs = cat["freqs"]
    stored = False
    kernel = f"k_eval<{cat['spectral_type']}, coherency>"
    if name == "full":
        sky.at_frequencies(freqs, run_check=False)  # K3: the (4, 1024, n) cube, on the device
        freqs, stored = None, True
        kernel = "k_eval<full, coherency>"
    if name == "c4":
        kernel = "k_eval<subband (banded), coherency>"
    r = measure_coherency(ctx, sky, cat, freqs, steps, warmup, stored_full=stored)
    rec = coherency_record(ctx, r, steps, kernel)
    rec["components"] = WORKLOADS[name][1]
    rec["nfreqs"] = WORKLOADS[name][2]
    rec["workload"] = WORKLOADS[name][0]
    del sky
    ctx.free()
    return rec


def run_sweep(ctx, steps, warmup):
    """BASELINE configs[4]: 1M sources x 1000 contexts in one launch per step (K1 sweep), sources sharded."""
    torch = ctx.torch
    desc, n_total, nt = C5
    lo, hi = shard_bounds(n_total, ctx.world, ctx.rank)
    bsize = -(-n_total // NBLOCKS)
    lon, lat = [], []
    for b in range(lo // bsize, -(-hi // bsize)):
        rng = np.random.default_rng([5, b])
        nb = min(n_total, (b + 1) * bsize) - b * bsize
        lon.append(rng.uniform(0, 2 * np.pi, nb))
        lat.append(np.arcsin(rng.uniform(-1, 1, nb)))
    lon = np.concatenate(lon)[lo - (lo // bsize) * bsize:][:hi - lo]
    lat = np.concatenate(lat)[lo - (lo // bsize) * bsize:][:hi - lo]
    n = hi - lo
    st = np.zeros((4, 1, n))
    st[0] = 1.0
    sky = ctx.rsk.SkyModel(ra=lon, dec=lat, frame="icrs", stokes=st, spectral_type="flat", device=ctx.dev)
    t0 = ctx.rsk.Time.from_utc("2015-03-01 00:00:00")
    th = time.perf_counter()
    astroms = ctx.rsk.astrometry.astrom_contexts([t0 + k * 86.164 / 86400.0 for k in range(nt)], ctx.loc)
    host_ctx_s = time.perf_counter() - th
    ctx_dev = sky.sweep_contexts(None, ctx.loc, astroms=astroms)  # packed and uploaded once (272 KB)
    reps = max(3, min(steps, 5))
    kev = []
    best = 1e30
    for r in range(warmup + reps):
        ctx.barrier()
        sky._timing_events = kev if r >= warmup else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ctx.stream)
        alt_az, _, above = sky.positions_sweep(None, ctx.loc, contexts=ctx_dev)
        e1.record(ctx.stream)
        torch.cuda.synchronize()
        (ms,), _ = ctx.reduce([e0.elapsed_time(e1)], [0.0])
        if r >= warmup:
            best = min(best, ms)
        if r < warmup + reps - 1:
            del alt_az, above
    sky._timing_events = None
    kernel_ms = min(a.elapsed_time(b) for a, b in kev)
    ach = n * nt * 17 / (kernel_ms * 1e-3) / 1e9
    (kmax, neg), _ = ctx.reduce([kernel_ms, -ach], [0.0])
    rec = {
        "value": n_total * nt / (best * 1e-3), "unit": "source.time position evals/s", "ms_per_step": best, "steps": reps,
        "components": n_total, "times": nt, "scaling": "strong", "workload": desc,
        "roofline": {"bound": "fp64 pipe (see DESIGN.md K1 sweep); reported against hbm", "kernel": "k_positions_sweep",
                     "kernel_ms": kmax, "algorithmic_bytes": n * nt * 17, "achieved": -neg, "peak": ctx.peak,
                     "unit": "GB/s", "frac": -neg / ctx.peak, "kernel_share_of_step": kmax / best,
                     "per": "GPU; 17 B per (source, time): alt, az, mask"},
        "host_context_build_ms": host_ctx_s * 1e3,
    }
    # this kernel is bound by fp64 issue, not by HBM (profiles/r02_sweep_floor.md): 67 fp64-pipe instructions per
    # evaluation in the SASS of the time loop, 2 dispatch cycles each on the 4 x 148 schedulers
    sm_hz = 1.965e9
    pipe_ms = (n * nt / 32.0) / (4 * 148) * 67 * 2 / sm_hz * 1e3
    rec["fp64_pipe"] = {"fp64_instr_per_eval": 67, "pipe_floor_ms": pipe_ms, "frac_of_pipe_floor": pipe_ms / kmax,
                        "hbm_frac_at_pipe_floor": n * nt * 17 / (pipe_ms * 1e-3) / 1e9 / ctx.peak,
                        "note": "floor = fp64 instructions alone at 16 lanes per scheduler and 1965 MHz, per GPU; even at "
                                "the floor the kernel would sit at 0.72 of the 17-byte HBM roofline (0.63 with the 77 instructions of the "
                                "full-accuracy-margin routines of round 1)"}
    if ctx.rank == 0:
        k = nt // 3
        ref_alt_az, _ = positions_cpu_baseline(lon[:2000], lat[:2000], astroms[k:k + 1])
        got = alt_az[k][:, :2000].cpu().numpy()
        rec["max_alt_err_uas_vs_oracle"] = float(np.abs(got[0] - ref_alt_az[0]).max() / (np.pi / 180 / 3600 * 1e-6))
    del alt_az, above, sky
    ctx.free()
    return rec


# ---- end to end ------------------------------------------------------------------------------------
def host_ceilings(ctx, nthreads, n_above, nf):
    """What the box can do, measured now, all ranks at once: pinned device->host copy rate and the rate at which this
    rank's host threads expand Hermitian-packed slabs into the reference-layout array."""
    torch = ctx.torch
    nbytes = 1 << 30
    d = torch.empty(nbytes, dtype=torch.uint8, device=ctx.dev)
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h.copy_(d, non_blocking=True)
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ctx.stream)
    for _ in range(3):
        h.copy_(d, non_blocking=True)
    e1.record(ctx.stream)
    ctx.barrier()
    d2h = 3 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
    del d, h
    lib = ctx._lib.load()
    slab = 64
    ld = -(-n_above // 32) * 32
    packed = torch.zeros((4, slab, ld), dtype=torch.float64).pin_memory()
    out = torch.zeros(4 * slab * n_above, dtype=torch.complex128)
    lib.rsk_unpack_coherency_host(packed.data_ptr(), slab, n_above, ld, out.data_ptr(), slab, 0, n_above, nthreads)
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        lib.rsk_unpack_coherency_host(packed.data_ptr(), slab, n_above, ld, out.data_ptr(), slab, 0, n_above, nthreads)
    expand = 3 * 64.0 * slab * n_above / (time.perf_counter() - t0) / 1e9
    # both at once, free-running (no dependency between them): what the host memory system sustains when the DMA
    # writes of the copies compete with the expansion's reads and non-temporal writes
    dslab = torch.zeros((4, slab, ld), dtype=torch.float64, device=ctx.dev)
    hstage = [torch.empty((4, slab, ld), dtype=torch.float64).pin_memory() for _ in range(2)]
    stop, copied = threading.Event(), [0]

    def copier():
        torch.cuda.set_device(ctx.dev)
        s = torch.cuda.Stream(ctx.dev)
        with torch.cuda.stream(s):
            while not stop.is_set():
                for hbuf in hstage:
                    hbuf.copy_(dslab, non_blocking=True)
                s.synchronize()
                copied[0] += 2

    ctx.barrier()
    th = threading.Thread(target=copier)
    th.start()
    time.sleep(0.03)
    c0, t0 = copied[0], time.perf_counter()
    nrep = 6
    for _ in range(nrep):
        lib.rsk_unpack_coherency_host(packed.data_ptr(), slab, n_above, ld, out.data_ptr(), slab, 0, n_above, nthreads)
    dt = time.perf_counter() - t0
    c1 = copied[0]
    stop.set()
    th.join()
    ctx.barrier()
    conc_expand = nrep * 64.0 * slab * n_above / dt / 1e9
    conc_d2h = (c1 - c0) * dslab.numel() * 8 / dt / 1e9
    # host DRAM traffic the two generate together: DMA writes + (32 B read + 64 B written per 64 B expanded)
    traffic = conc_d2h + 1.5 * conc_expand
    del dslab, hstage, packed, out
    (_, _), (d2h_sum, expand_sum, traffic_sum) = ctx.reduce([0.0, 0.0], [d2h, expand, traffic])
    return {"d2h_GBps_per_gpu": d2h, "d2h_GBps_all_ranks": d2h_sum, "expand_GBps_out_this_rank": expand,
            "expand_GBps_out_all_ranks": expand_sum, "unpack_threads_per_rank": nthreads,
            "concurrent_d2h_GBps_this_rank": conc_d2h, "concurrent_expand_GBps_out_this_rank": conc_expand,
            "host_dram_traffic_GBps_this_rank": traffic, "host_dram_traffic_GBps_all_ranks": traffic_sum,
            "how": "1 GiB pinned D2H copies x3 on every rank at once; rsk_unpack_coherency_host on a 64-channel slab x3 on "
                   "every rank at once (no copies in flight); then both at once, free-running, which gives the host DRAM "
                   "traffic the box sustains (DMA writes + 1.5 x expanded bytes)"}


def measure_e2e(ctx, workload, counts_hint, e2e_steps):
    """The metric through the public API with HOST buffers; returns the ``e2e`` record (rank 0) or None."""
    torch = ctx.torch
    _, n_total, nf, stype, frame = WORKLOADS[workload]
    lo, hi = shard_bounds(n_total, ctx.world, ctx.rank)
    cat = synth_catalog(workload, 2, lo, hi)
    freqs = cat["freqs"]
    n = hi - lo
    kw0 = model_kwargs(cat)
    host_in = {k: torch.as_tensor(np.ascontiguousarray(v)).pin_memory() for k, v in kw0.items()
               if isinstance(v, np.ndarray) and k != "freq_array"}
    h2d_bytes = sum(t.numel() * t.element_size() for t in host_in.values()) + freqs.nbytes
    n_above_max = int(min(n, 1.1 * max(counts_hint) + 1024))
    need = 4 * nf * n_above_max
    # Destination: host memory.  Two routes give the same array: "packed" (32 B/eval over PCIe + host-thread
    # expansion) and "plain" (64 B/eval over PCIe straight into the pinned array, no host work).
    try:
        host_out = torch.empty(need, dtype=torch.complex128).pin_memory()
        pinned_out = True
    except RuntimeError:
        host_out = torch.empty(need, dtype=torch.complex128)
        pinned_out = False
    host_out.zero_()
    nthreads = int(os.environ.get("RSK_E2E_THREADS", max(1, min(32, ctx.cpus - (1 if ctx.cpus > 4 else 0)))))
    slab = int(os.environ.get("RSK_E2E_SLAB", 64))
    ceil = host_ceilings(ctx, nthreads, int(max(counts_hint)), nf)
    state = {"evals": 0, "d2h": 0.0}

    def e2e_step(k, record, packed):
        kw = dict(kw0)
        kw.update({key: t.to(ctx.dev, non_blocking=True) for key, t in host_in.items()})
        s = ctx.rsk.SkyModel(**kw, device=ctx.dev)
        s.update_positions(ctx.time_at(k), ctx.loc)
        view = s.coherency_calc_to_host(freqs, host_out, slab=slab, packed=packed, nthreads=nthreads)
        if packed is False:
            torch.cuda.synchronize()
        if record:
            ev = view.shape[2] * view.shape[3]
            state["evals"] += ev
            if packed == "auto":
                npk, npl = s._e2e_route_counts
                state["d2h"] += ev * (32 * npk + 64 * npl) / max(1, npk + npl)
            else:
                state["d2h"] += ev * (32 if packed else 64)
        return view

    routes = (True, False, "auto") if pinned_out else (True,)
    trial = {}
    for packed in routes:
        e2e_step(0, False, packed)
        if packed == "auto":
            e2e_step(0, False, packed)
        best = float("inf")
        for _ in range(2):  # the hosts of the pool are shared: the better of two trials per route
            ctx.barrier()
            t0 = time.perf_counter()
            e2e_step(0, False, packed)
            ctx.barrier()
            best = min(best, time.perf_counter() - t0)
        trial[packed] = best
    choice, _ = ctx.reduce([trial[r] for r in routes], [0.0])
    use = routes[int(np.argmin(choice))]
    label = {"auto": "mixed", True: "packed", False: "plain"}
    ctx.barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        e2e_step(1 + k, True, use)
    ctx.barrier()
    e2e_s = time.perf_counter() - t0
    (e2e_s,), (evals, d2h) = ctx.reduce([e2e_s], [state["evals"], state["d2h"]])
    # the fused consumer: the source sum of the apparent coherency comes back, a few KB per step
    vis_s = float("nan")
    vis_evals = 0
    try:
        def vis_step(k):
            kw = dict(kw0)
            kw.update({key: t.to(ctx.dev, non_blocking=True) for key, t in host_in.items()})
            s = ctx.rsk.SkyModel(**kw, device=ctx.dev)
            s.update_positions(ctx.time_at(k), ctx.loc)
            v = s.visibility_calc(beam="short_dipole", freqs=freqs).cpu()
            return s._last_n_above * nf, v
        vis_step(0)
        ctx.barrier()
        t0 = time.perf_counter()
        for k in range(e2e_steps):
            ev, _ = vis_step(1 + k)
            vis_evals += ev
        ctx.barrier()
        vis_s = time.perf_counter() - t0
    except Exception as exc:  # the consumer is an extra; never lose the e2e record over it
        print(f"[bench] visibility e2e skipped: {exc}", file=sys.stderr)
    (vis_s,), (vis_evals,) = ctx.reduce([vis_s], [vis_evals])
    del host_out, host_in
    ctx.free()
    if ctx.rank != 0:
        return None
    ms_step = e2e_s / e2e_steps * 1e3
    ev_step_rank = evals / e2e_steps / ctx.world
    wire = (d2h / max(evals, 1))  # bytes per evaluation that crossed PCIe
    link_ms = ev_step_rank * wire / (ceil["d2h_GBps_per_gpu"] * 1e9) * 1e3
    host_ms = ev_step_rank * 64.0 / (ceil["expand_GBps_out_this_rank"] * 1e9) * 1e3 if use is not False else 0.0
    # host DRAM bytes per evaluation of the chosen route: the DMA write of what crosses PCIe, plus 32 B read and 64 B
    # written by the expansion for the packed share
    packed_share = (64.0 - wire) / 32.0 if use is not False else 0.0
    dram_per_eval = wire + 96.0 * packed_share
    dram_ms = ev_step_rank * dram_per_eval / (ceil["host_dram_traffic_GBps_this_rank"] * 1e9) * 1e3
    ceil.update({"link_floor_ms_per_step": link_ms, "host_expand_floor_ms_per_step": host_ms,
                 "frac_of_link_floor": link_ms / ms_step, "frac_of_host_floor": host_ms / ms_step if host_ms else None,
                 "frac_of_binding_floor": max(link_ms, host_ms) / ms_step,
                 "host_dram_model": {"bytes_per_eval": dram_per_eval, "floor_ms_per_step": dram_ms,
                                     "frac": dram_ms / ms_step,
                                     "what": "estimate, not a bound: host DRAM bytes of the route (DMA writes + 32 B read and 64 B "
                                             "written per expanded evaluation) over the DRAM traffic measured with copies and "
                                             "expansion free-running side by side"},
                 "note": "floors per rank, each measured with all ranks active at once: bytes over PCIe / concurrent D2H rate "
                         "(on these hosts the ranks share one host bridge: the rate per GPU falls from 57 GB/s alone to 35 at "
                         "2, 18 at 4 and 11-15 at 8 ranks); 64 B per evaluation / expansion rate (packed and mixed routes). "
                         "The two share the host memory system (side by side each runs at 0.5-0.8 of its rate alone), which is "
                         "why a packed step sits between the floors and their sum"})
    return {
        "value": evals / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes * ctx.world,
        "d2h_bytes_per_step": d2h / e2e_steps, "steps": e2e_steps, "ms_per_step": ms_step,
        "route": label[use], "route_trials_ms": {label[r]: c * 1e3 for r, c in zip(routes, choice)},
        "host_out": "pinned" if pinned_out else "pageable", "host_cores_per_rank": ctx.cpus,
        "roofline": ceil,
        "visibility": ({"value": vis_evals / vis_s, "unit": UNIT, "ms_per_step": vis_s / e2e_steps * 1e3,
                        "d2h_bytes_per_step": 64 * nf * ctx.world,
                        "what": "fused consumer, reported beside the coherency e2e: host inputs -> H2D -> SkyModel -> "
                                "update_positions -> visibility_calc (zero-baseline source sum of J C J^H, short dipole) -> "
                                "(1,2,2,Nf) complex128 back to the host; evaluations = above-horizon components x Nf summed"}
                       if vis_s == vis_s else None),
        "how": "pinned host inputs -> H2D -> SkyModel -> update_positions -> coherency_calc_to_host into the "
               "(2,2,Nf,N_above) complex128 array in host memory; route 'packed' = Hermitian-packed slabs "
               "(32 B/eval over PCIe) expanded by host threads, 'plain' = 64 B/eval copied as is, 'mixed' = per slab "
               "whichever finishes earlier given the measured host rate; the fastest of these (best of two untimed "
               "trials each) is used for the timed steps",
    }


def measure_gather(ctx, steps):
    """Assembly of the gathered (2,2,Nf,sum N_above) array on the c4 split, timed apart from the step."""
    torch, dist = ctx.torch, ctx.dist
    from pyradiosky_b200.dist import PeerGather, exchange_counts, gather_columns

    name = "c4"
    nf_g = 256  # 1M sources x 256 channels: 8.2 GB gathered (every GPU holds the whole array in the all-gather case)
    sky, cat = make_sky(ctx, name)
    freqs = cat["freqs"][::4][:nf_g]
    sky.update_positions(ctx.time_at(3), ctx.loc)
    _, n_above = sky.compact_above_horizon()
    counts, col0, total = exchange_counts(n_above, ctx.dev)
    nbytes = 64.0 * len(freqs) * total
    rec = {"workload": f"c4 split over {ctx.world} GPUs, {len(freqs)} of its 1024 channels", "columns": total,
           "bytes_gathered": nbytes}
    # peer peak, measured now: every rank copies 1 GiB into its neighbour's HBM at the same time
    try:
        import torch.distributed._symmetric_memory as symm_mem

        pbuf = symm_mem.empty(1 << 27, dtype=torch.float64, device=ctx.dev)
        hdl = symm_mem.rendezvous(pbuf, dist.group.WORLD)
        nxt = hdl.get_buffer((ctx.rank + 1) % ctx.world, (1 << 27,), torch.float64)
        src = torch.zeros(1 << 27, dtype=torch.float64, device=ctx.dev)
        nxt.copy_(src)
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ctx.stream)
        for _ in range(4):
            nxt.copy_(src)
        e1.record(ctx.stream)
        ctx.barrier()
        (pms,), _ = ctx.reduce([e0.elapsed_time(e1)], [0.0])
        rec["peer_copy_GBps_per_gpu"] = 4 * (1 << 30) / (pms * 1e-3) / 1e9
        del pbuf, hdl, nxt, src
    except Exception as exc:
        rec["peer_copy_GBps_per_gpu"] = None
        rec["peer_copy_error"] = str(exc)[:200]
    peak = rec["peer_copy_GBps_per_gpu"] or 770.0

    def timed(fn):
        fn()
        best = 1e30
        for _ in range(max(2, min(steps, 3))):
            ctx.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ctx.stream)
            fn()
            e1.record(ctx.stream)
            ctx.barrier()
            (ms,), _ = ctx.reduce([e0.elapsed_time(e1)], [0.0])
            best = min(best, ms)
        return best

    try:
        # (a rank's launch is sized for all of its components -- the count stays on the device -- so the buffer needs
        # room for the columns before it plus its whole block)
        (nmax,), _ = ctx.reduce([sky.Ncomponents], [0.0])
        pg = PeerGather(len(freqs), total + int(nmax), device=ctx.dev, streams=2)
        recv_all = nbytes * (ctx.world - 1) / ctx.world  # bytes each GPU receives from its peers
        for route, what in (("copy", "kernel writes its columns locally in up to 4 column slabs (>= 32k columns each); copy "
                                     "engines push each finished slab to the peers while the next slab is computed"),
                            ("store", "one kernel launch per destination, the kernel's own stores travel over NVLink")):
            ms_all = timed(lambda: pg.coherency_calc(sky, freqs=freqs, route=route))
            ms_one = timed(lambda: pg.coherency_calc(sky, freqs=freqs, dst=0, route=route))
            rec[f"peer_{route}_all_gather"] = {"ms": ms_all, "GBps_received_per_gpu": recv_all / (ms_all * 1e-3) / 1e9,
                                               "frac_of_peer_copy": recv_all / (ms_all * 1e-3) / 1e9 / peak, "what": what}
            rec[f"peer_{route}_to_rank0"] = {"ms": ms_one, "GBps_received": recv_all / (ms_one * 1e-3) / 1e9,
                                             "frac_of_peer_copy": recv_all / (ms_one * 1e-3) / 1e9 / peak}
        del pg
    except Exception as exc:
        rec["peer_store_error"] = str(exc)[:300]
    ctx.free()
    try:
        def nccl():
            local = sky.coherency_calc(freqs=freqs, pad=False)
            return gather_columns(local, counts)
        ms_nccl = timed(nccl)
        recv_all = nbytes * (ctx.world - 1) / ctx.world
        rec["nccl_all_gather_and_repack"] = {"ms": ms_nccl, "GBps_received_per_gpu": recv_all / (ms_nccl * 1e-3) / 1e9,
                                             "frac_of_peer_copy": recv_all / (ms_nccl * 1e-3) / 1e9 / peak}
    except Exception as exc:
        rec["nccl_error"] = str(exc)[:300]
    local_ms = timed(lambda: sky.coherency_calc(freqs=freqs, pad=False))
    rec["local_compute_ms"] = local_ms
    rec["note"] = ("frac_of_peer_copy: received bytes per GPU / time, over the neighbour-to-neighbour copy rate measured in "
                   "this run (copy engines, all ranks at once); the peer-store route is the coherency kernel itself "
                   "writing into the destination's HBM")
    del sky
    ctx.free()
    return rec


def run_gpu(args):
    ctx = Ctx(args)
    torch = ctx.torch
    workload = args.workload
    desc, n_total, nf, stype, frame = WORKLOADS[workload]
    sky, cat = make_sky(ctx, workload)
    freqs = cat["freqs"]
    stored = workload == "full"
    if stored:
        sky.at_frequencies(freqs, run_check=False)
        freqs = None
    sampler = ClockSampler(ctx.local_rank if ctx.world > 1 else 0)
    with sampler:
        r = measure_coherency(ctx, sky, cat, freqs, args.steps, args.warmup, stored_full=stored)
    kernel = "k_eval<full, coherency>" if stored else f"k_eval<{stype}, coherency>"
    head = coherency_record(ctx, r, args.steps, kernel)
    counts = r["counts"]
    # the reference-shaped call (exactly sized result: one host read of the count per step) and the CUDA-graph step
    sync_api = graph_api = None
    if r["mode"] == "async" and not args.quick:
        def sync_steps(first, nsteps):
            c = 0
            for k in range(first, first + nsteps):
                sky.update_positions(ctx.time_at(k), ctx.loc)
                o = sky.coherency_calc(freqs=freqs)
                c += o.shape[3]
                del o
            return c
        sync_steps(0, args.warmup)
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ctx.stream)
        c = sync_steps(args.warmup, args.steps)
        e1.record(ctx.stream)
        ctx.barrier()
        (ms,), (ev,) = ctx.reduce([e0.elapsed_time(e1)], [c * nf])
        sync_api = {"value": ev / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / args.steps,
                    "what": "coherency_calc(freqs) returning an exactly sized array: one host read of the count per step"}
        ctx.free()
        graph = sky.capture_step(freqs=freqs)
        for k in range(args.warmup):
            graph.run(ctx.time_at(k), ctx.loc)
        ctx.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gcounts = []
        g0.record(ctx.stream)
        for k in range(args.warmup, args.warmup + args.steps):
            gcounts.append(graph.run(ctx.time_at(k), ctx.loc)[1].clone())  # (no name may keep graph.out alive)
        g1.record(ctx.stream)
        ctx.barrier()
        (ms,), (ev,) = ctx.reduce([g0.elapsed_time(g1)], [int(torch.stack(gcounts).sum().item()) * nf])
        graph_api = {"value": ev / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / args.steps,
                     "what": "SkyModel.capture_step(freqs): the whole step (positions, compaction, coherency) "
                             "replayed as one CUDA graph after a 352-byte host->device copy"}
        del graph
    del sky
    ctx.free()

    # ---- every BASELINE config, sharded like the headline ------------------------------------------------
    configs = {}
    if args.configs != "none":
        names = ["c1", "c2", "c3", "c4", "c5", "full"] if args.configs == "all" else args.configs.split(",")
        csteps = max(3, min(args.steps, args.config_steps))
        for name in names:
            try:
                configs[name] = run_config(ctx, name, csteps, max(3, min(args.warmup, 3)))
            except Exception as exc:  # one config must not take the line down
                configs[name] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
                ctx.free()
                ctx.barrier()

    # ---- end to end through the public API with host buffers ---------------------------------------------
    e2e = None
    if args.e2e_steps > 0:
        e2e = measure_e2e(ctx, "c2" if stored else workload, counts, max(2, min(args.steps, args.e2e_steps)))

    gather = None
    if ctx.world > 1 and not args.no_gather:
        try:
            gather = measure_gather(ctx, args.steps)
        except Exception as exc:
            gather = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    if ctx.rank == 0:
        cfg = config_record(workload, ctx.world)
        tr = traffic_record(workload)
        if ctx.world > 1 and tr["traffic"] is not None:  # a rank launches on 1/N of the catalogue
            tr = {"traffic": tr["traffic"] / ctx.world, "traffic_source": tr["traffic_source"] + f", divided by {ctx.world} ranks"}
        head["roofline"].update({"peak_source": ctx.peak_src, **tr})
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": ctx.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg, "n_above": head["n_above"], "headline_mode": r["mode"],
            "roofline": head["roofline"],
            "sync_api": sync_api, "graph_api": graph_api,
            "configs": configs,
            "e2e": e2e,
            "gather": gather,
            "gpu_launches": r["launches"],
            "gpu_launches_how": "kernel launches of librsk counted by the ctypes layer during the timed headline region on "
                                "rank 0 (per step: k_positions 1, k_compact_* 3, k_eval 1)",
            "clocks": sampler.summary(),
        }
        if ctx.world == 1 and not args.no_cpu_baseline:
            n_sample = min(n_total, args.cpu_sample)
            v, s_per = time_oracle("c2" if stored else workload, n_sample, 1, 0, 1)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": f"one step on {n_sample} of {n_total} components x {nf} frequencies ({s_per:.1f} s); restated "
                          "oracle (numpy/scipy; astropy absent so pyradiosky itself cannot run), per-source loops kept"}
        emit(line)
    if ctx.world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


def _claim_stdout():
    """The contract is ONE JSON line on stdout, but libraries (e.g. the NCCL version banner) also print
    there.  Point fd 1 at stderr for the duration of the run and keep the real stdout for the result."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


_RESULT_OUT = None


def emit(line):
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _RESULT_OUT
    _RESULT_OUT = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="n1m", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--configs", default="all", help="'all', 'none' or a comma list of c1,c2,c3,c4,c5,full")
    ap.add_argument("--config-steps", type=int, default=8)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--quick", action="store_true", help="skip the sync / graph variants of the headline")
    ap.add_argument("--no-gather", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=20000, help="components in the cpu_baseline sample")
    ap.add_argument("--ref-sample-per-core", type=int, default=2500)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
