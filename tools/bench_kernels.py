#!/usr/bin/env python
"""Throughput of the secondary kernels against their HBM rooflines (numbers for DESIGN.md section 3).

K1 positions (57 B/component), K1b compaction, polarised mask, HEALPix centres, K3 at_frequencies
(32 B/eval), K2 in 'full' mode = the reference's own call sequence at_frequencies() once, then
update_positions() + coherency_calc() per time step (96 B/eval algorithmic), stokes_to_coherency
(96 B/element), K4 unit conversion (64 B/element), spline coefficients.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timeit(fn, reps=5):
    import torch

    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    import torch

    import __graft_entry__ as entry

    entry.build()
    from pyradiosky_b200 import EarthLocation, SkyModel, Time, _lib

    peak = 6547.8
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    rng = np.random.default_rng(7)
    loc = EarthLocation.from_dms((-30, 43, 17.5), (21, 25, 41.9), 1073.0)
    t0 = Time.from_utc("2015-03-01 00:00:00")
    out = {}

    def rec(name, ms, nbytes, note=""):
        out[name] = {"ms": ms, "GBps": nbytes / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / peak,
                     "algorithmic_bytes": nbytes, "note": note}

    # --- K1 / K1b / polarised mask on 16M point components -------------------------------------------------
    n = 16_000_000
    lon = torch.rand(n, dtype=torch.float64, device="cuda") * (2 * np.pi)
    lat = torch.asin(torch.rand(n, dtype=torch.float64, device="cuda") * 2 - 1)
    st = torch.zeros((4, 1, n), dtype=torch.float64, device="cuda")
    st[0] = 1.0
    st[1] = 0.1
    sky = SkyModel(lon=lon, lat=lat, stokes=st, spectral_type="flat")
    k = [0]

    from pyradiosky_b200 import astrometry

    # contexts built beforehand so that the events bracket GPU work, not the host-side context builder
    ctxs = [astrometry.astrom_context(t0 + 1e-3 * j, loc) for j in range(1, 40)]

    def upd():
        k[0] += 1
        sky.update_positions(t0 + 1e-3 * k[0], loc, astrom=ctxs[k[0] % len(ctxs)])

    rec("k_positions_16M_points(+rot)", timeit(upd), n * (57 + 24), "57 B/component + 24 B rotation coefficients")
    rec("k_compact_16M", timeit(lambda: sky._compact_async()), n * 1 * 2 + 8 * n // 2, "mask read twice + idx written")
    pol = torch.empty(n, dtype=torch.uint8, device="cuda")
    rec("k_polarized_16M", timeit(lambda: sky._call("rsk_polarized_mask", _lib.ptr(sky.stokes), 1, n, _lib.ptr(pol))),
        n * 25, "24 B read + 1 B written")
    del sky, lon, lat, st, pol
    torch.cuda.empty_cache()

    # --- HEALPix nside 1024: positions with on-device centres, and get_lon_lat ------------------------------
    nside = 1024
    npix = 12 * nside * nside
    st = torch.ones((4, 1, npix), dtype=torch.float64, device="cuda")
    st[1:] = 0
    for order in ("ring", "nested"):
        hp = SkyModel(nside=nside, hpx_order=order, stokes=st, spectral_type="flat", component_type="healpix",
                      stokes_unit="K")
        k2 = [0]

        def upd_h():
            k2[0] += 1
            hp.update_positions(t0 + 1e-3 * k2[0], loc, astrom=ctxs[k2[0] % len(ctxs)])

        rec(f"k_positions_healpix_nside1024_{order}", timeit(upd_h), npix * 41, "41 B/pixel, centres generated on device")
        rec(f"k_healpix_lonlat_nside1024_{order}", timeit(lambda: hp.get_lon_lat()), npix * 16)
        del hp
    del st
    torch.cuda.empty_cache()

    # --- spectra: K3, K2 'full', stokes_to_coherency, K4 ------------------------------------------------------
    n, nf = 300_000, 1024
    lon = rng.uniform(0, 2 * np.pi, n)
    lat = np.arcsin(rng.uniform(-1, 1, n))
    i = 10 ** rng.uniform(-2, 1, n)
    s1 = np.zeros((4, 1, n))
    s1[0, 0] = i
    s1[1, 0] = 0.05 * i
    s1[2, 0] = -0.02 * i
    s1[3, 0] = 0.01 * i
    freqs = np.linspace(100e6, 200e6, nf)
    sky = SkyModel(lon=lon, lat=lat, frame="fk5", stokes=s1, spectral_type="spectral_index",
                   spectral_index=np.clip(rng.normal(-0.8, 0.2, n), -2, 1), reference_frequency=np.full(n, 200e6))
    # run_check=False: the kernel alone (the reference's default run_check=True adds SkyModel.check: three scans of the new
    # cube for NaNs / negative Stokes I, measured on the next line)
    rec("at_frequencies 300k x 1024 with run_check (3 acceptability scans of the 9.8 GB cube, host syncs)",
        timeit(lambda: sky.at_frequencies(freqs, inplace=False)), 32 * freqs.size * n, "32 B/eval written by K3; the scans read 9.8 + 2.5 GB more")
    rec("k_eval<spectral_index,stokes> at_frequencies 300k x 1024", timeit(lambda: sky.at_frequencies(freqs, inplace=False, run_check=False)),
        32 * n * nf, "32 B/eval written")
    # N4: coherency_calc + Jones contraction fused (apparent coherency), analytic and caller-supplied beams
    sky.update_positions(t0, loc)
    na = int(sky.above_horizon.sum().item())
    rec("k_eval<spectral_index,apparent> short dipole 300k x 1024",
        timeit(lambda: sky.apparent_coherency_calc(beam="short_dipole", freqs=freqs)), 64 * na * nf, "64 B/eval written")
    j1 = torch.randn((2, 2, nf, na), dtype=torch.complex128, device="cuda")
    rec("k_eval<spectral_index,apparent> Jones arrays, one antenna 300k x 1024",
        timeit(lambda: sky.apparent_coherency_calc(jones=j1, freqs=freqs)), 128 * na * nf, "64 B/eval read + 64 written")
    j2 = torch.randn((2, 2, nf, na), dtype=torch.complex128, device="cuda")
    rec("k_eval<spectral_index,apparent> Jones arrays, two antennas 300k x 1024",
        timeit(lambda: sky.apparent_coherency_calc(jones=j1, jones2=j2, freqs=freqs)), 192 * na * nf,
        "128 B/eval read + 64 written")
    # source-summed consumer: nothing per-component is written; "bytes" = what the materialising route would move
    uvw = np.array([[14.6, 0.0, 0.0]])
    for label, kwv in (("short dipole, zero baseline", dict(beam="short_dipole")),
                       ("short dipole, 1 baseline", dict(beam="short_dipole", uvw=uvw)),
                       ("short dipole, 8 baselines", dict(beam="short_dipole", uvw=np.repeat(uvw, 8, axis=0) * np.arange(1, 9)[:, None])),
                       ("Jones arrays (1 antenna), 1 baseline", dict(jones=j1, uvw=uvw))):
        ms = timeit(lambda: sky.visibility_calc(freqs=freqs, **kwv))
        nb = 1 if "uvw" not in kwv else len(kwv["uvw"])
        out["k_visibility<spectral_index> " + label + " 300k x 1024"] = {
            "ms": ms, "evals_per_s": na * nf * nb / (ms * 1e-3), "GBps": 64 * na * nf / (ms * 1e-3) / 1e9,
            "frac_of_hbm_peak": 64 * na * nf / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": 64 * na * nf,
            "note": "compute-bound; GBps/frac = the 64 B/eval the materialised route would have written, for comparison"}
    del j1, j2
    torch.cuda.empty_cache()
    full = sky.at_frequencies(freqs, inplace=False)  # (4, 1024, 300k) = 9.8 GB, spectral_type 'full'
    full.update_positions(t0, loc)
    n_above = int(full.above_horizon.sum().item())
    ld = -(-n // 16) * 16
    buf = torch.empty((2, 2, nf, ld), dtype=torch.complex128, device="cuda")
    rec("k_eval<full,coherency> coherency_calc() on stored Stokes 300k x 1024",
        timeit(lambda: full.coherency_calc(out=buf, sync=False)), 96 * n_above * nf,
        "96 B/eval algorithmic; read from the component-major copy of the cube, 8 loads in flight per thread (the whole call: compaction launches + kernel)")
    del buf
    torch.cuda.empty_cache()
    rec("rsk_stokes_to_coherency 300k x 1024", timeit(lambda: full.calc_frame_coherency(store=False), reps=3), 96 * n * nf)
    rec("k_convert_units 300k x 1024 (Jy->K sr, per channel)",
        timeit(lambda: (full.jansky_to_kelvin(), full.kelvin_to_jansky()), reps=3) / 2, 64 * n * nf, "32 B read + 32 B written")
    del full, sky
    torch.cuda.empty_cache()

    # --- spline coefficients for 1M x 20 subbands -------------------------------------------------------------
    n = 1_000_000
    bands = np.array([76, 84, 92, 99, 107, 115, 122, 130, 143, 151, 158, 166, 174, 181, 189, 197, 204, 212, 220, 227]) * 1e6
    st = torch.rand((4, 20, n), dtype=torch.float64, device="cuda") + 1
    sb = SkyModel(lon=torch.zeros(n, dtype=torch.float64, device="cuda"), lat=torch.zeros(n, dtype=torch.float64, device="cuda"),
                  stokes=st, spectral_type="subband", freq_array=bands)

    def coef():
        sb.__dict__.pop("_subband_cache", None)
        sb._subband_operator("cubic")

    rec("k_spline_coefficients 1M x 20 (once per catalogue)", timeit(coef, reps=3), 2 * 640 * n, "640 B read + 640 B written per component")
    del sb, st
    torch.cuda.empty_cache()

    # --- select: predicate mask, compaction, gather on 4M x 20 subbands (GLEAM-like, 2.6 GB of Stokes) -----------
    n, nf = 4_000_000, 20
    lon = torch.rand(n, dtype=torch.float64, device="cuda") * (2 * np.pi)
    lat = torch.asin(torch.rand(n, dtype=torch.float64, device="cuda") * 2 - 1)
    st = torch.rand((4, nf, n), dtype=torch.float64, device="cuda") + 0.01
    sky = SkyModel(lon=lon, lat=lat, stokes=st, spectral_type="subband", freq_array=bands)
    sel = _lib.RskSelect()
    sel.use_lat, sel.lat_lo, sel.lat_hi = 1, -0.6, 0.9
    sel.use_min, sel.min_brightness, sel.non_nan, sel.non_negative = 1, 0.02, 1, 1
    keep = torch.empty(n, dtype=torch.uint8, device="cuda")
    coords = sky._coords_struct()
    import ctypes

    rec("k_select_mask 4M x 20 (lat range, min brightness, non_nan any, non_negative)",
        timeit(lambda: sky._call("rsk_select_mask", ctypes.byref(coords), _lib.ptr(st), nf, n, None, n, ctypes.byref(sel),
                                 _lib.ptr(keep))), n * (4 * nf * 8 + 8 + 1), "all four Stokes planes + latitude read, 1 B written")
    idx, m = sky._compact_mask(keep)
    idx = idx[:m].contiguous()
    outbuf = torch.empty((4, nf, m), dtype=torch.float64, device="cuda")
    rec(f"k_gather_rows stokes 4M -> {m} x 80 rows",
        timeit(lambda: sky._call("rsk_gather_rows", _lib.ptr(st), 4 * nf, n, _lib.ptr(idx), m, 8, _lib.ptr(outbuf))),
        m * (4 * nf * 16 + 8), "8 B read + 8 B written per kept element, index read once")
    rise = torch.empty(n, dtype=torch.float64, device="cuda")
    sett = torch.empty(n, dtype=torch.float64, device="cuda")
    rec("k_rise_set 4M", timeit(lambda: sky._call("rsk_rise_set_lsts", ctypes.byref(coords), n, -0.536, 0.04364,
                                                    _lib.ptr(rise), _lib.ptr(sett))), n * 32, "16 B read + 16 B written")

    def whole():
        s2 = SkyModel.__new__(SkyModel)
        s2.__dict__.update(sky.__dict__)
        s2.select(lat_range=np.array([-0.6, 0.9]), min_brightness=0.02, non_nan="any", non_negative=True)

    rec("SkyModel.select end to end 4M x 20 (mask + compaction + all gathers + host sync)", timeit(whole, reps=3),
        n * (4 * nf * 8 + 9) + m * (4 * nf * 16 + 8 + 32), "sum of the above")
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
