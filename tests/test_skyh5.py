"""SkyH5 loader: librsk's HDF5 reader/writer against the independent pure-Python
parser in oracle/h5_np.py, both pinned by the reference's own fixture files and the values its
tutorial prints for them (tests/golden/skyh5_known_answers.json).

CPU tests cover the host reader/writer through the C ABI; ``-m gpu`` tests cover the file -> HBM path
and ``SkyModel.from_skyh5`` / ``write_skyh5`` with the reference's warnings and error texts
(tests/test_skymodel.py:3851-4220, :4507-4522)."""

import json
import os
import subprocess
import sys
import warnings

import numpy as np
import pytest

from oracle import h5_np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
GSM = os.path.join(GOLDEN, "ref_fixture_gsm_icrs_nside8.skyh5")  # the reference's gsm_icrs.skyh5
OLD_POINT = os.path.join(GOLDEN, "ref_fixture_gleam50_old_header.skyh5")  # its old_skyh5_point_sources.skyh5


@pytest.fixture(scope="module")
def answers():
    with open(os.path.join(GOLDEN, "skyh5_known_answers.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="module")
def skyh5(librsk):
    from pyradiosky_b200 import skyh5 as mod

    return mod


# ---------------------------------------------------------------------------------------------------
# the oracle parser is pinned by what the reference's tutorial prints for its own files
# ---------------------------------------------------------------------------------------------------
def test_oracle_parser_matches_tutorial_values(answers):
    f = h5_np.File(GSM)
    st = f["/Data/stokes"].read()
    assert st.shape == (4, 10, 768)
    assert int(f["/Header/Ncomponents"].read()) == answers["gsm_icrs"]["Ncomponents"]
    np.testing.assert_allclose(st[0, 0, :5], answers["gsm_icrs"]["stokes_I_first5_K"], rtol=0, atol=5e-9)
    assert f["/Data/stokes"].attrs["unit"] == b"K"
    g = h5_np.File(OLD_POINT)
    st = g["/Data/stokes"].read()
    np.testing.assert_allclose(st[0, 0, :5], answers["old_skyh5_point_sources"]["stokes_I_first5_Jy"], rtol=0, atol=5e-7)
    assert bytes(g["/Header/spectral_type"].read()) == b"subband"
    assert g["/Header/ra"].attrs == {"unit": b"deg", "angtype": b"longitude"}


# ---------------------------------------------------------------------------------------------------
# product reader (C ABI) == oracle parser, value for value
# ---------------------------------------------------------------------------------------------------
def _compare_file(skyh5, path):
    ref = h5_np.tree(path)
    fo = h5_np.File(path)
    with skyh5.H5File(path) as f:
        groups = {"/"} | {k for k, v in ref.items() if v[0] == "group"}
        for g in groups:
            want = sorted(k[len(g.rstrip("/")) + 1:] for k in ref if os.path.dirname(k) == (g.rstrip("/") or "/"))
            assert f.keys(g) == want
        for k, (kind, shape, _dt, attrs) in ref.items():
            assert f.kind(k) == (1 if kind == "group" else 2)
            if kind != "dataset":
                continue
            assert f.shape(k) == shape
            want = fo[k].read()
            got = f.read(k)
            if isinstance(got, str) or (isinstance(got, np.ndarray) and got.dtype.kind == "U"):
                wb = np.char.decode(np.asarray(want, dtype="S"), "utf8") if np.ndim(want) else bytes(want).decode()
                assert np.array_equal(np.asarray(got), np.asarray(wb))
            else:
                assert np.array_equal(np.asarray(got), np.asarray(want)), k
            for name, val in attrs.items():
                assert f.attr(k, name) == bytes(val).decode()
            assert f.attr(k, "no_such_attribute") is None
        assert f.kind("/Header/nope") == 0 and "/Nope/nope" not in f


def test_reader_matches_oracle_on_reference_fixtures(skyh5):
    _compare_file(skyh5, GSM)
    _compare_file(skyh5, OLD_POINT)


def test_reader_known_answers(skyh5, answers):
    with skyh5.H5File(GSM) as f:
        assert f.read("/Header/Ncomponents") == 768 and f.read("/Header/nside") == 8
        assert f.read_strings("/Header/component_type") == "healpix"
        st = f.read("/Data/stokes")
        np.testing.assert_allclose(st[0, 0, :5], answers["gsm_icrs"]["stokes_I_first5_K"], rtol=0, atol=5e-9)
        info = f.info("/Data/stokes")
        assert info.layout == 2 and tuple(info.chunk[:3]) == (1, 5, 384) and info.nchunks == 16
        for nthreads in (1, 3, 16):
            assert np.array_equal(f.read("/Data/stokes", nthreads=nthreads), st)


def test_reader_errors(skyh5, tmp_path):
    p = tmp_path / "not_hdf5.skyh5"
    p.write_bytes(b"x" * 4096)
    with pytest.raises(skyh5.H5Error, match="not an HDF5 file"):
        skyh5.H5File(p)
    with pytest.raises(skyh5.H5Error, match="cannot open"):
        skyh5.H5File(tmp_path / "missing.skyh5")
    raw = bytearray(open(GSM, "rb").read())
    raw[8] = 2  # a libver='latest' superblock
    p2 = tmp_path / "v2.skyh5"
    p2.write_bytes(bytes(raw))
    with pytest.raises(skyh5.H5Error, match="outside the supported"):
        skyh5.H5File(p2)
    p3 = tmp_path / "truncated.skyh5"
    p3.write_bytes(open(GSM, "rb").read()[:20000])
    with skyh5.H5File(p3) as f, pytest.raises(skyh5.H5Error):
        f.read("/Data/stokes")
    with skyh5.H5File(GSM) as f:
        with pytest.raises(skyh5.H5Error, match="no such object"):
            f.read("/Header/nope")
        with pytest.raises(skyh5.H5Error):
            f.read("/Header")


_FUZZ = r"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, {root!r})
from pyradiosky_b200 import skyh5
raw = open({src!r}, 'rb').read()
rng = np.random.default_rng(7)
meta_end = 12000   # object headers, B-trees and heaps of this file live below this offset
bad = 0
for trial in range(300):
    b = bytearray(raw)
    for _ in range(int(rng.integers(1, 6))):
        pos = int(rng.integers(8, meta_end))
        b[pos] = int(rng.integers(0, 256))
    path = {tmp!r}
    with open(path, 'wb') as fh:
        fh.write(b)
    try:
        with skyh5.H5File(path) as f:
            for g in ('/', '/Header', '/Data'):
                for k in f.keys(g):
                    p = g.rstrip('/') + '/' + k
                    if f.kind(p) == 2:
                        f.read(p)
                        f.attr(p, 'unit')
    except (skyh5.H5Error, MemoryError, UnicodeDecodeError, ValueError):
        bad += 1
print('survived', bad)
"""


def test_reader_survives_corrupt_metadata(librsk, tmp_path):
    """Random byte flips in the metadata must give errors (or garbage values), never a crash."""
    code = _FUZZ.format(root=ROOT, src=OLD_POINT, tmp=str(tmp_path / "fuzz.skyh5"))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "survived" in res.stdout


# ---------------------------------------------------------------------------------------------------
# writer: files it produces are read back by both readers
# ---------------------------------------------------------------------------------------------------
def _write_sample(skyh5, path, gzip, rng):
    items = {}
    with skyh5.H5Writer(path) as w:
        w.group("/Header")
        w.group("/Data")
        w.group("/Header/skycoord")
        st = rng.normal(size=(4, 37, 1001))
        w.dataset("/Data/stokes", st, chunks=(1, 5, 384), gzip=gzip, unit="Jy")
        items["/Data/stokes"] = st
        for i in range(23):  # more members than one symbol node holds (8)
            v = rng.integers(0, 100, size=(i,)).astype(np.int64) if i % 2 else rng.normal(size=(3, i + 1)).astype(np.float32)
            items[f"/Header/par{i:02d}"] = v
        items["/Header/name"] = np.array([f"src{i}" for i in range(1001)])
        items["/Header/component_type"] = "point"
        items["/Header/Ncomponents"] = np.int64(1001)
        items["/Header/skycoord/ra"] = rng.uniform(size=1001)
        items["/Header/skycoord/frame"] = "icrs"
        items["/Header/small_ints"] = rng.integers(-5, 5, size=(700, 300)).astype(np.int16)
        items["/Header/unsigned"] = rng.integers(0, 255, size=(17,)).astype(np.uint8)
        items["/Header/rank4"] = rng.normal(size=(3, 4, 5, 6))
        items["/Header/chunked1d"] = rng.normal(size=(1000,))
        items["/Header/empty"] = np.zeros((0,), dtype=np.float64)
        chunk_of = {"/Header/small_ints": (64, 64), "/Header/rank4": (2, 2, 3, 4), "/Header/chunked1d": (128,)}
        for k, v in items.items():
            if k == "/Data/stokes":
                continue
            w.dataset(k, v, chunks=chunk_of.get(k), gzip=gzip)
        w.attr("/Header/skycoord/ra", "unit", "rad")
        w.attr("/Header/skycoord/ra", "object_type", "longitude")
        w.attr("/Header", "note", "group attribute")
    return items


@pytest.mark.parametrize("gzip", [0, 4])
def test_writer_round_trip(skyh5, tmp_path, gzip):
    rng = np.random.default_rng(11 + gzip)
    path = str(tmp_path / f"sample_{gzip}.skyh5")
    items = _write_sample(skyh5, path, gzip, rng)
    fo = h5_np.File(path)
    assert len(fo.b) == fo.eof
    for k, v in items.items():
        got = fo[k].read()
        if isinstance(v, str) or np.asarray(v).dtype.kind == "U":
            want = np.char.encode(np.asarray(v), "utf8")
            assert np.array_equal(np.asarray(got, dtype="S"), want.astype(np.asarray(got).dtype if np.ndim(got) else want.dtype))
        else:
            assert np.array_equal(np.asarray(got), np.asarray(v)), k
            assert np.asarray(got).dtype == np.asarray(v).dtype, k
    assert fo["/Header/skycoord/ra"].attrs == {"unit": b"rad", "object_type": b"longitude"}
    assert fo["/Header"].attrs == {"note": b"group attribute"}
    assert fo["/Data/stokes"].attrs == {"unit": b"Jy"}
    if gzip:
        assert [fid for fid, _ in fo["/Data/stokes"].filters] == [2, 1]
        assert os.path.getsize(path) < 4 * 37 * 1001 * 8 + 700 * 300 * 2
    _compare_file(skyh5, path)
    with skyh5.H5File(path) as f:
        assert f.info("/Data/stokes").nchunks == 4 * 8 * 3  # > 64: a two-level chunk B-tree
        assert f.attr("/Header", "note") == "group attribute"


def test_writer_state_errors(skyh5, tmp_path):
    with skyh5.H5Writer(tmp_path / "e.skyh5") as w:
        w.group("/A")
        with pytest.raises(skyh5.H5Error):
            w.group("/A")  # exists
        with pytest.raises(skyh5.H5Error):
            w.dataset("/B/x", np.zeros(3))  # parent missing
        with pytest.raises(skyh5.H5Error):
            w.attr("/nope", "unit", "Jy")
        with pytest.raises(TypeError):
            w.dataset("/A/c", np.zeros(3, dtype=np.complex128))
        w.dataset("/A/empty", np.zeros((0,), dtype=np.float64))
        w.dataset("/A/scalar", 3.5)
    fo = h5_np.File(str(tmp_path / "e.skyh5"))
    assert fo["/A/scalar"].read() == 3.5 and fo["/A/empty"].read().size == 0


def test_freq_edges_from_centers(skyh5):
    """skymodel.py:305-329"""
    edges = skyh5.get_freq_edges_from_centers(np.array([100e6, 101e6, 102e6]))
    assert np.allclose(edges, [[99.5e6, 100.5e6, 101.5e6], [100.5e6, 101.5e6, 102.5e6]])
    with pytest.raises(ValueError, match="only one frequency center"):
        skyh5.get_freq_edges_from_centers(np.array([1e8]))
    with pytest.raises(ValueError, match="spacing is not constant"):
        skyh5.get_freq_edges_from_centers(np.array([1e8, 1.1e8, 1.3e8]))


# ---------------------------------------------------------------------------------------------------
# GPU: file -> HBM and the SkyModel methods
# ---------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_read_device_matches_host(skyh5, tmp_path):
    import torch

    rng = np.random.default_rng(5)
    for gzip in (0, 3):
        path = str(tmp_path / f"dev_{gzip}.skyh5")
        items = _write_sample(skyh5, path, gzip, rng)
        with skyh5.H5File(path) as f:
            for k in ("/Data/stokes", "/Header/small_ints", "/Header/par20", "/Header/par21", "/Header/skycoord/ra",
                      "/Header/unsigned", "/Header/Ncomponents", "/Header/rank4", "/Header/chunked1d", "/Header/empty"):
                want = np.asarray(items[k])
                dtype = torch.float64 if want.dtype.kind == "f" else torch.int64
                for slab in (0, 4096, 200000):  # default, tiny (ring reuse), medium
                    got = f.read_device(k, "cuda", dtype=dtype, slab_bytes=slab, nthreads=3)
                    torch.cuda.synchronize()
                    assert got.shape == want.shape
                    assert np.array_equal(got.cpu().numpy(), want.astype(np.float64 if dtype == torch.float64 else np.int64)), (k, slab)
    with skyh5.H5File(GSM) as f:
        got = f.read_device("/Data/stokes", "cuda")
        assert np.array_equal(got.cpu().numpy(), h5_np.File(GSM)["/Data/stokes"].read())


@pytest.mark.gpu
def test_from_skyh5_gsm_tutorial_values(answers):
    """docs/tutorial.rst:1288-1297: read the GSM map, print K, convert to Jy / sr, print."""
    import torch

    from pyradiosky_b200 import SkyModel

    with warnings.catch_warnings():
        warnings.simplefilter("error")  # a current-format file must read without warnings
        sky = SkyModel.from_file(GSM)
    a = answers["gsm_icrs"]
    assert sky.Ncomponents == a["Ncomponents"] and sky.Nfreqs == 10
    assert sky.component_type == "healpix" and sky.nside == 8 and sky.hpx_order == "ring" and sky.frame == "icrs"
    assert sky.spectral_type == "full" and sky.stokes_unit == "K"
    assert sky.filename == [os.path.basename(GSM)]
    assert torch.equal(sky.hpx_inds, torch.arange(768, device=sky.device))
    np.testing.assert_allclose(sky.stokes[0, 0, :5].cpu().numpy(), a["stokes_I_first5_K"], rtol=0, atol=5e-9)
    np.testing.assert_allclose(sky.freq_array[:2], [50e6, 61111111.11111111])
    sky.kelvin_to_jansky()
    assert sky.stokes_unit == "Jy / sr"
    np.testing.assert_allclose(sky.stokes[0, 0, :5].cpu().numpy(), a["stokes_I_first5_Jy_per_sr"], rtol=0, atol=5e-9)


@pytest.mark.gpu
def test_from_skyh5_old_point_file(answers):
    """tests/test_skymodel.py:4507-4522 and docs/tutorial.rst:1279-1286."""
    from pyradiosky_b200 import SkyModel

    a = answers["old_skyh5_point_sources"]
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        sky = SkyModel.from_file(OLD_POINT)
    texts = [str(w.message) for w in rec]
    for expected in a["warnings"]:
        assert any(expected in t for t in texts), (expected, texts)
    assert sky.spectral_type == a["spectral_type_after_read"]
    assert sky.component_type == "point" and sky.Ncomponents == 50 and sky.Nfreqs == 20 and sky.frame == "icrs"
    assert sky.name[0] == "J235139-894114" and sky.stokes_unit == "Jy"
    ref = h5_np.File(OLD_POINT)
    np.testing.assert_allclose(sky.lon.cpu().numpy(), np.deg2rad(ref["/Header/ra"].read()), rtol=0, atol=1e-15)
    np.testing.assert_allclose(sky.lat.cpu().numpy(), np.deg2rad(ref["/Header/dec"].read()), rtol=0, atol=1e-15)
    np.testing.assert_allclose(sky.stokes[0, 0, :5].cpu().numpy(), a["stokes_I_first5_Jy"], rtol=0, atol=5e-7)
    sky.jansky_to_kelvin()
    np.testing.assert_allclose(sky.stokes[0, 0, :5].cpu().numpy(), a["stokes_I_first5_K_sr"], rtol=0, atol=5e-9)


def _mock_sky(kind, rng):
    from pyradiosky_b200 import SkyModel

    n = 333
    ra = rng.uniform(0, 2 * np.pi, n)
    dec = np.arcsin(rng.uniform(-1, 1, n))
    names = [f"src{i:04d}" for i in range(n)]
    if kind == "spectral_index":
        st = rng.normal(size=(4, 1, n))
        return SkyModel(name=names, ra=ra, dec=dec, frame="fk5", equinox=1990.5, stokes=st, spectral_type="spectral_index",
                        reference_frequency=np.full(n, 150e6), spectral_index=rng.normal(-0.8, 0.1, n),
                        history="made by a test")
    if kind == "subband":
        st = rng.normal(size=(4, 6, n))
        fa = 100e6 + 8e6 * np.arange(6)
        fe = np.stack([fa - 4e6, fa + 4e6])
        return SkyModel(name=names, ra=ra, dec=dec, frame="galactic", stokes=st, spectral_type="subband", freq_array=fa,
                        freq_edge_array=fe, stokes_error=np.abs(st) * 0.1, extended_model_group=["g%d" % (i % 3) for i in range(n)],
                        extra_columns={"flag": np.arange(n) % 2, "weight": rng.uniform(size=n)})
    if kind == "healpix":
        nside = 4
        st = rng.normal(size=(4, 3, 100))
        return SkyModel(nside=nside, hpx_inds=np.sort(rng.choice(12 * nside**2, 100, replace=False)), hpx_order="nested",
                        hpx_frame="icrs", stokes=st, spectral_type="full", freq_array=[1e8, 1.1e8, 1.2e8], stokes_unit="K",
                        beam_amp=rng.uniform(size=(4, 3, 100)))
    raise ValueError(kind)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["spectral_index", "subband", "healpix"])
@pytest.mark.parametrize("compression", [None, "gzip"])
def test_skyh5_file_loop(kind, compression, tmp_path):
    """tests/test_skymodel.py:3851-3990: write, read back, compare."""
    import torch

    from pyradiosky_b200 import SkyModel

    rng = np.random.default_rng(3)
    sky = _mock_sky(kind, rng)
    path = str(tmp_path / "testfile.skyh5")
    sky.write_skyh5(path, data_compression=compression)
    with pytest.raises(OSError, match="File exists; If overwriting is desired set the clobber"):
        sky.write_skyh5(path)
    sky.write_skyh5(path, clobber=True, data_compression=compression)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        sky2 = SkyModel.from_file(path, filetype="skyh5", run_check=False)
    assert sky2.filename == ["testfile.skyh5"]
    for attr in ("component_type", "spectral_type", "frame", "equinox", "Ncomponents", "Nfreqs", "stokes_unit", "nside",
                 "hpx_order"):
        assert getattr(sky, attr) == getattr(sky2, attr), attr
    assert sky.history in sky2.history and "pyradiosky_b200" in sky2.history
    for attr in ("stokes", "lon", "lat", "hpx_inds", "reference_frequency", "spectral_index", "stokes_error", "beam_amp"):
        a, b = getattr(sky, attr), getattr(sky2, attr)
        assert (a is None) == (b is None), attr
        if a is not None:
            assert torch.equal(a, b), attr
    for attr in ("freq_array", "freq_edge_array", "name", "extended_model_group"):
        a, b = getattr(sky, attr), getattr(sky2, attr)
        assert (a is None) == (b is None), attr
        if a is not None:
            assert np.array_equal(np.asarray(a).astype(str) if np.asarray(a).dtype.kind in "US" else a, b), attr
    if sky.extra_columns:
        assert set(sky2.extra_columns) == set(sky.extra_columns)
        for k in sky.extra_columns:
            assert np.array_equal(sky.extra_columns[k], sky2.extra_columns[k])
    # the file is also readable by the independent parser
    fo = h5_np.File(path)
    assert np.array_equal(fo["/Data/stokes"].read(), sky.stokes.cpu().numpy())
    assert fo["/Data/stokes"].attrs["unit"].decode() == sky.stokes_unit


@pytest.mark.gpu
def test_skyh5_read_errors(tmp_path, skyh5):
    """tests/test_skymodel.py:4099-4155: the reference's ValueError texts."""
    from pyradiosky_b200 import SkyModel

    rng = np.random.default_rng(9)

    def rewrite(path, drop=(), replace=None, drop_header=False):
        """Copy of a written file with members dropped or replaced (h5py's ``del fileobj[...]``)."""
        src = h5_np.File(path)
        out = str(tmp_path / "edited.skyh5")
        with skyh5.H5Writer(out) as w:
            def rec(group, prefix):
                for k in group.keys():
                    p = f"{prefix}/{k}"
                    if p in drop or (drop_header and p == "/Header"):
                        continue
                    o = group[k]
                    if isinstance(o, h5_np.Group):
                        w.group(p)
                        rec(o, p)
                    else:
                        v = o.read() if not (replace and p in replace) else replace[p]
                        if isinstance(v, bytes) or (isinstance(v, np.ndarray) and v.dtype.kind == "S"):
                            v = np.char.decode(np.asarray(v, dtype="S"), "utf8")
                            v = str(v) if v.ndim == 0 else v
                        w.dataset(p, v)
                        for an, av in o.attrs.items():
                            w.attr(p, an, bytes(av).decode())
            rec(src.root, "")
        return out

    sky = _mock_sky("subband", rng)
    path = str(tmp_path / "full.skyh5")
    sky.write_skyh5(path)
    cases = [
        (dict(drop=("/Header/name",)), "Expected parameter name is missing in file"),
        (dict(replace={"/Header/Ncomponents": np.int64(5)}), "Ncomponents is not equal to the size of 'name'."),
        (dict(replace={"/Header/Nfreqs": np.int64(10)}), "Nfreqs is not equal to the size of 'freq_array'."),
        (dict(drop=("/Header/skycoord",)), "No component location information found in file."),
        (dict(drop_header=True), "This is not a proper skyh5 file."),
    ]
    for kw, msg in cases:
        with pytest.raises(ValueError, match=msg):
            SkyModel.from_file(rewrite(path, **kw))
    hp = _mock_sky("healpix", rng)
    hpath = str(tmp_path / "hp.skyh5")
    hp.write_skyh5(hpath)
    for kw, msg in [
        (dict(drop=("/Header/nside",)), "Expected parameter nside is missing in file."),
        (dict(drop=("/Header/hpx_inds",)), "Expected parameter hpx_inds is missing in file."),
        (dict(replace={"/Header/Ncomponents": np.int64(10)}), "Ncomponents is not equal to the size of 'hpx_inds'."),
    ]:
        with pytest.raises(ValueError, match=msg):
            SkyModel.from_file(rewrite(hpath, **kw))
    # no frame information at all: warning + icrs (tests/test_skymodel.py:4402-4420)
    with pytest.warns(UserWarning, match="No frame available in this file, assuming 'icrs'"):
        sky3 = SkyModel.from_file(rewrite(hpath, drop=("/Header/hpx_frame",)))
    assert sky3.frame == "icrs"
    # tests/test_skymodel.py:2910-2918
    with pytest.raises(ValueError, match="Cannot determine the file type."):
        SkyModel.from_file(str(tmp_path / "catalog.foo"))
    with pytest.raises(ValueError, match="Invalid filetype. Filetype options are"):
        SkyModel.from_file(hpath, filetype="foo")


@pytest.mark.gpu
def test_loaded_healpix_model_runs_the_hot_path():
    """File -> HBM -> update_positions -> coherency_calc, against the oracle on the same file contents."""
    import torch

    from oracle import healpix_np, skymodel_np
    from pyradiosky_b200 import EarthLocation, SkyModel, Time

    sky = SkyModel.from_skyh5(GSM)
    loc = EarthLocation.from_dms((-30, 43, 17.5), (21, 25, 41.9), 1073.0)
    sky.update_positions(Time.from_utc("2015-03-01 00:00:00"), loc)
    coh = sky.coherency_calc()
    torch.cuda.synchronize()
    fo = h5_np.File(GSM)
    stokes = fo["/Data/stokes"].read()
    lon, lat = healpix_np.healpix_to_lonlat(8, fo["/Header/hpx_inds"].read(), nested=False)
    # the file also stores the centres (old-style header): the generated ones agree with the stored ones
    np.testing.assert_allclose(lon, fo["/Header/lon"].read(), rtol=0, atol=1e-14)
    np.testing.assert_allclose(lat, fo["/Header/lat"].read(), rtol=0, atol=1e-14)
    alt_az, _, above = skymodel_np.update_positions(lon, lat, sky._astrom)
    assert np.array_equal(sky.above_horizon.cpu().numpy(), above)
    ravg = skymodel_np.calc_average_rotation_matrix(sky._astrom)
    ref = skymodel_np.coherency_calc(stokes, lon, lat, alt_az, above, ravg, polarized=skymodel_np.polarized_indices(stokes))
    got = coh.cpu().numpy()
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(stokes[0]))


# ---------------------------------------------------------------------------------------------------
# The reference's own GLEAM NaN-handling tests (tests/test_skymodel.py:3630-3846), on the same 50 GLEAM
# sources (old_skyh5_point_sources.skyh5 holds the catalogue the reference reads from gleam_50srcs.vot).
# ---------------------------------------------------------------------------------------------------
def _gleam_subband_model():
    from pyradiosky_b200 import SkyModel

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        full = SkyModel.from_file(OLD_POINT)
    return SkyModel(name=full.name, ra=full.lon, dec=full.lat, frame="icrs", stokes=full.stokes.clone(),
                    spectral_type="subband", freq_array=full.freq_array)


def _collect_warnings(fn):
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        out = fn()
    return out, [str(w.message) for w in rec]


@pytest.mark.gpu
@pytest.mark.parametrize("nan_handling", ["propagate", "interp", "clip"])
def test_reference_gleam_nan_handling(nan_handling):
    """tests/test_skymodel.py:3630-3793, line for line."""
    skyobj = _gleam_subband_model()
    interp_freqs = np.asarray([77, 154, 225]) * 1e6
    skyobj_interp = skyobj.at_frequencies(interp_freqs, inplace=False)

    skyobj2 = skyobj.copy()
    nan = float("nan")
    skyobj2.stokes[0, 0:2, 0] = nan  # no low freq support
    skyobj2.stokes[0, 10:11, 1] = nan  # missing freqs in middle
    skyobj2.stokes[0, -2:, 2] = nan  # no high freq support
    skyobj2.stokes[0, :, 3] = nan  # all NaNs
    skyobj2.stokes[0, 1:-2, 4] = nan  # only 2 good freqs
    skyobj2.stokes[0, 0, 5] = nan  # no low or high frequency support
    skyobj2.stokes[0, -1, 5] = nan
    skyobj2.stokes[0, 1:, 6] = nan  # only 1 good freqs

    message = ["Some Stokes values are NaNs."]
    if nan_handling == "propagate":
        message[0] += " All output Stokes values for sources with any NaN values will be NaN."
    else:
        message[0] += " Interpolating using the non-NaN values only."
        message.extend([
            "1 components had all NaN Stokes values. ",
            "3 components had all NaN Stokes values above one or more of the requested frequencies. ",
            "2 components had all NaN Stokes values below one or more of the requested frequencies. ",
            "1 components had too few non-NaN Stokes values for chosen interpolation. Using linear "
            "interpolation for these components instead.",
        ])
        if nan_handling == "interp":
            message[2] += "The Stokes for these components at these frequencies will be NaN."
            message[3] += "The Stokes for these components at these frequencies will be NaN."
        else:
            message[2] += "Using the Stokes value at the highest frequency "
            message[3] += "Using the Stokes value at the lowest frequency "
    message[0] += " You can change the way NaNs are handled using the `nan_handling` keyword."
    skyobj2_interp, texts = _collect_warnings(
        lambda: skyobj2.at_frequencies(interp_freqs, inplace=False, nan_handling=nan_handling))
    for m in message:
        assert any(t.startswith(m) for t in texts), (m, texts)

    got = skyobj2_interp.stokes.cpu().numpy()
    ref = skyobj_interp.stokes.cpu().numpy()
    orig = skyobj.stokes.cpu().numpy()
    if nan_handling == "propagate":
        assert np.all(np.isnan(got[:, :, 0:7]))
        assert np.all(~np.isnan(got[:, :, 7:]))
    elif nan_handling == "interp":
        assert np.all(np.isnan(got[:, 0, 0]))
        assert np.allclose(got[:, 1:, 0], ref[:, 1:, 0], atol=1e-5, rtol=0)
        assert np.all(np.isnan(got[:, 2, 2]))
        assert np.allclose(got[:, 0:-1, 2], ref[:, 0:-1, 2], atol=1e-5, rtol=0)
        assert np.all(np.isnan(got[:, 0, 5]))
        assert np.all(np.isnan(got[:, 2, 5]))
        assert np.allclose(got[:, 1, 2], ref[:, 1, 2], atol=1e-5, rtol=0)
    else:  # clip
        assert np.all(~np.isnan(got[:, :, 0:3]))
        assert np.allclose(got[:, 0, 0], orig[:, 2, 0])
        assert np.allclose(got[:, 1:, 0], ref[:, 1:, 0], atol=1e-5, rtol=0)
        assert np.allclose(got[:, 2, 2], orig[:, -3, 2])
        assert np.allclose(got[:, 0:-1, 2], ref[:, 0:-1, 2], atol=1e-5, rtol=0)
        assert np.allclose(got[:, 0, 5], orig[:, 1, 5])
        assert np.allclose(got[:, 2, 5], orig[:, -2, 5])
        assert np.allclose(got[:, 1, 5], ref[:, 1, 5], atol=1e-5, rtol=0)
    if nan_handling in ["interp", "clip"]:
        assert np.all(np.isnan(got[:, :, 3]))
        assert np.all(~np.isnan(got[:, :, 1]))
        assert np.allclose(got[:, 0, 1], ref[:, 0, 1])
        assert np.allclose(got[:, 2, 1], ref[:, 2, 1])
        assert not np.allclose(got[:, 1, 1], ref[:, 1, 1])
        assert np.allclose(got[:, 1, 1], ref[:, 1, 1], atol=1e-2, rtol=0)
        assert np.all(~np.isnan(got[:, :, 4]))
        assert np.allclose(got[:, 0, 4], ref[:, 0, 4], atol=1e-1, rtol=0)
        assert np.allclose(got[:, 2, 4], ref[:, 2, 4], atol=1e-1, rtol=0)
        assert not np.allclose(got[:, 1, 4], ref[:, 1, 4], atol=1e-1, rtol=0)
        assert np.allclose(got[:, 1, 4], ref[:, 1, 4], atol=2e-1, rtol=0)
    assert np.allclose(got[:, :, 7:], ref[:, :, 7:])


@pytest.mark.gpu
@pytest.mark.parametrize("nan_handling", ["propagate", "interp", "clip"])
def test_reference_gleam_nan_handling_allsrc(nan_handling):
    """tests/test_skymodel.py:3796-3846."""
    skyobj = _gleam_subband_model()
    interp_freqs = np.asarray([77, 154, 225]) * 1e6
    ref = skyobj.at_frequencies(interp_freqs, inplace=False).stokes.cpu().numpy()
    skyobj2 = skyobj.copy()
    skyobj2.stokes[0, 10:11, :] = float("nan")
    out, texts = _collect_warnings(lambda: skyobj2.at_frequencies(interp_freqs, inplace=False, nan_handling=nan_handling))
    head = "Some Stokes values are NaNs." + (
        " All output Stokes values for sources with any NaN values will be NaN." if nan_handling == "propagate"
        else " Interpolating using the non-NaN values only.")
    assert any(t.startswith(head) for t in texts), texts
    got = out.stokes.cpu().numpy()
    if nan_handling == "propagate":
        assert np.all(np.isnan(got))
    else:
        assert np.all(~np.isnan(got))
        assert np.allclose(got[:, 0, :], ref[:, 0, :], atol=1e-5, rtol=0)
        assert np.allclose(got[:, 2, :], ref[:, 2, :], atol=1e-5, rtol=0)
        assert not np.allclose(got[:, 1, :], ref[:, 1, :], atol=1e-5, rtol=0)
        assert np.allclose(got[:, 1, :], ref[:, 1, :], atol=1e-1, rtol=0)


@pytest.mark.gpu
@pytest.mark.parametrize("gzip", [0, 2])
def test_dataset_device_round_trip(skyh5, tmp_path, gzip):
    """HBM -> file (rsk_h5w_dataset_device): 1-, 2- and 3-D float64 / int64 tensors, chunked (several slabs, edge
    chunks), contiguous and empty, read back by the independent parser and by the product reader."""
    import torch

    g = torch.Generator(device="cuda").manual_seed(5)
    cube = torch.randn((3, 41, 1003), dtype=torch.float64, device="cuda", generator=g)
    big = torch.randn((2, 700, 9000), dtype=torch.float64, device="cuda", generator=g)  # 100 MB: more than one 64 MiB slab
    mat = torch.randint(-1000, 1000, (257, 130), dtype=torch.int64, device="cuda", generator=g)
    vec = torch.randn((5000,), dtype=torch.float64, device="cuda", generator=g)
    path = str(tmp_path / "dev.skyh5")
    with skyh5.H5Writer(path) as w:
        w.group("/Data")
        w.dataset_device("/Data/cube", cube, chunks=(1, 8, 256), gzip=gzip, unit="Jy")
        w.dataset_device("/Data/big", big, chunks=(1, 64, 4096), gzip=gzip)
        w.dataset_device("/Data/mat", mat, chunks=(100, 64), gzip=gzip)
        w.dataset_device("/Data/vec", vec, chunks=(1024,), gzip=gzip)
        w.dataset_device("/Data/flat", cube)  # contiguous layout
        w.dataset_device("/Data/bigflat", big)
        w.dataset_device("/Data/empty", torch.empty((0,), dtype=torch.float64, device="cuda"))
        with pytest.raises(TypeError):
            w.dataset_device("/Data/bad", torch.zeros((2, 2, 2, 2), dtype=torch.float64, device="cuda"))
    fo = h5_np.File(path)
    want = {"cube": cube, "big": big, "mat": mat, "vec": vec, "flat": cube, "bigflat": big}
    for k, t in want.items():
        assert np.array_equal(fo[f"/Data/{k}"].read(), t.cpu().numpy()), k
    assert fo["/Data/empty"].read().size == 0 and fo["/Data/cube"].attrs == {"unit": b"Jy"}
    with skyh5.H5File(path) as f:
        for k, t in want.items():
            assert torch.equal(f.read_device(f"/Data/{k}", "cuda", dtype=t.dtype), t), k
