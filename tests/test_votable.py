"""VOTable catalogues: the stdlib parser (CPU) and SkyModel.read_votable_catalog / read_gleam_catalog / from_file (GPU).

Mirrors tests/test_skymodel.py:2879-3077 of the reference on its own data files (tests/golden/ref_fixture_*.vot).  The
parser is pinned by two independent known answers: the tutorial's print of the first five 76 MHz fluxes
(docs/tutorial.rst:1279-1283) and the reference's old_skyh5_point_sources.skyh5, which an older pyradiosky wrote from
this very catalogue through astropy's VOTable reader.
"""

import os
import re

import numpy as np
import pytest

from oracle import h5_np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
GLEAM_VOT = os.path.join(GOLDEN, "ref_fixture_gleam_50srcs.vot")
SIMPLE_VOT = os.path.join(GOLDEN, "ref_fixture_simple_test.vot")
OLD_VOT = os.path.join(GOLDEN, "ref_fixture_single_source_old.vot")
GLEAM_H5 = os.path.join(GOLDEN, "ref_fixture_gleam50_old_header.skyh5")
BANDS = ["076", "084", "092", "099", "107", "115", "122", "130", "143", "151", "158", "166", "174", "181", "189", "197",
         "204", "212", "220", "227"]


def test_votable_parser_known_answers(known_answers):
    from pyradiosky_b200 import votable as vot

    (table,) = vot.parse(GLEAM_VOT)
    assert table.ID == "VIII_100_gleamegc" and table.name == "VIII/100/gleamegc" and len(table) == 50
    assert table.units["RAJ2000"] == "deg" and table.units["Fint076"] == "Jy" and table.units["Fpwide"] == "Jy/beam"
    ka = known_answers["tutorial_unit_conversion"]
    assert np.array_equal(table.columns["Fint076"][:5], ka["gleam_jy"])
    g = h5_np.File(GLEAM_H5)
    st = np.array(g["/Data/stokes"].read(), dtype=np.float64)
    cube = np.array([table.columns["Fint" + b] for b in BANDS])
    assert np.array_equal(cube, st[0], equal_nan=True)
    assert np.array_equal(table.columns["RAJ2000"], np.array(g["/Header/ra"].read()))
    assert np.array_equal(table.columns["DEJ2000"], np.array(g["/Header/dec"].read()))
    assert [n.decode() for n in g["/Header/name"].read()] == table.columns["GLEAM"].tolist()
    assert table.columns["GLEAM"][0] == "J235139-894114"
    (simple,) = vot.parse(SIMPLE_VOT)
    assert simple.colnames == ["source_id", "RAJ2000", "DEJ2000", "Si"] and simple.columns["source_id"].tolist() == ["TEST", "TEST2"]
    assert simple.columns["DEJ2000"].tolist() == [-30.0, 0.0]
    assert any(t.ID is None for t in vot.parse(OLD_VOT))


def test_get_matching_fields():
    """tests/test_skymodel.py:2788-2830."""
    from pyradiosky_b200.votable import get_matching_fields

    names = ["_RAJ2000", "_DEJ2000", "RAJ2000", "DEJ2000", "GLEAM"]
    assert get_matching_fields("gle", names) == "GLEAM"
    assert get_matching_fields("raj2000", names, exclude_start_pattern="_") == "RAJ2000"
    assert get_matching_fields("RAJ2000", names) == "RAJ2000"
    with pytest.raises(ValueError, match="More than one match for aj2000"):
        get_matching_fields("aj2000", names)
    with pytest.raises(ValueError, match="No match for foo"):
        get_matching_fields("foo", names)


@pytest.fixture(scope="module")
def rsk(librsk):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pyradiosky_b200

    return pyradiosky_b200


@pytest.mark.gpu
@pytest.mark.parametrize("spec_type", ["flat", "subband", "spectral_index"])
def test_read_gleam(rsk, spec_type):
    """tests/test_skymodel.py:2879-2907 (+ the spectral_index flavour)."""
    sky = rsk.SkyModel.from_file(GLEAM_VOT, spectral_type=spec_type, with_error=True, run_check=False)
    assert sky.Ncomponents == 50 and sky.frame == "fk5" and sky.stokes_unit == "Jy"
    assert sky.stokes_error is not None and sky.stokes_error.shape == sky.stokes.shape
    assert sky.filename == ["ref_fixture_gleam_50srcs.vot"] and sky.name[0] == "J235139-894114"
    g = h5_np.File(GLEAM_H5)
    assert np.allclose(sky.lon.cpu().numpy(), np.deg2rad(np.array(g["/Header/ra"].read())), rtol=0, atol=1e-15)
    assert np.allclose(sky.lat.cpu().numpy(), np.deg2rad(np.array(g["/Header/dec"].read())), rtol=0, atol=1e-15)
    if spec_type == "subband":
        assert sky.Nfreqs == 20 and sky.spectral_type == "subband"
        assert np.array_equal(sky.stokes.cpu().numpy(), np.array(g["/Data/stokes"].read(), dtype=np.float64), equal_nan=True)
        assert np.array_equal(sky.freq_array, np.array(g["/Header/freq_array"].read(), dtype=np.float64))
        assert sky.freq_edge_array.shape == (2, 20) and sky.freq_edge_array[0, 0] == 72e6 and sky.freq_edge_array[1, -1] == 231e6
    else:
        assert sky.Nfreqs == 1 and float(sky.reference_frequency[0]) == 200e6
        if spec_type == "spectral_index":
            assert sky.spectral_index is not None and sky.spectral_index.shape == (50,)
    sky.select(non_negative=True)
    assert sky.Ncomponents == (32 if spec_type == "subband" else 50)
    if spec_type == "subband":
        sky2 = rsk.SkyModel.from_file(GLEAM_VOT, spectral_type=spec_type, with_error=True, use_paper_freqs=True)
        sky2.select(non_negative=True)
        assert sky2 != sky
        assert np.max(np.abs(sky2.freq_array - sky.freq_array)) <= 0.6e6
        assert np.max(np.abs(sky2.freq_edge_array - sky.freq_edge_array)) <= 1.08e6
        # the catalogue goes straight into the hot path
        loc = rsk.EarthLocation.from_dms((-30, 43, 17.5), (21, 25, 41.9), 1073.0)
        sky.update_positions(rsk.Time.from_utc("2015-03-01 00:00:00"), loc)
        out = sky.coherency_calc(freqs=np.linspace(80e6, 220e6, 5))
        assert out.shape == (2, 2, 5, int(sky.above_horizon.sum()))


@pytest.mark.gpu
def test_read_votable_and_errors(rsk, tmp_path):
    """tests/test_skymodel.py:2910-3077."""
    with pytest.raises(ValueError, match="Invalid filetype. Filetype options are"):
        rsk.SkyModel.from_file(GLEAM_VOT, filetype="foo")
    with pytest.raises(ValueError, match="Cannot determine the file type."):
        rsk.SkyModel.from_file(str(tmp_path / "catalog.foo"))
    with pytest.raises(ValueError, match="spectral_type full is not an allowed type"):
        rsk.SkyModel.from_gleam_catalog(GLEAM_VOT, spectral_type="full")
    sky = rsk.SkyModel.from_votable_catalog(SIMPLE_VOT, "VIII_1000_single", "source_id", "RAJ2000", "DEJ2000", "Si", frame="fk5")
    assert sky.Ncomponents == 2 and sky.spectral_type == "flat" and sky.name.tolist() == ["TEST", "TEST2"]
    assert sky.stokes[0, 0].tolist() == [1.0, 0.0] and abs(float(sky.lat[0]) + np.pi / 6) < 1e-15
    sky2 = rsk.SkyModel.from_votable_catalog(SIMPLE_VOT, "VIII/1000/single", "source_id", "RAJ2000", "DEJ2000", "Si",
                                             frame="fk5", run_check=False)
    assert sky2 == sky
    sky3 = rsk.SkyModel.from_file(SIMPLE_VOT, table_name="VIII_1000_single", id_column="source_id", lon_column="RAJ2000",
                                  lat_column="DEJ2000", flux_columns="Si", frame="fk5")
    assert sky3 == sky
    with pytest.raises(ValueError, match=re.escape(f"File {OLD_VOT} contains tables with no name or ID.")):
        rsk.SkyModel.from_votable_catalog(OLD_VOT, "GLEAM", "GLEAM", "RAJ2000", "DEJ2000", "Fintwide", frame="fk5")
    flux_columns = ["Fint" + b for b in BANDS]
    flux_error_columns = ["e_Fint" + b for b in BANDS]
    freq_array = np.array([76, 84, 92, 99, 107, 115, 122, 130, 143, 151, 158, 166, 174, 181, 189, 197, 204, 212, 220, 227]) * 1e6
    lower = np.array([72, 80, 88, 95, 103, 111, 118, 126, 139, 147, 154, 162, 170, 177, 185, 193, 200, 208, 216, 223]) * 1e6
    upper = np.array([80, 88, 95, 103, 111, 118, 126, 134, 147, 154, 162, 170, 177, 185, 193, 200, 208, 216, 223, 231]) * 1e6
    base = {"filename": GLEAM_VOT, "filetype": "vot", "table_name": "GLEAM", "id_column": "GLEAM", "lon_column": "RAJ2000",
            "lat_column": "DEJ2000", "frame": "fk5", "flux_columns": flux_columns, "flux_error_columns": flux_error_columns,
            "freq_array": freq_array, "freq_edge_array": np.stack([lower, upper])}
    good = rsk.SkyModel.from_file(**base)
    assert good == rsk.SkyModel.from_gleam_catalog(GLEAM_VOT, with_error=True)
    cases = [
        ({"reference_frequency": 200e6}, ["freq_array", "freq_edge_array"],
         "Frequency information must be provided with multiple flux columns."),
        ({"freq_array": np.array([150e6, 200e6]), "flux_columns": ["Fintwide", "Fpwide"]}, ["freq_edge_array"],
         "All flux columns must have compatible units"),
        ({}, ["freq_edge_array"], "freq_edge_array must be provided for multiple flux columns if freq_array is not regularly spaced."),
        ({"flux_columns": ["Fint076", "Fint084"], "flux_error_columns": ["e_Fp076", "e_Fint084"],
          "freq_edge_array": np.array([[72, 80], [80, 88]])}, ["freq_array"],
         "All flux error columns must have units compatible with"),
        ({}, ["table_name"], "table_name is required when reading vot files."),
        ({}, ["id_column"], "id_column is required when reading vot files."),
        ({}, ["lon_column"], "lon_column is required when reading vot files."),
        ({}, ["lat_column"], "lat_column is required when reading vot files."),
        ({}, ["frame"], "frame is required when reading vot files."),
        ({}, ["flux_columns"], "flux_columns is required when reading vot files."),
    ]
    for update, drop, msg in cases:
        kw = {k: v for k, v in base.items() if k not in drop}
        kw.update(update)
        with pytest.raises(ValueError, match=msg):
            rsk.SkyModel.from_file(**kw)
    # regularly spaced centres: the edges are derived, with the reference's warning
    with pytest.warns(UserWarning, match="freq_edge_array not set, calculating it from the freq_array."):
        reg = rsk.SkyModel.from_votable_catalog(GLEAM_VOT, "GLEAM", "GLEAM", "RAJ2000", "DEJ2000", ["Fint076", "Fint084", "Fint092"],
                                                frame="fk5", freq_array=np.array([76e6, 84e6, 92e6]))
    assert np.array_equal(reg.freq_edge_array, np.array([[72e6, 80e6, 88e6], [80e6, 88e6, 96e6]]))
    with pytest.warns(UserWarning, match="freq_array not set, calculating it from the freq_edge_array."):
        reg2 = rsk.SkyModel.from_votable_catalog(GLEAM_VOT, "GLEAM", "GLEAM", "RAJ2000", "DEJ2000", ["Fint076", "Fint084"],
                                                 frame="fk5", freq_edge_array=np.array([[72e6, 80e6], [80e6, 88e6]]))
    assert np.array_equal(reg2.freq_array, [76e6, 84e6])


# ---------------------------------------------------------------------------------------------------
# text catalogues (skymodel.py:4204-4413, :5031-5182; tests/test_skymodel.py:3214-3470)
# ---------------------------------------------------------------------------------------------------
POINT_TXT = os.path.join(GOLDEN, "ref_fixture_pointsource_catalog.txt")
MOCK_TXT = os.path.join(GOLDEN, "ref_fixture_mock_hera_text.txt")


def test_text_reader_on_reference_files():
    """tests/test_skymodel.py:3214-3238 and :3436-3441 (host side)."""
    from pyradiosky_b200 import textcat

    kw = textcat.read(POINT_TXT)
    assert kw["frame"] == "fk5" and kw["equinox"] == 2000.0 and kw["spectral_type"] == "flat"
    assert kw["name"].tolist() == ["src1", "src2", "src3"]
    assert np.allclose(np.rad2deg(kw["lon"]), [1.0, 1.3, 1.35]) and np.allclose(np.rad2deg(kw["lat"]), [-30, -31, -32])
    assert kw["stokes"][0, 0].tolist() == [0.5, 2.0, 5.0] and kw["reference_frequency"].tolist() == [1e6] * 3
    mock = textcat.read(MOCK_TXT)
    assert mock["name"].tolist() == ["src" + str(k) for k in range(mock["stokes"].shape[2])]
    assert mock["reference_frequency"][0] == 150e6 and mock["stokes"][0, 0, 0] == 1.0


def test_text_frame_columns_and_frequencies():
    """skymodel.py:116-203 for the frames carried here; tests/test_skymodel.py:2832-2875."""
    from pyradiosky_b200.textcat import frame_comp_cols, frame_desc_str, name2freq

    assert frame_comp_cols(["source_id", "RA_J2000", "Dec_J2000", "Flux"]) == ("fk5", 2000.0, "RA_J2000", "Dec_J2000")
    assert frame_comp_cols(["source_id", "ra_j2000", "de_j2000"])[2:] == ("ra_j2000", "de_j2000")
    assert frame_comp_cols(["source_id", "ra_icrs", "dec_icrs"]) == ("icrs", 2000.0, "ra_icrs", "dec_icrs")
    assert frame_comp_cols(["source_id", "l_galactic", "b_galactic"]) == ("galactic", 2000.0, "l_galactic", "b_galactic")
    assert frame_comp_cols(["source_id", "ra_fk5_j2002.000", "dec_fk5_j2002.000"])[:2] == ("fk5", 2002.0)
    assert frame_comp_cols(["source_id", "ra_b1950", "dec_b1950"])[0] == "fk4"
    with pytest.raises(ValueError, match="frame not recognized from coordinate column"):
        frame_comp_cols(["source_id", "ra_foo", "dec_foo"])
    with pytest.raises(ValueError, match="Longitudinal and Latidudinal component columns not identified."):
        frame_comp_cols(["source_id", "foo_icrs", "dec_icrs"])
    with pytest.raises(ValueError, match="Latitudinal component column not identified."):
        frame_comp_cols(["source_id", "ra_j2000", "foo_j2000x"])
    assert frame_desc_str("fk5", 2000.0) == "j2000" and frame_desc_str("fk5", 2002.0) == "fk5_J2002.000"
    assert frame_desc_str("icrs", 2000.0) == "icrs" and frame_desc_str("galactic", 2000.0) == "galactic"
    with pytest.raises(ValueError, match="fk4 is not supported for writing text files."):
        frame_desc_str("fk4", 1950.0)
    assert name2freq("Flux_150_MHz") == 150e6 and name2freq("Flux_1.5_GHz") == 1.5e9
    assert name2freq("Flux_200_kHz") == 200e3 and name2freq("Flux_76_Hz") == 76.0


@pytest.mark.gpu
@pytest.mark.parametrize("with_error", [False, True])
@pytest.mark.parametrize("spec_type", ["flat", "spectral_index", "full", "subband"])
def test_text_catalog_loop(rsk, tmp_path, spec_type, with_error):
    """tests/test_skymodel.py:3285-3337: write_text_catalog -> from_file gives the object back."""
    sky = rsk.SkyModel.from_file(GLEAM_VOT, spectral_type="subband" if spec_type == "full" else spec_type,
                                 with_error=with_error)
    if spec_type in ("full", "subband"):
        sky.select(non_nan="any")  # ('nan' cells round-trip through text, but NaN != NaN in the numpy reader's eyes)
    if spec_type == "spectral_index":
        sky.select(component_inds=np.nonzero(~np.isnan(sky.spectral_index.cpu().numpy()))[0])
    if spec_type == "full":
        sky.at_frequencies(sky.freq_array)
        sky.freq_edge_array = None
    sky.calculate_rise_set_lsts(-0.5361)
    sky.frame_coherency = None
    fname = str(tmp_path / "temp_cat.txt")
    if spec_type == "subband":
        with pytest.raises(ValueError, match="Text files do not support subband types, use write_skyh5."):
            sky.write_text_catalog(fname)
        return
    sky.write_text_catalog(fname)
    sky2 = rsk.SkyModel.from_file(fname)
    assert sky2.spectral_type == spec_type and sky2.filename == ["temp_cat.txt"] and sky2.frame == "fk5"
    sky2.filename, sky2.history = sky.filename, sky.history
    assert sky2 == sky  # values written with 8 decimals: equal at the UVParameter tolerances
    if spec_type == "flat":
        # again with no reference_frequency field
        ref = sky.reference_frequency
        sky.reference_frequency = None
        sky.write_text_catalog(fname)
        sky2 = rsk.SkyModel.from_file(fname, run_check=False)
        sky2.filename, sky2.history = sky.filename, sky.history
        assert sky2 == sky and sky2.reference_frequency is None
        # again with flat & freq_array, in other units too (:3340-3351)
        for mult in (1.0, 1e-6, 1e-3, 1e3):
            sky.freq_array = np.atleast_1d(np.unique(ref.cpu().numpy())) * mult
            sky.write_text_catalog(fname)
            sky2 = rsk.SkyModel.from_file(fname, filetype="text")
            sky2.filename, sky2.history = sky.filename, sky.history
            assert sky2 == sky, mult


@pytest.mark.gpu
def test_text_catalog_files_and_errors(rsk, tmp_path):
    """tests/test_skymodel.py:3214-3238, :3354-3470."""
    sky = rsk.SkyModel.from_file(POINT_TXT)
    assert sky.filename == ["ref_fixture_pointsource_catalog.txt"] and sky.Ncomponents == 3 and sky.frame == "fk5"
    assert np.allclose(np.rad2deg(sky.lon.cpu().numpy()), [1.0, 1.3, 1.35]) and sky.stokes[0, 0].tolist() == [0.5, 2.0, 5.0]
    mock = rsk.SkyModel.from_file(MOCK_TXT)
    assert mock.name.tolist() == ["src" + str(k) for k in range(mock.Ncomponents)]
    fname = str(tmp_path / "temp_cat.txt")
    hpx = rsk.SkyModel(nside=2, stokes=np.ones((4, 1, 48)), spectral_type="flat", hpx_frame="icrs")
    with pytest.raises(ValueError, match="component_type must be 'point' to use this method."):
        hpx.write_text_catalog(fname)
    g = rsk.SkyModel.from_file(GLEAM_VOT, spectral_type="flat")
    g.jansky_to_kelvin()
    with pytest.raises(ValueError, match="Stokes units must be equivalent to Jy to use this method."):
        g.write_text_catalog(fname)
    g = rsk.SkyModel.from_gleam_catalog(GLEAM_VOT, spectral_type="flat")
    g.extended_model_group = np.asarray(["foo"] * g.Ncomponents)
    with pytest.raises(ValueError, match="Text files do not support catalogs with extended_model_group"):
        g.write_text_catalog(fname)
    # a frame the text format can name, a frame it cannot find, lost coordinate columns (:3393-3433)
    for frame in ("icrs", "galactic"):
        z = rsk.SkyModel(name=["zen_source"], ra=[1.1], dec=[-0.5], frame=frame, stokes=np.array([1.0, 0, 0, 0]).reshape(4, 1, 1),
                         spectral_type="flat")
        z.write_text_catalog(fname)
        back = rsk.SkyModel.from_text_catalog(fname)
        assert back.frame == frame and abs(float(back.lon[0]) - 1.1) < 1e-9 and abs(float(back.lat[0]) + 0.5) < 1e-9
    z = rsk.SkyModel(name=["zen_source"], ra=[1.1], dec=[-0.5], frame="icrs", stokes=np.array([1.0, 0, 0, 0]).reshape(4, 1, 1),
                     spectral_type="flat")
    for old, new, msg in (("icrs", "foo", "frame not recognized from coordinate column"),
                          ("ra_", "foo_", "Longitudinal and Latidudinal component columns not identified."),
                          ("dec_", "foo_", "Longitudinal and Latidudinal component columns not identified.")):
        z.write_text_catalog(fname)
        text = open(fname).read().split("\n")
        text[0] = text[0].replace(old, new)
        open(fname, "w").write("\n".join(text))
        with pytest.raises(ValueError, match=msg):
            rsk.SkyModel.from_text_catalog(fname)
    # flux / error column bookkeeping (:3452-3470)
    e = rsk.SkyModel.from_file(GLEAM_VOT, with_error=True, run_check=False)
    e.select(non_negative=True, non_nan="any")
    e.at_frequencies(e.freq_array)
    e.freq_edge_array = None
    e.write_text_catalog(fname)
    text = open(fname).read().replace("Flux_76_MHz [Jy]", "Frequency [Hz]")
    open(fname, "w").write(text)
    with pytest.raises(ValueError, match="Number of flux error fields does not match number of flux fields."):
        rsk.SkyModel.from_file(fname)


# ---------------------------------------------------------------------------------------------------
# FHD catalogues (skymodel.py:4415-4627; tests/test_skymodel.py:3079-3211)
# ---------------------------------------------------------------------------------------------------
FHD = {k: os.path.join(GOLDEN, f"ref_fixture_{k}.sav") for k in ("extended_source_test", "fhd_catalog_no_extend",
                                                                  "fhd_catalog_bad", "fhd_source_array")}


def test_fhd_reader_host():
    """Labels of expanded extended sources (tests/test_skymodel.py:3183-3195), component counts (:3104-3133), beam
    amplitudes (:3136-3180), extra columns (:3088-3095), errors (:3198-3211)."""
    import scipy.io

    from pyradiosky_b200 import fhdcat

    with pytest.warns(UserWarning, match="Source IDs are not unique. Defining unique IDs."):
        kw = fhdcat.read(FHD["extended_source_test"], expand_extended=True)
    assert kw["extended_model_group"].tolist() == ["0", "0", "0", "1", "1"]
    assert kw["name"].tolist() == ["0_1", "0_2", "0_3", "1_1", "1_2"]
    assert kw["spectral_type"] == "spectral_index" and kw["frame"] == "icrs" and np.all(kw["reference_frequency"] > 50e6)
    for key, var in (("fhd_catalog_no_extend", "catalog"), ("fhd_source_array", "source_array")):
        catalog = scipy.io.readsav(FHD[key])[var]
        ext = [i for i in range(len(catalog)) if catalog["extend"][i] is not None]
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            flat = fhdcat.read(FHD[key], expand_extended=False, extra_columns={"x": "image_x", "y": "image_y"})
            full = fhdcat.read(FHD[key], expand_extended=True, extra_columns={"x": "image_x", "y": "image_y"})
        assert flat["stokes"].shape == (4, 1, len(catalog)) and list(flat["extra_columns"]) == ["image_x", "image_y"]
        assert flat["extended_model_group"] is None and flat["filename"] == [f"ref_fixture_{key}.sav"]
        assert full["stokes"].shape[2] == len(catalog) - len(ext) + sum(len(catalog[e]["extend"]) for e in ext)
        assert (full["extended_model_group"] is None) == (len(ext) == 0)
        assert full["extra_columns"]["image_x"].shape == (full["stokes"].shape[2],)
        assert np.array_equal(flat["stokes"][0, 0], [catalog[i]["flux"]["I"][0] for i in range(len(catalog))])
        if "BEAM" in catalog.dtype.names:
            beam = np.array([[catalog[i]["beam"]["XX"][0], catalog[i]["beam"]["YY"][0], abs(catalog[i]["beam"]["XY"][0]),
                              abs(catalog[i]["beam"]["YX"][0])] for i in range(len(catalog))]).T
            assert np.array_equal(flat["beam_amp"][:, 0, :], beam)
            # components of an extended source sit where the source was, each with its own beam values
            first = ext[0]
            ncomp = len(catalog[first]["extend"])
            assert full["name"][first:first + ncomp].tolist() == [f"{flat['name'][first]}_{k}" for k in range(1, ncomp + 1)]
            assert np.array_equal(full["beam_amp"][:, 0, :first], beam[:, :first])
        else:
            assert flat["beam_amp"] is None
    with pytest.raises(KeyError, match="does not contain a known catalog name. "):
        fhdcat.read(FHD["fhd_catalog_bad"])
    with pytest.warns(UserWarning), pytest.raises(KeyError, match="foo in extra_columns not available in catalog. Available extra_columns are:"):
        fhdcat.read(FHD["extended_source_test"], extra_columns={"foo": "foo"})


@pytest.mark.gpu
@pytest.mark.filterwarnings("ignore:Source IDs are not unique")
def test_fhd_catalog_on_gpu(rsk):
    sky = rsk.SkyModel.from_file(FHD["fhd_source_array"], expand_extended=True, extra_columns={"x": "image_x"})
    assert sky.spectral_type == "spectral_index" and sky.Ncomponents == 1055 and sky.beam_amp.shape == (4, 1, 1055)
    assert sky.extended_model_group is not None and sky.extra_columns["image_x"].shape == (1055,)
    assert sky.filename == ["ref_fixture_fhd_source_array.sav"]
    flat = rsk.SkyModel.from_fhd_catalog(FHD["fhd_catalog_no_extend"], expand_extended=False)
    assert flat.Ncomponents == 100 and flat.extended_model_group is None and flat.extra_columns is None
    assert bool((flat.reference_frequency > 50e6).all())
    # straight into the hot path
    loc = rsk.EarthLocation.from_dms((-30, 43, 17.5), (21, 25, 41.9), 1073.0)
    sky.update_positions(rsk.Time.from_utc("2015-03-01 00:00:00"), loc)
    out = sky.coherency_calc(freqs=np.array([150e6, 180e6]))
    assert out.shape == (2, 2, 2, int(sky.above_horizon.sum()))
    # groups survive a select; a select that keeps no extended component drops the attribute (skymodel.py:3260-3265)
    ext = np.nonzero(sky.extended_model_group != "")[0]
    plain = np.nonzero(sky.extended_model_group == "")[0]
    assert sky.select(component_inds=ext[:7], inplace=False).extended_model_group.tolist() == sky.extended_model_group[ext[:7]].tolist()
    assert sky.select(component_inds=plain[:7], inplace=False).extended_model_group is None


def test_votable_fast_scan_equals_xml_parser(tmp_path):
    """The bytes.split scan of large tables (votable._parse_fast) against the ElementTree stream on the reference's
    files and on a synthetic table with empty cells, self-closed cells, entities, a short row and a column filter; files
    the scan does not cover fall back to the XML parser."""
    from pyradiosky_b200 import votable as vot

    def same(a, b):
        assert a is not None and len(a) == len(b)
        for ta, tb in zip(a, b):
            assert (ta.ID, ta.name, ta.colnames, ta.units) == (tb.ID, tb.name, tb.colnames, tb.units)
            assert list(ta.columns) == list(tb.columns)
            for k in ta.columns:
                x, y = ta.columns[k], tb.columns[k]
                assert x.dtype == y.dtype, k
                assert np.array_equal(x, y, equal_nan=True) if x.dtype.kind == "f" else np.array_equal(x, y), k

    for path in (GLEAM_VOT, SIMPLE_VOT, OLD_VOT):
        same(vot._parse_fast(path, None), vot._parse_stream(path, None))
    head = ('<?xml version="1.0"?>\n<VOTABLE version="1.4" xmlns="http://www.ivoa.net/xml/VOTable/v1.4"><RESOURCE>'
            '<TABLE ID="t1" name="first"><FIELD name="id" datatype="char" arraysize="*"/>'
            '<FIELD name="x" datatype="double" unit="deg"><VALUES null="-99"/></FIELD>'
            '<FIELD name="flux" datatype="float" unit="mJy"/><FIELD name="n" datatype="int"/><DATA><TABLEDATA>\n')
    rows = ('<TR><TD>A &amp; B</TD><TD> 1.5 </TD><TD>2e3</TD><TD>7</TD></TR>\n'
            '<TR><TD>c&lt;d</TD><TD>-99</TD><TD/><TD></TD></TR>\n'
            '<TR><TD/><TD>nan</TD><TD>4</TD></TR>\n')
    tail = '</TABLEDATA></DATA></TABLE><TABLE ID="t2" name="empty"><FIELD name="q" datatype="double"/></TABLE></RESOURCE></VOTABLE>\n'
    p = tmp_path / "syn.vot"
    p.write_text(head + rows * 3 + tail)
    fast, slow = vot._parse_fast(str(p), None), vot._parse_stream(str(p), None)
    same(fast, slow)
    t1, t2 = vot.parse(str(p))
    assert len(t1) == 9 and len(t2) == 0 and t2.colnames == ["q"]
    assert t1.columns["id"][:3].tolist() == ["A & B", "c<d", ""]
    assert t1.columns["x"][0] == 1.5 and np.isnan(t1.columns["x"][1]) and np.isnan(t1.columns["x"][2])
    assert t1.columns["flux"][:3].tolist()[0] == 2000.0 and np.isnan(t1.columns["flux"][1]) and t1.columns["flux"][2] == 4.0
    assert t1.columns["n"][0] == 7.0 and np.isnan(t1.columns["n"][1]) and np.isnan(t1.columns["n"][2])
    (only,) = vot.parse(str(p), columns=lambda name: name in ("id", "flux"))[:1]
    assert list(only.columns) == ["id", "flux"] and only.colnames == ["id", "x", "flux", "n"]
    with pytest.raises(KeyError, match="was not kept"):
        only.column("x")
    # not covered by the scan: comments inside the data, namespace prefixes -> the XML parser reads them
    p2 = tmp_path / "comment.vot"
    p2.write_text(head + "<!-- a comment -->" + rows + tail)
    assert vot._parse_fast(str(p2), None) is None and len(vot.parse(str(p2))[0]) == 3
    p3 = tmp_path / "binary.vot"
    p3.write_text(head.replace("<TABLEDATA>", "<BINARY><STREAM>AAAA</STREAM></BINARY>").rstrip("\n")
                  + "</DATA></TABLE></RESOURCE></VOTABLE>")
    with pytest.raises(ValueError, match="only the TABLEDATA serialisation"):
        vot.parse(str(p3))
