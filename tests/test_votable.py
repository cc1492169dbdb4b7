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
