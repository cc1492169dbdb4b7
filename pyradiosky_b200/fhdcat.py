"""FHD (Fast Holographic Deconvolution) IDL save-file catalogues: SkyModel.read_fhd_catalog (skymodel.py:4415-4621),
host side, through scipy.io.readsav as upstream.  The reference edits its arrays in place for every extended source
(np.delete / np.insert); here the component list is built once, in catalogue order, with the same result: a source
with an ``EXTEND`` table is replaced by its components (``<id>_<k>``, k from 1, grouped under the parent's id)."""

import os
import warnings

import numpy as np


def read(filename_sav, expand_extended=True, extra_columns=None):
    import scipy.io

    sav = scipy.io.readsav(filename_sav)
    if "catalog" in sav:
        catalog = sav["catalog"]
    elif "source_array" in sav:
        catalog = sav["source_array"]
    else:
        raise KeyError(f"File {filename_sav} does not contain a known catalog name. "
                       f"File variables include {list(sav.keys())}")
    ids = np.array(catalog["id"]).copy()
    nsrcs = len(catalog)
    use_beam = "BEAM" in catalog.dtype.names
    unique_ids, cts = np.unique(ids, return_counts=True)
    if cts.max() > 1:
        warnings.warn("Source IDs are not unique. Defining unique IDs.")
        for this_id in unique_ids[np.nonzero(cts > 1)]:
            # leave the first instance of this id alone; new ids above the current maximum keep all of them unique
            where = np.nonzero(ids == this_id)[0][1:]
            ids[where] = np.arange(where.size) + (ids.max() + 1)
    ids = ids.astype(str)
    known = {"ID", "RA", "DEC", "FREQ", "ALPHA", "FLUX", "BEAM", "EXTEND"}
    extra_available = [c for c in catalog.dtype.names if c not in known]
    extra_map = {}
    if extra_columns is not None:
        for key, value in extra_columns.items():
            if key.upper() not in extra_available:
                raise KeyError(f"{key} in extra_columns not available in catalog. "
                               f"Available extra_columns are: {extra_available}")
            extra_map[key.upper()] = value

    def stokes_of(rec):
        return [rec["flux"]["I"][0], rec["flux"]["Q"][0], rec["flux"]["U"][0], rec["flux"]["V"][0]]

    def beam_of(rec):
        return [rec["beam"]["XX"][0], rec["beam"]["YY"][0], np.abs(rec["beam"]["XY"][0]), np.abs(rec["beam"]["YX"][0])]

    names, groups, ra, dec, freq, alpha, stokes, beam = [], [], [], [], [], [], [], []
    extras = {v: [] for v in extra_map.values()}
    any_extended = False
    for i in range(nsrcs):
        src = catalog[i]
        ext = src["extend"] if "EXTEND" in catalog.dtype.names else None
        if expand_extended and ext is not None:
            any_extended = True
            for k in range(len(ext)):
                comp = ext[k]
                names.append(f"{ids[i]}_{k + 1}")
                groups.append(ids[i])
                ra.append(comp["ra"]); dec.append(comp["dec"]); freq.append(comp["freq"]); alpha.append(comp["alpha"])
                stokes.append(stokes_of(comp))
                if use_beam:
                    beam.append(beam_of(comp))
                for key, out in extra_map.items():
                    extras[out].append(comp[key])
        else:
            names.append(ids[i])
            groups.append("")
            ra.append(src["ra"]); dec.append(src["dec"]); freq.append(src["freq"]); alpha.append(src["alpha"])
            stokes.append(stokes_of(src))
            if use_beam:
                beam.append(beam_of(src))
            for key, out in extra_map.items():
                extras[out].append(src[key])
    n = len(names)
    st = np.asarray(stokes, dtype=np.float64).T.reshape(4, 1, n)
    return dict(
        name=np.array(names, dtype=str), lon=np.deg2rad(np.mod(np.asarray(ra, dtype=np.float64), 360.0)),
        lat=np.deg2rad(np.asarray(dec, dtype=np.float64)), frame="icrs", stokes=st, stokes_unit="Jy",
        spectral_type="spectral_index", reference_frequency=np.asarray(freq, dtype=np.float64) * 1e6,  # MHz in the file
        spectral_index=np.asarray(alpha, dtype=np.float64),
        beam_amp=np.asarray(beam, dtype=np.float64).T.reshape(4, 1, n) if use_beam else None,
        extended_model_group=np.array(groups, dtype=str) if (expand_extended and any_extended) else None,
        extra_columns={k: np.asarray(v) for k, v in extras.items()} or None,
        filename=[os.path.basename(os.fspath(filename_sav))],
    )
