"""Tab / space separated text catalogues: SkyModel.read_text_catalog (skymodel.py:4204-4407) and write_text_catalog
(:5031-5182), host side.  Coordinate columns name their frame as the reference's ``_get_frame_comp_cols`` (:116-203)
expects -- ``ra_j2000`` / ``dec_j2000`` (FK5, J2000), ``ra_icrs`` / ``dec_icrs``, ``l_galactic`` / ``b_galactic``,
``ra_fk5_j2015.5`` ... -- for the frames this build carries (icrs, fk5(equinox), galactic); FK4 columns are recognised
and refused."""

import re

import numpy as np

_LONLAT = {"icrs": ("ra", "dec"), "fk5": ("ra", "dec"), "fk4": ("ra", "dec"), "galactic": ("l", "b")}


def frame_desc_str(frame, equinox):
    """skymodel.py:91-113."""
    if frame in ("icrs", "galactic"):
        return frame
    if frame == "fk5":
        return "j2000" if float(equinox) == 2000.0 else f"fk5_J{float(equinox):.3f}"
    raise ValueError(f"{frame} is not supported for writing text files.")


def frame_comp_cols(colnames):
    """(frame, equinox, lon column, lat column) from the header's column names (skymodel.py:116-203)."""
    ra_fk5, dec_fk5 = re.compile("ra_j2000"), re.compile("de[c]?_j2000")
    ra_fk4, dec_fk4 = re.compile("ra_b1950"), re.compile("de[c]?_b1950")
    frame, short = None, False
    for name in colnames:
        low = name.casefold()
        if ra_fk5.match(low) or dec_fk5.match(low):
            frame, short = "fk5", True
            break
        if ra_fk4.match(low) or dec_fk4.match(low):
            frame, short = "fk4", True
            break
        frame = next((f for f in ("galactic", "icrs", "fk5", "fk4") if f in low), None)
        if frame is not None:
            break
    if frame is None:
        raise ValueError("frame not recognized from coordinate column")
    lon_col = lat_col = None
    equinox = 2000.0
    if short:
        ra_pat, dec_pat = (ra_fk5, dec_fk5) if frame == "fk5" else (ra_fk4, dec_fk4)
        for name in colnames:
            if lon_col is None and ra_pat.match(name.casefold()):
                lon_col = name
            if lat_col is None and dec_pat.match(name.casefold()):
                lat_col = name
    else:
        lon_name, lat_name = _LONLAT[frame]
        for name in colnames:
            low = name.casefold()
            if low.startswith(lon_name + "_" + frame):
                lon_col = name
            if low.startswith(lat_name + "_" + frame):
                lat_col = name
        if lon_col is None or lat_col is None:
            raise ValueError("Longitudinal and Latidudinal component columns not identified.")
        lon_str = lon_col.casefold().split(lon_name + "_", 1)[1]
        lat_str = lat_col.casefold().split(lat_name + "_", 1)[1]
        if lon_str != lat_str:
            raise ValueError("Longitudinal and Latitudinal columns have different frame indicators. "
                             f"Columns are: [{lon_col}, {lat_col}].")
        if frame == "fk5" and "_j" in lon_str:
            equinox = float(lon_str.split("_j", 1)[1])
    if lon_col is None:
        raise ValueError("Longitudinal component column not identified.")
    if lat_col is None:
        raise ValueError("Latitudinal component column not identified.")
    return frame, equinox, lon_col, lat_col


def name2freq(colname):
    """Frequency in Hz from a column name such as ``Flux_150_MHz`` (pyuvdata CSTBeam.name2freq)."""
    fi = colname.rfind("Hz")
    number = float(re.findall(r"\d*\.\d+|\d+", colname[:fi])[-1])
    return number * {"k": 1e3, "M": 1e6, "G": 1e9}.get(colname[fi - 1], 1.0)


def read(catalog_csv):
    """Everything SkyModel.__init__ needs from a text catalogue (skymodel.py:4254-4405), angles in radians."""
    with open(catalog_csv) as cfile:
        header = cfile.readline()
    header = [h.strip() for h in header.split() if h[0] != "["]  # ignore units in header
    frame, equinox, lon_col, lat_col = frame_comp_cols(header)
    flux_fields = [c for c in header if c.casefold().startswith("flux")]
    flux_error_fields = [c for c in flux_fields if "error" in c.casefold()]
    for c in flux_error_fields:
        flux_fields.remove(c)
    flux_lower = [c.casefold() for c in flux_fields]
    if flux_error_fields and len(flux_error_fields) != len(flux_fields):
        raise ValueError("Number of flux error fields does not match number of flux fields.")
    err_lower = [c.casefold() for c in flux_error_fields]
    header_lower = [c.casefold() for c in header]
    expected = ["source_id", lon_col.casefold(), lat_col.casefold()]
    freq_array = None
    if "frequency" in header_lower:
        if len(flux_fields) != 1:
            raise ValueError("If frequency column is present, only one flux column allowed.")
        expected.append(flux_lower[0])
        if flux_error_fields:
            expected.append("flux_error")
        expected.append("frequency")
        spectral_type = "flat"
        if "spectral_index" in header_lower:
            spectral_type = "spectral_index"
            expected.append("spectral_index")
        n_freqs = 1
    else:
        frequencies = []
        for fluxname in flux_fields:
            if "Hz" in fluxname:
                frequencies.append(name2freq(fluxname))
            elif len(flux_fields) > 1:
                raise ValueError("Multiple flux fields, but they do not all contain a frequency.")
        if frequencies:
            n_freqs = len(frequencies)
            spectral_type = "full" if n_freqs > 1 else "flat"
            if flux_error_fields:
                for k in range(n_freqs):
                    expected += [flux_lower[k], err_lower[k]]
            else:
                expected += flux_lower
            freq_array = np.array(frequencies)
        else:
            n_freqs, spectral_type = 1, "flat"
            expected.append("flux")
            if flux_error_fields:
                expected.append("flux_error")
    if expected != header_lower:
        raise ValueError(f"Header does not match expectations. Expected columnsare: {expected}, header columns were: "
                         f"{header_lower}")
    table = np.atleast_1d(np.genfromtxt(catalog_csv, autostrip=True, skip_header=1, dtype=None, encoding="utf-8"))
    cols = table.dtype.names
    n = len(table)
    lon_ind, lat_ind = header.index(lon_col), header.index(lat_col)
    stokes = np.zeros((4, n_freqs, n))
    stokes_error = np.zeros((4, n_freqs, n)) if flux_error_fields else None
    for k in range(n_freqs):
        if flux_error_fields:
            stokes[0, k] = table[cols[k * 2 + 3]]
            stokes_error[0, k] = table[cols[k * 2 + 4]]
        else:
            stokes[0, k] = table[cols[k + 3]]
    reference_frequency = spectral_index = None
    if "frequency" in header_lower and freq_array is None:
        reference_frequency = np.asarray(table[cols[header_lower.index("frequency")]], dtype=np.float64)
        if "spectral_index" in header_lower:
            spectral_index = np.asarray(table[cols[header_lower.index("spectral_index")]], dtype=np.float64)
    return dict(name=table[cols[0]].astype(str), lon=np.deg2rad(np.mod(table[cols[lon_ind]].astype(np.float64), 360.0)),
                lat=np.deg2rad(table[cols[lat_ind]].astype(np.float64)), frame=frame, equinox=equinox, stokes=stokes,
                stokes_unit="Jy", spectral_type=spectral_type, freq_array=freq_array,
                reference_frequency=reference_frequency, spectral_index=spectral_index, stokes_error=stokes_error)


def write(filename, sky):
    """skymodel.py:5031-5182 for a point-source SkyModel whose Stokes are in Jy (checked by the caller)."""
    lon_name, lat_name = _LONLAT[sky.frame]
    desc = frame_desc_str(sky.frame, sky.equinox)
    header = f"source_id\t{lon_name}_{desc} [deg]\t{lat_name}_{desc} [deg]"
    fmt = "{}\t{:0.8f}\t{:0.8f}"
    have_err = sky.stokes_error is not None
    ref = None if sky.reference_frequency is None else sky.reference_frequency.cpu().numpy()
    spi = None if sky.spectral_index is None else sky.spectral_index.cpu().numpy()
    if ref is not None:
        header += "\tFlux [Jy]"
        if have_err:
            header += "\tFlux_error [Jy]"
            fmt += "\t{:0.8f}"
        header += "\tFrequency [Hz]"
        fmt += "\t{:0.8f}\t{:0.8f}"
        if spi is not None:
            header += "\tSpectral_Index"
            fmt += "\t{:0.8f}"
    elif sky.freq_array is not None:
        for hz in sky.freq_array:
            if hz > 1e9:
                freq_str = f"{hz * 1e-9:g}_GHz"
            elif hz > 1e6:
                freq_str = f"{hz * 1e-6:g}_MHz"
            elif hz > 1e3:
                freq_str = f"{hz * 1e-3:g}_kHz"
            else:
                freq_str = f"{hz:g}_Hz"
            fmt += "\t{:0.8f}"
            header += f"\tFlux_{freq_str} [Jy]"
            if have_err:
                header += f"\tFlux_error_{freq_str} [Jy]"
                fmt += "\t{:0.8f}"
    else:
        header += "\tFlux [Jy]"
        fmt += "\t{:0.8f}"
        if have_err:
            header += "\tFlux_error [Jy]"
            fmt += "\t{:0.8f}"
    lon = np.rad2deg(sky.lon.cpu().numpy())
    lat = np.rad2deg(sky.lat.cpu().numpy())
    flux = sky.stokes[0].cpu().numpy().T          # (N, Nfreqs)
    err = sky.stokes_error[0].cpu().numpy().T if have_err else None
    with open(filename, "w+") as fo:
        fo.write(header + "\n")
        for i in range(sky.Ncomponents):
            if have_err:
                fluxes = [v for pair in zip(flux[i], err[i]) for v in pair]
            else:
                fluxes = list(flux[i])
            tail = []
            if ref is not None:
                tail.append(ref[i])
                if spi is not None:
                    tail.append(spi[i])
            fo.write(fmt.format(sky.name[i], lon[i], lat[i], *fluxes, *tail) + "\n")
