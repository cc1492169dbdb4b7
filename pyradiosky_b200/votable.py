"""VOTable catalogues (TABLEDATA serialisation) without astropy: the front door of the GLEAM-type workloads.

Stands in for ``astropy.io.votable.parse(...)`` + ``table.to_table()`` at the reference call site
src/pyradiosky/skymodel.py:3871-3887 (``SkyModel.read_votable_catalog``).  Only what that reader needs is parsed: the
tables of a file with their ID / name, the FIELDs (name, datatype, unit) and the TABLEDATA rows; BINARY / FITS
serialisations are refused.  Numeric columns come back as float64 (empty cells and the declared null value -> NaN),
``char`` columns as numpy str arrays.  Host-side only; the arrays reach the GPU through the SkyModel constructor.
"""

import xml.etree.ElementTree as ET

import numpy as np

_NUMERIC = {"double", "float", "int", "long", "short", "unsignedByte", "bit", "boolean"}


def _local(tag):
    return tag.rsplit("}", 1)[-1]


class Table:
    def __init__(self, table_id, name, colnames, columns, units):
        self.ID = table_id
        self.name = name
        self.colnames = colnames
        self.columns = columns  # dict name -> numpy array
        self.units = units      # dict name -> unit string or None

    def __len__(self):
        return len(next(iter(self.columns.values()))) if self.columns else 0


def parse(path):
    """All tables of a VOTable file, in file order."""
    root = ET.parse(path).getroot()
    tables = []
    for tab in root.iter():
        if _local(tab.tag) != "TABLE":
            continue
        fields = [f for f in tab if _local(f.tag) == "FIELD"]
        names = [f.get("name") or f.get("ID") for f in fields]
        data = next((d for d in tab if _local(d.tag) == "DATA"), None)
        rows = []
        if data is not None:
            td = next((t for t in data if _local(t.tag) == "TABLEDATA"), None)
            if td is None:
                kinds = [_local(t.tag) for t in data]
                raise ValueError(f"only the TABLEDATA serialisation of VOTable is supported, {path} has {kinds}")
            for tr in td:
                if _local(tr.tag) == "TR":
                    rows.append([(c.text or "") for c in tr if _local(c.tag) == "TD"])
        columns, units = {}, {}
        for k, (f, name) in enumerate(zip(fields, names)):
            cells = [r[k] if k < len(r) else "" for r in rows]
            units[name] = f.get("unit")
            if f.get("datatype") in _NUMERIC and f.get("arraysize") in (None, "1"):
                null = next((v.get("null") for v in f if _local(v.tag) == "VALUES"), None)
                vals = np.empty(len(cells), dtype=np.float64)
                for i, c in enumerate(cells):
                    c = c.strip()
                    vals[i] = np.nan if (c == "" or c == null or c.lower() == "nan") else float(c)
                columns[name] = vals
            else:
                columns[name] = np.array([c.strip() for c in cells], dtype=str)
        tables.append(Table(tab.get("ID"), tab.get("name"), names, columns, units))
    return tables


def get_matching_fields(name_to_match, name_list, exclude_start_pattern=None):
    """The one name of ``name_list`` that contains ``name_to_match`` (case-insensitively); several matches are narrowed
    to the exact one, then to the one not starting with ``exclude_start_pattern`` (skymodel.py:47-81, brittle form)."""
    match_list = [name for name in name_list if name_to_match.casefold() in name.casefold()]
    if len(match_list) > 1:
        exact = [name for name in match_list if name_to_match.casefold() == name.casefold()]
        if len(exact) == 1:
            match_list = exact
        elif exclude_start_pattern is not None:
            kept = [name for name in match_list if not name.startswith(exclude_start_pattern)]
            if len(kept) == 1:
                match_list = kept
        if len(match_list) > 1:
            raise ValueError(f"More than one match for {name_to_match} in {name_list}.")
    elif len(match_list) == 0:
        raise ValueError(f"No match for {name_to_match} in {name_list}.")
    return match_list[0]


# flux units the reader accepts (skymodel.py:3951-3966) and their factor to the unit kept on the object
FLUX_UNITS = {
    "Jy": ("Jy", 1.0), "mJy": ("Jy", 1e-3), "uJy": ("Jy", 1e-6), "kJy": ("Jy", 1e3),
    "Jy/sr": ("Jy / sr", 1.0), "Jy / sr": ("Jy / sr", 1.0), "Jy.sr-1": ("Jy / sr", 1.0),
    "K": ("K", 1.0), "mK": ("K", 1e-3),
    "K sr": ("K sr", 1.0), "K.sr": ("K sr", 1.0),
}
ANGLE_UNITS = {"deg": np.pi / 180.0, "rad": 1.0, "hourangle": np.pi / 12.0, "h": np.pi / 12.0, "arcsec": np.pi / 648000.0,
               "arcmin": np.pi / 10800.0}
