"""Minimal native FITS I/O for HEALPix maps (survey ingestion, reference
surveys.py:1013-1019,1116-1119: ``hp.read_map``).

Only what a HEALPix map file needs: a primary HDU followed by a BINTABLE
extension whose columns hold the map, either one pixel per row (TFORM 'E'/'D')
or the usual 1024 pixels per row ('1024E'), big-endian, optionally gzip
compressed.  ``read_healpix_map`` returns what ``healpy.read_map(path)`` returns
with its defaults: field 0 as float64 in RING ordering (a NESTED file is
reordered), UNSEEN kept as is.  Partial-sky (INDXSCHM = EXPLICIT) files are
expanded to full sky with UNSEEN.  healpy is not required.
"""
from __future__ import annotations

import gzip

import numpy as np

from . import healpix as hp

_BLOCK = 2880
_TFORM = {"L": ("i1", 1), "B": ("u1", 1), "I": (">i2", 2), "J": (">i4", 4), "K": (">i8", 8),
          "E": (">f4", 4), "D": (">f8", 8)}


def _open(path):
    with open(path, "rb") as f:
        magic = f.read(2)
    return gzip.open(path, "rb") if magic == b"\x1f\x8b" else open(path, "rb")


def _parse_value(raw):
    raw = raw.split("/")[0].strip() if not raw.strip().startswith("'") else raw.strip()
    if raw.startswith("'"):
        end = raw.find("'", 1)
        return raw[1:end].strip()
    if raw in ("T", "F"):
        return raw == "T"
    try:
        return int(raw)
    except ValueError:
        try:
            return float(raw.replace("D", "E"))
        except ValueError:
            return raw


def _read_header(f):
    cards = {}
    while True:
        block = f.read(_BLOCK)
        if len(block) < _BLOCK:
            raise ValueError("truncated FITS header")
        for i in range(0, _BLOCK, 80):
            card = block[i:i + 80].decode("ascii", "replace")
            key = card[:8].strip()
            if key == "END":
                return cards
            if card[8:10] == "= ":
                cards[key] = _parse_value(card[10:])


def _skip_data(f, header):
    naxis = header.get("NAXIS", 0)
    if naxis == 0:
        return
    n = abs(header.get("BITPIX", 8)) // 8 * header.get("GCOUNT", 1)
    size = header.get("PCOUNT", 0)
    prod = 1
    for k in range(1, naxis + 1):
        prod *= header[f"NAXIS{k}"]
    size = n * (prod + size)
    f.read(-(-size // _BLOCK) * _BLOCK)


def read_healpix_map(path, field=0, nest=False):
    """Dense float64 HEALPix map from a FITS file (RING unless ``nest=True``)."""
    with _open(path) as f:
        primary = _read_header(f)
        if not primary.get("SIMPLE", False):
            raise ValueError(f"{path}: not a FITS file")
        _skip_data(f, primary)
        hdr = _read_header(f)
        if hdr.get("XTENSION") != "BINTABLE":
            raise ValueError(f"{path}: first extension is not a BINTABLE")
        nrows, rowbytes, nfields = hdr["NAXIS2"], hdr["NAXIS1"], hdr["TFIELDS"]
        fields = []
        for k in range(1, nfields + 1):
            tform = str(hdr[f"TFORM{k}"]).strip()
            code = tform[-1]
            rep = int(tform[:-1]) if tform[:-1] else 1
            if code not in _TFORM:
                raise ValueError(f"{path}: unsupported TFORM '{tform}'")
            fields.append((f"c{k}", _TFORM[code][0], (rep,)))
        dt = np.dtype(fields)
        if dt.itemsize != rowbytes:
            raise ValueError(f"{path}: row size {rowbytes} does not match TFORMs")
        table = np.frombuffer(f.read(nrows * rowbytes), dtype=dt, count=nrows)
    explicit = str(hdr.get("INDXSCHM", "IMPLICIT")).upper() == "EXPLICIT"
    col = field + 1 + (1 if explicit else 0)
    values = np.asarray(table[f"c{col}"], dtype=np.float64).ravel()
    nside = hdr.get("NSIDE") or hp.npix2nside(values.size)
    if explicit:
        index = np.asarray(table["c1"], dtype=np.int64).ravel()
        full = np.full(hp.nside2npix(nside), hp.UNSEEN)
        full[index] = values
        values = full
    values = values[:hp.nside2npix(nside)]
    file_nest = str(hdr.get("ORDERING", "RING")).upper().startswith("NEST")
    if file_nest and not nest:
        values = hp.reorder_nest_to_ring(values)
    elif not file_nest and nest:
        values = hp.reorder_ring_to_nest(values)
    return values


def _card(key, value, comment=""):
    if isinstance(value, bool):
        v = f"{'T' if value else 'F':>20}"
    elif isinstance(value, (int, np.integer)):
        v = f"{int(value):>20}"
    elif isinstance(value, float):
        v = f"{value:>20.12E}"
    else:
        v = f"'{str(value):<8}'"
        v = f"{v:<20}"
    return f"{key:<8}= {v} / {comment}"[:80].ljust(80)


def write_healpix_map(path, m, nest=False, dtype=np.float32, per_row=1024):
    """Write a full-sky map as a standard HEALPix BINTABLE (for tests and for
    exporting synthetic surveys)."""
    m = np.asarray(m)
    nside = hp.get_nside(m)
    per_row = min(per_row, m.size)
    code = "E" if np.dtype(dtype) == np.float32 else "D"
    be = np.ascontiguousarray(m, dtype=">f4" if code == "E" else ">f8")
    rowbytes = per_row * be.itemsize
    cards = [_card("SIMPLE", True), _card("BITPIX", 8), _card("NAXIS", 0), _card("EXTEND", True)]
    head = "".join(cards) + "END".ljust(80)
    head = head.ljust(-(-len(head) // _BLOCK) * _BLOCK)
    ext = [("XTENSION", "BINTABLE"), ("BITPIX", 8), ("NAXIS", 2), ("NAXIS1", rowbytes),
           ("NAXIS2", m.size // per_row), ("PCOUNT", 0), ("GCOUNT", 1), ("TFIELDS", 1),
           ("TTYPE1", "T"), ("TFORM1", f"{per_row}{code}"), ("PIXTYPE", "HEALPIX"),
           ("ORDERING", "NESTED" if nest else "RING"), ("NSIDE", nside), ("INDXSCHM", "IMPLICIT"),
           ("FIRSTPIX", 0), ("LASTPIX", m.size - 1)]
    ehead = "".join(_card(k, v) for k, v in ext) + "END".ljust(80)
    ehead = ehead.ljust(-(-len(ehead) // _BLOCK) * _BLOCK)
    data = be.tobytes()
    data += b"\0" * (-len(data) % _BLOCK)
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "wb") as f:
        f.write(head.encode("ascii") + ehead.encode("ascii") + data)
