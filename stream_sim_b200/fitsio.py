"""Minimal native FITS I/O for HEALPix maps (survey ingestion, reference
surveys.py:1013-1019,1116-1119: ``hp.read_map``).

Only what a HEALPix map file needs: a primary HDU followed by a BINTABLE
extension whose columns hold the map, either one pixel per row (TFORM 'E'/'D')
or the usual 1024 pixels per row ('1024E'), big-endian, optionally gzip
compressed.  ``read_healpix_map`` returns what ``healpy.read_map(path)`` returns
with its defaults: field 0 as float64 in RING ordering (a NESTED file is
reordered), UNSEEN kept as is.  Partial-sky (INDXSCHM = EXPLICIT) files are
expanded to full sky with UNSEEN.  healpy is not required.

``read_healsparse`` reads a HealSparse ``.hsp`` map (reference surveys.py:1255-1284:
``healsparse.HealSparseMap.read`` + ``generate_healpix_map``) without the healsparse
package: the file is a FITS container with a coverage-index image (``COV``) and the
sparse pixel image (``SPARSE``), the latter usually tile-compressed (GZIP_1 / GZIP_2,
lossless).  Layout per the published healsparse format (v1.x); no ``.hsp`` file ships
with the reference, so this reader is pinned only by round trips through
``write_healsparse`` (parity unpinned).  Record-array and wide-mask maps are not maps a
survey depth / extinction entry can be and raise ``ValueError``.
"""
from __future__ import annotations

import gzip
import zlib

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


# ---------------------------------------------------------------- HealSparse (.hsp)
_BITPIX = {8: "u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}


def _data_bytes(header):
    naxis = header.get("NAXIS", 0)
    if naxis == 0:
        return 0
    prod = 1
    for k in range(1, naxis + 1):
        prod *= header[f"NAXIS{k}"]
    return abs(header.get("BITPIX", 8)) // 8 * header.get("GCOUNT", 1) * (prod + header.get("PCOUNT", 0))


def _read_hdus(f):
    """[(header, raw data bytes)] for every HDU of an open FITS stream."""
    hdus = []
    while True:
        try:
            hdr = _read_header(f)
        except ValueError:
            break
        size = _data_bytes(hdr)
        data = f.read(-(-size // _BLOCK) * _BLOCK)[:size]
        if len(data) < size:
            raise ValueError("truncated FITS data unit")
        hdus.append((hdr, data))
    return hdus


def _scaled(arr, hdr, prefix=""):
    """Apply BZERO/BSCALE of integer images (FITS unsigned-integer convention)."""
    zero, scale = hdr.get(prefix + "BZERO", hdr.get("BZERO", 0)), hdr.get("BSCALE", 1)
    if arr.dtype.kind in "iu" and (zero != 0 or scale != 1):
        if scale == 1 and float(zero) == int(zero):
            return arr.astype(np.int64) + int(zero)
        return arr * float(scale) + float(zero)
    return arr


def _image_1d(hdr, data):
    """1-D image of a plain IMAGE/primary HDU or of a tile-compressed one (ZIMAGE)."""
    if hdr.get("ZIMAGE", False):
        if hdr.get("ZNAXIS", 1) != 1:
            raise ValueError("HealSparse images are one-dimensional")
        n, tile = hdr["ZNAXIS1"], hdr.get("ZTILE1", hdr["ZNAXIS1"])
        dt = np.dtype(_BITPIX[hdr["ZBITPIX"]])
        cmp_type = str(hdr.get("ZCMPTYPE", "")).strip().upper()
        if cmp_type not in ("GZIP_1", "GZIP_2", "NOCOMPRESS"):
            raise ValueError(f"tile compression '{cmp_type}' is not supported (GZIP_1/GZIP_2 are)")
        # Lossless float compression (ZQUANTIZ = 'NONE', what healsparse writes) keeps the
        # tile bytes in COMPRESSED_DATA.  Quantised files keep scaled integers there and fall
        # back to GZIP_COMPRESSED_DATA for tiles that could not be quantised: only the
        # former need ZSCALE / ZZERO, and those are refused tile by tile below.
        quantised = (dt.kind == "f" and str(hdr.get("ZQUANTIZ", "NONE")).strip().upper() != "NONE"
                     and any(str(hdr.get(f"TTYPE{k}", "")).strip().upper() == "ZSCALE"
                             for k in range(1, hdr["TFIELDS"] + 1)))
        nrows, rowbytes = hdr["NAXIS2"], hdr["NAXIS1"]
        heap0 = hdr.get("THEAP", nrows * rowbytes)
        # column layout: variable-length descriptors 'P' (2 x int32) or 'Q' (2 x int64)
        cols, off = {}, 0
        for k in range(1, hdr["TFIELDS"] + 1):
            tform = str(hdr[f"TFORM{k}"]).strip()
            name = str(hdr.get(f"TTYPE{k}", "")).strip().upper()
            body = tform.lstrip("0123456789")
            if body[0] in "PQ":
                width = 8 if body[0] == "P" else 16
                cols[name] = (off, body[0], body[1])
            else:
                rep = int(tform[:len(tform) - len(body)] or 1)
                width = rep * _TFORM[body[0]][1]
            off += width
        out = np.empty(n, dtype=dt.newbyteorder("="))
        for r in range(nrows):
            lo, hi = r * tile, min((r + 1) * tile, n)
            raw, shuffled = None, cmp_type == "GZIP_2"
            for name in ("COMPRESSED_DATA", "GZIP_COMPRESSED_DATA", "UNCOMPRESSED_DATA"):
                if name not in cols:
                    continue
                coff, kind, _ = cols[name]
                desc = np.frombuffer(data, dtype=">i4" if kind == "P" else ">i8", count=2,
                                     offset=r * rowbytes + coff)
                nb, ho = int(desc[0]), int(desc[1])
                if nb == 0:
                    continue
                if quantised and name == "COMPRESSED_DATA":
                    raise ValueError("quantised (lossy) float tile compression is not supported")
                chunk = data[heap0 + ho: heap0 + ho + nb * (1 if name != "UNCOMPRESSED_DATA" else dt.itemsize)]
                if name == "UNCOMPRESSED_DATA" or cmp_type == "NOCOMPRESS":
                    raw, shuffled = chunk, False
                else:
                    raw = zlib.decompress(chunk, 47)          # gzip or zlib framing
                    shuffled = shuffled and name == "COMPRESSED_DATA"
                break
            if raw is None:
                raise ValueError(f"tile {r} of the compressed image holds no data")
            cnt = hi - lo
            if shuffled:                                      # GZIP_2: byte planes, MSB first
                planes = np.frombuffer(raw, dtype=np.uint8, count=cnt * dt.itemsize)
                raw = np.ascontiguousarray(planes.reshape(dt.itemsize, cnt).T).tobytes()
            out[lo:hi] = np.frombuffer(raw, dtype=dt, count=cnt)
        return _scaled(out, hdr, "Z")
    if hdr.get("NAXIS", 0) != 1:
        raise ValueError("HealSparse images are one-dimensional")
    dt = np.dtype(_BITPIX[hdr["BITPIX"]])
    return _scaled(np.frombuffer(data, dtype=dt, count=hdr["NAXIS1"]).astype(dt.newbyteorder("=")), hdr)


def read_healsparse(path):
    """(nside_coverage, nside_sparse, cov_index_map, sparse_map, sentinel) of a ``.hsp`` file."""
    with _open(path) as f:
        hdus = [(h, d) for h, d in _read_hdus(f) if h.get("NAXIS", 0) > 0]
    by_name = {str(h.get("EXTNAME", "")).strip().upper(): (h, d) for h, d in hdus}
    cov = by_name.get("COV", hdus[0] if hdus else None)
    sparse = by_name.get("SPARSE", hdus[1] if len(hdus) > 1 else None)
    if cov is None or sparse is None:
        raise ValueError(f"{path}: not a HealSparse file (needs COV and SPARSE HDUs)")
    chdr, shdr = cov[0], sparse[0]
    if str(shdr.get("PIXTYPE", chdr.get("PIXTYPE", ""))).strip().upper() != "HEALSPARSE":
        raise ValueError(f"{path}: PIXTYPE is not HEALSPARSE")
    if shdr.get("XTENSION") == "BINTABLE" and not shdr.get("ZIMAGE", False):
        raise ValueError(f"{path}: record-array HealSparse maps are not supported")
    if shdr.get("WIDEMASK", False):
        raise ValueError(f"{path}: wide-mask HealSparse maps are not supported")
    cov_index = np.asarray(_image_1d(chdr, cov[1]), dtype=np.int64)
    values = _image_1d(shdr, sparse[1])
    nside_cov = int(chdr.get("NSIDE") or hp.npix2nside(cov_index.size))
    nside_sparse = int(shdr["NSIDE"])
    sentinel = shdr.get("SENTINEL", hp.UNSEEN)
    return nside_cov, nside_sparse, cov_index, values, sentinel


def healsparse_to_ring(nside_cov, nside_sparse, cov_index, values, sentinel, nest=False):
    """``HealSparseMap.generate_healpix_map(nside=nside_sparse, nest=nest)``: dense float64
    map with UNSEEN where the sparse map holds its sentinel (or is not covered)."""
    shift = 2 * int(round(np.log2(nside_sparse / nside_cov)))
    nfine = 1 << shift
    npix = hp.nside2npix(nside_sparse)
    out = np.full(npix, hp.UNSEEN)
    covered = np.nonzero(cov_index + np.arange(cov_index.size, dtype=np.int64) * nfine >= nfine)[0]
    for c in covered:                                   # one coverage pixel = nfine NEST pixels
        start = int(cov_index[c]) + int(c) * nfine
        block = np.asarray(values[start:start + nfine])
        if isinstance(sentinel, float) and np.isnan(sentinel):
            good = ~np.isnan(block)
        else:
            good = block != np.asarray(sentinel).astype(block.dtype)
        dst = out[c * nfine:(c + 1) * nfine]
        dst[good] = block[good].astype(np.float64)
    return out if nest else hp.reorder_nest_to_ring(out)


def read_healsparse_map(path, nest=False):
    """Dense float64 RING (or NEST) map of a ``.hsp`` file at its own nside_sparse."""
    return healsparse_to_ring(*read_healsparse(path), nest=nest)


def write_healsparse(path, ring_map, nside_coverage=32, compress="GZIP_2", dtype=np.float32):
    """Write a dense RING map (UNSEEN/NaN = not observed) as a HealSparse ``.hsp`` file:
    coverage pixels holding no valid pixel are dropped.  ``compress``: None, "GZIP_1" or
    "GZIP_2" (lossless tile compression, one tile per coverage pixel as healsparse does)."""
    m = np.asarray(ring_map, dtype=np.float64)
    nside = hp.get_nside(m)
    nest = hp.reorder_ring_to_nest(np.where(np.isnan(m), hp.UNSEEN, m))
    shift = 2 * int(round(np.log2(nside / nside_coverage)))
    nfine = 1 << shift
    ncov = hp.nside2npix(nside_coverage)
    blocks = nest.reshape(ncov, nfine)
    has = (blocks != hp.UNSEEN).any(axis=1)
    cov_index = -np.arange(ncov, dtype=np.int64) * nfine              # not covered: overflow block
    order = np.nonzero(has)[0]
    cov_index[order] += (1 + np.arange(order.size, dtype=np.int64)) * nfine
    sparse = np.concatenate([np.full(nfine, hp.UNSEEN), blocks[order].ravel()]).astype(dtype)
    code = {np.dtype(np.float32): -32, np.dtype(np.float64): -64}[np.dtype(dtype)]
    be = sparse.astype(np.dtype(dtype).newbyteorder(">"))

    def unit(cards, payload):
        head = "".join(_card(k, v) for k, v in cards) + "END".ljust(80)
        head = head.ljust(-(-len(head) // _BLOCK) * _BLOCK)
        return head.encode("ascii") + payload + b"\0" * (-len(payload) % _BLOCK)

    cov_unit = unit([("SIMPLE", True), ("BITPIX", 64), ("NAXIS", 1), ("NAXIS1", ncov),
                     ("EXTEND", True), ("EXTNAME", "COV"), ("PIXTYPE", "HEALSPARSE"),
                     ("NSIDE", nside_coverage)], cov_index.astype(">i8").tobytes())
    common = [("EXTNAME", "SPARSE"), ("PIXTYPE", "HEALSPARSE"), ("NSIDE", nside),
              ("SENTINEL", float(hp.UNSEEN))]
    if compress is None:
        sp_unit = unit([("XTENSION", "IMAGE"), ("BITPIX", code), ("NAXIS", 1),
                        ("NAXIS1", sparse.size), ("PCOUNT", 0), ("GCOUNT", 1)] + common, be.tobytes())
    else:
        item = be.dtype.itemsize
        heap, desc = [], []
        for r in range(sparse.size // nfine):
            tile = be[r * nfine:(r + 1) * nfine]
            raw = tile.tobytes()
            if compress == "GZIP_2":
                raw = np.ascontiguousarray(np.frombuffer(raw, np.uint8).reshape(nfine, item).T).tobytes()
            z = gzip.compress(raw, 4)
            desc.append((len(z), sum(len(h) for h in heap)))
            heap.append(z)
        table = np.array(desc, dtype=">i4").tobytes()
        blob = b"".join(heap)
        sp_unit = unit([("XTENSION", "BINTABLE"), ("BITPIX", 8), ("NAXIS", 2), ("NAXIS1", 8),
                        ("NAXIS2", len(desc)), ("PCOUNT", len(blob)), ("GCOUNT", 1), ("TFIELDS", 1),
                        ("TTYPE1", "COMPRESSED_DATA"), ("TFORM1", "1PB(%d)" % max(len(h) for h in heap)),
                        ("ZIMAGE", True), ("ZBITPIX", code), ("ZNAXIS", 1), ("ZNAXIS1", sparse.size),
                        ("ZTILE1", nfine), ("ZCMPTYPE", compress)] + common, table + blob)
    with open(path, "wb") as f:
        f.write(cov_unit + sp_unit)
