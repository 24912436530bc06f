"""Columnar result of an injection, kept off pandas until someone asks for pandas.

The reference returns a pandas DataFrame whose ``mag_<band>_obs`` columns are *object*
columns (floats mixed with the string ``"BAD_MAG"``, reference observed.py:900-908) and
writes CSV (observed.py:961-982, bin/generate_stream.py:69).  Once the per-star arithmetic
runs at 10^10 stars/s, building N Python objects and formatting text dominate everything,
so ``StreamInjector.inject(..., output="table")`` returns a :class:`StarTable` instead:

* columns stay where they were produced (device tensors) until first use; each is copied to
  the host once, through pinned memory;
* a bad magnitude is a NaN plus one shared validity mask -- Arrow nulls / Parquet nulls, and
  ``"BAD_MAG"`` strings only in ``to_pandas(compat=True)`` / ``to_csv(compat=True)``, the
  reference's own format;
* sinks: ``to_arrow()``, ``to_parquet(path)``, ``to_csv(path)``, ``to_pandas()``.
"""
from __future__ import annotations

import numpy as np

BAD_MAG = "BAD_MAG"


class StarTable:
    """Ordered mapping column name -> 1-D array (NumPy array or torch tensor, host or device).

    ``bad_mag_columns``: float columns in which NaN stands for the reference's "BAD_MAG".
    """

    def __init__(self, columns, bad_mag_columns=()):
        self._cols = dict(columns)
        self.bad_mag_columns = tuple(c for c in bad_mag_columns if c in self._cols)
        lens = {int(v.shape[0]) for v in self._cols.values()}
        if len(lens) > 1:
            raise ValueError(f"columns of different lengths: {sorted(lens)}")
        self._n = lens.pop() if lens else 0

    # ---- mapping face
    def __len__(self):
        return self._n

    @property
    def columns(self):
        return list(self._cols)

    def __contains__(self, name):
        return name in self._cols

    def __getitem__(self, name):
        """Column as a NumPy array (materialised on the host on first use)."""
        v = self._cols[name]
        if not isinstance(v, np.ndarray):
            v = self._to_host(v)
            self._cols[name] = v
        return v

    def device_column(self, name):
        """The column as stored (a torch tensor if it still lives on the GPU)."""
        return self._cols[name]

    @staticmethod
    def _to_host(t):
        import torch
        if t.device.type == "cuda":
            host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            host.copy_(t, non_blocking=False)
            t = host
        return t.numpy()

    def __repr__(self):
        where = {k: ("device" if not isinstance(v, np.ndarray) and getattr(v, "is_cuda", False)
                     else "host") for k, v in self._cols.items()}
        return f"StarTable({self._n} rows; " + ", ".join(f"{k}[{w}]" for k, w in where.items()) + ")"

    # ---- sinks
    def to_pandas(self, compat=True):
        """DataFrame.  ``compat=True``: the reference's layout -- ``mag_<band>_obs`` object
        columns holding "BAD_MAG" where the noisy flux was not positive (costs one Python
        object per row and column); ``compat=False``: plain float columns, NaN = bad."""
        import pandas as pd
        data = {}
        for k in self._cols:
            v = self[k]
            if compat and k in self.bad_mag_columns:
                col = v.astype(object)
                col[np.isnan(v)] = BAD_MAG
                v = col
            elif v.dtype == np.uint8 and k.startswith("flag"):
                v = v.astype(bool)
            data[k] = v
        return pd.DataFrame(data)

    def to_arrow(self):
        """pyarrow.Table; bad magnitudes are nulls, flags are booleans.  Float columns are
        wrapped without a copy."""
        import pyarrow as pa
        arrays, names = [], []
        for k in self._cols:
            v = self[k]
            if k in self.bad_mag_columns:
                arr = pa.array(v, mask=np.isnan(v))
            elif v.dtype == np.uint8 and k.startswith("flag"):
                arr = pa.array(v.astype(bool))
            else:
                arr = pa.array(v)
            arrays.append(arr)
            names.append(k)
        return pa.Table.from_arrays(arrays, names=names)

    def to_parquet(self, path, compression="zstd", row_group_size=1 << 20):
        import pyarrow.parquet as pq
        pq.write_table(self.to_arrow(), path, compression=compression, row_group_size=row_group_size)
        return path

    def to_csv(self, path, compat=True):
        """CSV through Arrow's writer.  ``compat=True`` writes BAD_MAG where the reference's
        ``DataFrame.to_csv`` would (reference observed.py:961-982); else empty fields."""
        import pyarrow as pa
        import pyarrow.compute as pc
        import pyarrow.csv as pcsv
        t = self.to_arrow()
        for i, field in enumerate(t.schema):
            if compat and pa.types.is_boolean(field.type):         # pandas writes True / False
                t = t.set_column(i, field.name, pc.if_else(t.column(i), "True", "False"))
            if not pa.types.is_floating(field.type):
                continue
            # Arrow's CSV writer prints 15 significant digits; the cast to string gives the
            # shortest digits that round-trip, like pandas' to_csv
            txt = pc.cast(t.column(i), pa.string())
            if compat and field.name in self.bad_mag_columns:
                txt = pc.fill_null(txt, BAD_MAG)
            t = t.set_column(i, field.name, txt)
        pcsv.write_csv(t, path, pcsv.WriteOptions(quoting_style="none"))
        return path
