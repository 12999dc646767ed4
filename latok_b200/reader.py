"""Native csv / csv.gz ingest (latok_reader.cpp): rows -> packed batches without a Python object per row.

Counterpart of the reference's loop `for row in csv.reader(f): text = json.loads(row[1]).strip()`
(scripts/timing/time_tokenizer.py:25-40).  Host code only; the batches go to Engine.submit().
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class CsvReader:
    """Iterate over (uint8 buffer, int64 offsets[rows+1]) batches of a csv / csv.gz file.

    With pinned=True the buffers are page-locked (latok_b200_host_alloc) so Engine.submit() copies straight from
    them; `n_buffers` buffer sets rotate, i.e. a yielded batch stays valid until n_buffers - 1 further batches have
    been read."""

    def __init__(self, path: str, batch_rows: int = 200_000, batch_bytes: int = 64 << 20, column: int = 1,
                 pinned: bool = False, n_buffers: int = 2):
        self._L = _lib.load()
        h = C.c_void_p()
        _lib.check(self._L.latok_b200_reader_open(str(path).encode(), column, C.byref(h)))
        self._h = h
        self.batch_rows, self.batch_bytes = batch_rows, batch_bytes
        self._pinned_ptrs = []
        self._sets = [self._alloc(pinned) for _ in range(max(1, n_buffers))]
        self._turn = 0

    def _alloc(self, pinned):
        if not pinned:
            return np.empty(self.batch_bytes, dtype=np.uint8), np.empty(self.batch_rows + 1, dtype=np.int64)
        out = []
        for nbytes, dt in ((self.batch_bytes, np.uint8), (8 * (self.batch_rows + 1), np.int64)):
            p = C.c_void_p()
            _lib.check(self._L.latok_b200_host_alloc(C.byref(p), nbytes))
            self._pinned_ptrs.append(p)
            out.append(np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,)).view(dt))
        return tuple(out)

    def __iter__(self):
        return self

    def __next__(self):
        buf, off = self._sets[self._turn]
        self._turn = (self._turn + 1) % len(self._sets)
        n = C.c_int64(0)
        _lib.check(self._L.latok_b200_reader_next(self._h, self.batch_rows, buf.ctypes.data, len(buf), off.ctypes.data,
                                                  C.byref(n)))
        if n.value == 0:
            raise StopIteration
        return buf[:int(off[n.value])], off[:n.value + 1]

    def close(self):
        if getattr(self, "_h", None):
            self._L.latok_b200_reader_close(self._h)
            self._h = None
        for p in self._pinned_ptrs:
            self._L.latok_b200_host_free(p)
        self._pinned_ptrs = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
