"""Synthetic MSAreal generator (DataSimulator.py-shaped); binding of include/rr_msagen.h."""
import ctypes as C
import os

import numpy as np

from ._lib import MsagenParams, gen

TYPES = {"Tree": 0, "Distributed": 1, "EquiDistant": 2}


class MsaGen:
    def __init__(self, type="Tree", copies=100, coverage=40, repeat_len=30000, diff=0.01, seed=1001,
                 flank=10000, min_overlap=500, max_reads=0, threads=0):
        p = MsagenParams(TYPES[type], copies, coverage, repeat_len, diff, seed, flank, min_overlap, max_reads, threads)
        self._h = C.c_void_p(gen.rr_msagen_create(C.byref(p)))
        self.rows = gen.rr_msagen_rows(self._h)
        self.cols = gen.rr_msagen_cols(self._h)

    def codes(self, out=None):
        if out is None:
            out = np.empty((self.rows, self.cols), dtype=np.uint8)
        assert out.shape == (self.rows, self.cols) and out.dtype == np.uint8 and out.flags.c_contiguous
        gen.rr_msagen_fill_codes(self._h, out.ctypes.data)
        return out

    def text(self):
        buf = np.empty((self.rows, self.cols + 1), dtype=np.uint8)
        gen.rr_msagen_fill_text(self._h, buf.ctypes.data)
        return buf.tobytes()

    def write(self, path):
        rc = gen.rr_msagen_write(self._h, os.fsencode(path))
        if rc:
            raise OSError(f"rr_msagen_write({path}) -> {rc}")

    def read_copy(self):
        return np.ctypeslib.as_array(gen.rr_msagen_read_copy(self._h), shape=(self.rows,)).copy()

    def close(self):
        if self._h:
            gen.rr_msagen_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
