"""ctypes loader for oracle/libbp4_oracle.so (C/OpenMP restatement; test infrastructure
only -- see bp4_oracle.c).  Build with `make -C oracle`."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import bp4_oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _Tables(C.Structure):
    _fields_ = [("degree", C.c_int), ("S", C.c_void_p), ("D", C.c_void_p), ("xq", C.c_void_p),
                ("wq", C.c_void_p), ("ent", C.c_void_p), ("pos", C.c_void_p),
                ("n_colors", C.c_int), ("color_start", C.c_void_p), ("color_cells", C.c_void_p)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libbp4_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.oracle_num_threads.restype = C.c_int
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class COracle:
    """Holds the setup arrays of one rank (from bp4_oracle.build_problem) in C layout."""

    def __init__(self, rd: O.RankData):
        p = rd.degree
        self.rd = rd
        t = O.make_tables(p)
        n1 = p + 1
        ent = np.zeros(n1 ** 3, dtype=np.int32)
        pos = np.zeros(n1 ** 3, dtype=np.int32)

        def code(i):
            return 0 if i == 0 else (2 if i == p else 1)

        def off(i):
            return i - 1 if 0 < i < p else 0

        def size(e):
            return p - 1 if e == 1 else 1
        l = 0
        for k in range(n1):
            for j in range(n1):
                for i in range(n1):
                    ex, ey, ez = code(i), code(j), code(k)
                    ent[l] = ex + 3 * ey + 9 * ez
                    pos[l] = off(i) + size(ex) * (off(j) + size(ey) * off(k))
                    l += 1
        self._keep = [np.ascontiguousarray(a) for a in (t.S, t.D, t.xq, t.wq, ent, pos)]
        # 8-colouring of the structured mesh by lattice parity: cells of one colour share no DoF
        col = (rd.cells[:, 0] % 2) + 2 * (rd.cells[:, 1] % 2) + 4 * (rd.cells[:, 2] % 2)
        order = np.argsort(col, kind="stable").astype(np.int64)
        start = np.concatenate(([0], np.cumsum(np.bincount(col, minlength=8)))).astype(np.int64)
        self._keep += [np.ascontiguousarray(start), np.ascontiguousarray(order)]
        self.tab = _Tables(p, *[_p(a) for a in self._keep[:6]], 8, _p(self._keep[6]), _p(self._keep[7]))
        self.eidx = np.ascontiguousarray(rd.entity_index, dtype=np.uint32)
        assert rd.coefficients is None, "the C restatement evaluates tri-linear cells only (use the numpy oracle)"
        self.coef = np.ascontiguousarray(O.trilinear_coefficients(rd.vertices))
        self.con = np.ascontiguousarray(rd.constrained, dtype=np.uint32)
        self.n = rd.n_owned + rd.n_ghost

    def vmult_cells(self, src):
        src = np.ascontiguousarray(src, dtype=np.float64)
        dst = np.empty(self.n)
        e = lib().oracle_vmult_cells(C.byref(self.tab), C.c_long(self.rd.n_cells), C.c_long(self.n),
                                     _p(self.eidx), _p(self.coef), _p(src), _p(dst))
        assert e == 0
        return dst

    def vmult(self, src):
        src = np.ascontiguousarray(src, dtype=np.float64)
        dst = np.empty(self.n)
        e = lib().oracle_vmult(C.byref(self.tab), C.c_long(self.rd.n_cells), C.c_long(self.n),
                               _p(self.eidx), _p(self.coef), C.c_long(len(self.con)), _p(self.con),
                               _p(src), _p(dst))
        assert e == 0
        return dst

    def cg(self, b, diag, merged: bool, max_steps=100, tol=1e-15, reduce=1e-8, blocked=False):
        """returns (x, last_step, residual history).  blocked (merged only): the vector updates
        run per cell-batch range inside the loop, every thread on its own chunk of ranges -- the
        reference's cache-blocked cell_loop with pre/post hooks; used as the CPU baseline"""
        b = np.ascontiguousarray(b, dtype=np.float64)
        diag = np.ascontiguousarray(diag, dtype=np.float64)
        x = np.zeros(self.n)
        hist = np.zeros(max_steps + 2)
        if merged and blocked:
            assert self.rd.n_ghost == 0
            rc = np.ascontiguousarray(self.rd.range_cell_offset, dtype=np.int64)
            rp = np.ascontiguousarray(self.rd.range_private_offset, dtype=np.int64)
            it = lib().oracle_cg_merged_blocked(C.byref(self.tab), C.c_long(self.rd.n_cells), C.c_long(self.n),
                                                _p(self.eidx), _p(self.coef), _p(diag), _p(b), _p(x),
                                                C.c_int(max_steps), C.c_double(tol), C.c_double(reduce),
                                                _p(hist), C.c_long(len(rc) - 1), _p(rc), _p(rp))
            assert it >= 0
        elif merged:
            it = lib().oracle_cg_merged(C.byref(self.tab), C.c_long(self.rd.n_cells), C.c_long(self.n),
                                        _p(self.eidx), _p(self.coef), _p(diag), _p(b), _p(x),
                                        C.c_int(max_steps), C.c_double(tol), C.c_double(reduce), _p(hist))
        else:
            it = lib().oracle_cg_plain(C.byref(self.tab), C.c_long(self.rd.n_cells), C.c_long(self.n),
                                       _p(self.eidx), _p(self.coef), C.c_long(len(self.con)), _p(self.con),
                                       _p(diag), _p(b), _p(x), C.c_int(max_steps), C.c_double(tol),
                                       C.c_double(reduce), _p(hist))
        return x, it, hist[: it + 1]
