"""CPU emulation of the thread loops of the CUDA cell kernel around the SAME phase functions
(csrc/bp4_cell.cuh) the GPU runs: checks the shared-memory indexing, the gather/scatter table
and the contraction algebra of the device code against the oracle without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import bp4_oracle as O

from helpers import rel_l2

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(HERE, "emu", "libemu_cell.so")
    src = os.path.join(HERE, "emu", "emu_cell.cpp")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
    return C.CDLL(so)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("p,s", [(2, 4), (3, 5), (4, 4), (5, 3), (6, 3), (7, 3), (8, 3)])
def test_device_phases_match_oracle(emu, p, s):
    rd = O.build_problem(p, s)[0]
    t = O.make_tables(p)
    N, Q = p + 1, p + 2
    S, Dn, D, xq, wq = np.zeros((N, Q)), np.zeros((N, Q)), np.zeros((Q, Q)), np.zeros(Q), np.zeros(Q)
    emu.emu_tables(p, _p(S), _p(Dn), _p(D), _p(xq), _p(wq))
    assert np.allclose(S, t.S, atol=1e-14) and np.allclose(D, t.D, rtol=1e-13, atol=1e-12)
    assert np.allclose(xq, t.xq, atol=1e-15) and np.allclose(wq, t.wq, atol=1e-15)
    assert np.allclose(Dn, O.lagrange_derivs(t.xn, t.xq), rtol=1e-13, atol=1e-12)
    v = np.random.default_rng(p).standard_normal(rd.n_owned)
    want = O.vmult_cells(rd, t, v)
    got = np.zeros(rd.n_owned)
    e, vt = np.ascontiguousarray(rd.entity_index), np.ascontiguousarray(rd.vertices)
    assert emu.emu_vmult_cells(p, C.c_long(rd.n_cells), _p(e), _p(vt), _p(v), _p(got)) == 0
    assert rel_l2(got, want) <= 1e-13
    # the fine-grained sweeps of phases 1 and 3 used for the high degrees
    got2 = np.zeros(rd.n_owned)
    assert emu.emu_vmult_cells_fine(p, C.c_long(rd.n_cells), _p(e), _p(vt), _p(v), _p(got2)) == 0
    assert rel_l2(got2, want) <= 1e-13


@pytest.mark.parametrize("p", [2, 3, 4, 5, 6, 7, 8])
def test_even_odd_contractions_match_dense_sums(emu, p):
    """eo_first / eo_second (bp4_cell.cuh) against the plain sums for S, Dn, D, both directions:
    exercises the odd/even node counts, the middle row/column cases and both symmetry signs."""
    emu.emu_eo_check.restype = C.c_double
    assert emu.emu_eo_check(p) <= 2e-14
