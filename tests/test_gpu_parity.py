"""GPU parity tests: the CUDA path (through the C ABI of include/bp4.h) against the CPU
oracle on the same inputs.  Tolerances are BASELINE.json north_star's: operator apply
rel-L2 <= 1e-12, CG iteration count +-1, solution rel diff <= 1e-8."""
import numpy as np
import pytest

from oracle import bp4_oracle as O

from helpers import gpu_cg_merged, gpu_cg_plain, make_ctx, rel_l2, single

pytestmark = pytest.mark.gpu

VMULT_CASES = [(2, 3), (2, 7), (3, 3), (3, 6), (3, 10), (4, 4), (4, 9), (4, 11), (5, 5), (5, 7), (6, 3),
               (6, 8), (7, 4), (7, 6), (8, 3), (8, 7)]


@pytest.mark.parametrize("p,s", VMULT_CASES)
def test_vmult_matches_oracle(p, s, bp4_lib, c_oracle_lib):
    rd, co = single(p, s)
    ctx = make_ctx(rd)
    rng = np.random.default_rng(100 * p + s)
    v = rng.standard_normal(rd.n_owned)
    src, dst = ctx.vector(data=v), ctx.vector()
    ctx.vmult(dst, src)
    got = dst.download()
    want = co.vmult(v)
    assert rel_l2(got, want) <= 1e-12
    # constrained rows are the identity (poisson_operator.h:311-312)
    assert np.array_equal(got[rd.constrained], v[rd.constrained])
    ctx.close()


def test_vmult_zero_and_linearity(bp4_lib):
    rd, co = single(4, 6)
    ctx = make_ctx(rd)
    rng = np.random.default_rng(1)
    a, b = rng.standard_normal(rd.n_owned), rng.standard_normal(rd.n_owned)
    va, vb, vab, out = ctx.vector(data=a), ctx.vector(data=b), ctx.vector(data=2 * a - 3 * b), ctx.vector()
    ctx.vmult(out, va)
    ya = out.download()
    ctx.vmult(out, vb)
    yb = out.download()
    ctx.vmult(out, vab)
    yab = out.download()
    assert rel_l2(yab, 2 * ya - 3 * yb) <= 1e-12
    z = ctx.vector()
    ctx.vmult(out, z)
    assert not out.download().any()
    # symmetry u^T A v = v^T A u
    assert abs(a @ yb - b @ ya) <= 1e-11 * abs(a @ yb)
    ctx.close()


@pytest.mark.parametrize("p,s", [(2, 6), (3, 6), (4, 6), (5, 4), (6, 5), (8, 3)])
def test_inverse_diagonal(p, s, bp4_lib):
    rd, _ = single(p, s)
    ctx = make_ctx(rd)
    got = ctx.inverse_diagonal().download()
    want = O.finish_inverse_diagonal(O.inverse_diagonal(rd))
    assert rel_l2(got, want) <= 1e-12
    ctx.close()


def test_blas1(bp4_lib):
    rd, _ = single(3, 6)
    ctx = make_ctx(rd)
    n = rd.n_owned
    rng = np.random.default_rng(7)
    a, b = rng.standard_normal(n), rng.standard_normal(n)
    va, vb, vc = ctx.vector(data=a), ctx.vector(data=b), ctx.vector()
    assert abs(ctx.dot(va, vb) - a @ b) <= 1e-12 * np.linalg.norm(a) * np.linalg.norm(b)
    assert abs(ctx.l2_norm(va) - np.linalg.norm(a)) <= 1e-13 * np.linalg.norm(a)
    ctx.equ(vc, -1.5, va)
    assert np.array_equal(vc.download(), -1.5 * a)
    ctx.add(vc, 0.25, vb)
    np.testing.assert_allclose(vc.download(), -1.5 * a + 0.25 * b, rtol=1e-15, atol=1e-15)
    ctx.sadd(vc, 2.0, -1.0, va)
    np.testing.assert_allclose(vc.download(), 2.0 * (-1.5 * a + 0.25 * b) - a, rtol=1e-14, atol=1e-14)
    assert ctx.all_zero(ctx.vector()) and not ctx.all_zero(va)
    g = ctx.vector(data=a)
    r = ctx.add_and_dot(g, 0.5, vb, g)
    assert abs(r - (a + 0.5 * b) @ (a + 0.5 * b)) <= 1e-12 * r
    np.testing.assert_allclose(g.download(), a + 0.5 * b, rtol=1e-15, atol=1e-15)
    diag = rng.standard_normal(n // 3)
    vd = ctx.vector(n // 3, data=diag)
    ctx.jacobi_vmult(vc, va, vd)
    np.testing.assert_allclose(vc.download(), np.repeat(diag, 3) * a, rtol=1e-15)
    ctx.close()


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("constrained_rhs", [False, True])
@pytest.mark.parametrize("p,s", [(2, 7), (3, 6), (4, 7), (4, 9), (5, 6), (6, 4), (7, 5), (8, 4)])
def test_merged_sums_match_oracle(p, s, fused, constrained_rhs, bp4_lib, c_oracle_lib):
    """one vmult_with_merged_sums call in each of the three do_cg_update4b regimes, with the
    vector updates inside the cell loop (fused, default) and streamed (unfused).  With
    constrained_rhs the Dirichlet rows of x, g, d, h are non-zero: do_cg_update4b/3b sweep ALL
    owned entries (solver_cg_optimized.h:65-161, 12-61), not only those the cells touch."""
    rd, co = single(p, s)
    ctx = make_ctx(rd)
    _, n_private, n_units = ctx.fused_info()
    assert n_units > 0
    if p <= 4:
        assert n_private == rd.group_sizes[0]
        ctx.set_fused(fused)
    else:
        # no fused kernel above Q4 (a range's private run does not fit the staging rows left
        # in shared memory): everything is streamed, asking for the fused loop is an error
        from mf_data_locality_b200 import capi
        assert n_private == 0
        with pytest.raises(capi.Bp4Error, match="degree"):
            ctx.set_fused(True)
        if fused:
            ctx.close()
            return
    n = rd.n_owned
    rng = np.random.default_rng(3)
    free = np.ones(n)
    if not constrained_rhs:
        free[rd.constrained] = 0.0
    prec = O.finish_inverse_diagonal(O.inverse_diagonal(rd))
    prec3 = np.repeat(prec, 3)
    vp = ctx.vector(n // 3, data=prec)
    for (alpha, beta, alpha_old, beta_old) in [(0.0, 0.0, 0.0, 0.0), (0.7, 0.3, 0.0, 0.2), (0.7, 0.3, 0.4, 0.2)]:
        x, g, d, h = (rng.standard_normal(n) * free for _ in range(4))
        vx, vg, vd, vh = (ctx.vector(data=a) for a in (x, g, d, h))
        S = ctx.vmult_merged(vx, vg, vd, vh, vp, alpha, beta, alpha_old, beta_old)
        O.cg_update4b(h, x, g, d, prec3, alpha, beta, alpha_old, beta_old)
        dd = d.copy()
        dd[rd.constrained] = 0.0          # constrained entries are never read by the cells
        h[:] = co.vmult_cells(dd)
        want = O.cg_update3b(g, d, h, prec3)
        np.testing.assert_allclose(S, want, rtol=1e-11)
        for got, ref in ((vx, x), (vg, g), (vd, d), (vh, h)):
            assert rel_l2(got.download(), ref) <= 1e-12
    ctx.close()


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("p,s", [(3, 6), (4, 6), (2, 9), (5, 7), (6, 6)])
def test_cg_merged_parity(p, s, fused, bp4_lib, c_oracle_lib):
    rd, co = single(p, s)
    ctx = make_ctx(rd)
    if fused and p > 4:
        pytest.skip("no fused kernel above Q4")
    ctx.set_fused(fused)
    prec = O.finish_inverse_diagonal(O.inverse_diagonal(rd))
    vp = ctx.vector(rd.n_owned // 3, data=prec)
    for reduce in (1e-8, 1e-10):
        ctl = O.ReductionControl(100, 1e-15, reduce)
        x = gpu_cg_merged(ctx, rd.rhs, vp, ctl).download()
        xo, ito, hist = co.cg(rd.rhs, prec, merged=True, reduce=reduce)
        assert abs(ctl.last_step - ito) <= 1
        if ito < 100:
            assert rel_l2(x, xo) <= 1e-8
        else:
            assert rel_l2(x, xo) <= 1e-6
    ctx.close()


@pytest.mark.parametrize("p,s", [(3, 6), (4, 6)])
def test_cg_plain_parity(p, s, bp4_lib, c_oracle_lib):
    rd, co = single(p, s)
    ctx = make_ctx(rd)
    prec = O.finish_inverse_diagonal(O.inverse_diagonal(rd))
    vp = ctx.vector(rd.n_owned // 3, data=prec)
    ctl = O.ReductionControl(100, 1e-15, 1e-8)
    x = gpu_cg_plain(ctx, rd.rhs, vp, ctl).download()
    xo, ito, hist = co.cg(rd.rhs, prec, merged=False)
    assert abs(ctl.last_step - ito) <= 1
    assert rel_l2(x, xo) <= (1e-8 if ito < 100 else 1e-6)
    ctx.close()


def test_error_behaviour(bp4_lib):
    """argument errors are reported through negative codes + bp4_last_error, with the
    reference's wording where it has one (diagonal_matrix_blocked.h:17-20)"""
    from mf_data_locality_b200 import capi
    rd, _ = single(3, 5)
    ctx = make_ctx(rd)
    short, full = ctx.vector(10), ctx.vector()
    with pytest.raises(capi.Bp4Error, match="n_owned"):
        ctx.vmult(full, short)
    with pytest.raises(capi.Bp4Error, match="aliases"):
        ctx.vmult(full, full)
    with pytest.raises(capi.Bp4Error, match="Dimension mismatch"):
        ctx.jacobi_vmult(full, full, ctx.vector(3))
    with pytest.raises(capi.Bp4Error, match="degree"):
        capi.Context(9, rd.entity_index, rd.vertices, rd.n_owned)
    with pytest.raises(capi.Bp4Error, match="multiples of 3"):
        capi.Context(3, rd.entity_index, rd.vertices, rd.n_owned + 1)
    ctx.close()


def test_empty_and_ragged_inputs(bp4_lib, c_oracle_lib):
    """a context without cells applies the identity on constrained rows and zero elsewhere; a cell
    count that is not a multiple of the block's batch size is handled (ragged last batch)"""
    from mf_data_locality_b200 import capi
    ctx = capi.Context(4, np.zeros((0, 27), np.uint32), np.zeros((0, 8, 3)), 30, 0, np.arange(6, dtype=np.uint32))
    v = np.arange(30, dtype=np.float64) + 1
    src, dst = ctx.vector(data=v), ctx.vector()
    ctx.vmult(dst, src)
    got = dst.download()
    assert np.array_equal(got[:6], v[:6]) and not got[6:].any()
    ctx.close()
    for p, s in [(4, 3), (3, 4), (5, 3)]:          # 8, 16, 8 cells: never a multiple of 7 / 10 / 5
        rd, co = single(p, s)
        ctx = make_ctx(rd)
        w = np.random.default_rng(p).standard_normal(rd.n_owned)
        a, b = ctx.vector(data=w), ctx.vector()
        ctx.vmult(b, a)
        assert rel_l2(b.download(), co.vmult(w)) <= 1e-12
        ctx.close()


def test_fused_needs_range_tables(bp4_lib):
    """without range tables in the descriptor the vector updates are streamed; asking for the
    in-loop form is a state error, not a silent fallback"""
    from mf_data_locality_b200 import capi
    rd, _ = single(3, 5)
    ctx = make_ctx(rd, ranges=False)
    assert ctx.fused_info() == (False, 0, 0)
    with pytest.raises(capi.Bp4Error, match="range tables"):
        ctx.set_fused(True)
    ctx.close()
    bad = rd.range_private_offset.copy()
    bad[1] += 1                                    # not a multiple of 3
    with pytest.raises(capi.Bp4Error, match="multiple of 3"):
        capi.Context(rd.degree, rd.entity_index, rd.vertices, rd.n_owned, rd.n_ghost, rd.constrained,
                     ranges=(rd.range_cell_offset, bad))


@pytest.mark.parametrize("p,s", [(2, 6), (3, 6), (4, 7), (5, 5), (6, 5), (7, 4), (8, 4)])
def test_quadratic_geometry_matches_oracle(p, s, bp4_lib):
    """genuinely quadratic cells: all 27 coefficient vectors per cell cross the ABI
    (bp4_desc::coefficients) and the kernel evaluates the full form of
    poisson_operator.h:577-602.  Operator apply, GLL inverse diagonal and one merged iteration
    against the numpy oracle (which is pinned by a dense tri-quadratic assembly)."""
    from mf_data_locality_b200 import capi
    rd = O.build_problem(p, s, quadratic=True)[0]
    t = O.make_tables(p)
    ctx = make_ctx(rd)
    v = np.random.default_rng(p).standard_normal(rd.n_owned)
    src, dst = ctx.vector(data=v), ctx.vector()
    ctx.vmult(dst, src)
    got = dst.download()
    want = O.vmult(rd, t, v)
    assert rel_l2(got, want) <= 1e-12
    lin = O.vmult(O.build_problem(p, s)[0], t, v)
    assert rel_l2(got, lin) > 1e-3                      # not the tri-linear operator
    prec = O.finish_inverse_diagonal(O.inverse_diagonal(rd))
    assert rel_l2(ctx.inverse_diagonal().download(), prec) <= 1e-12
    # the in-loop vector updates are built for the tri-linear kernel only: asking is an error
    assert ctx.fused_info()[0] is False
    with pytest.raises(capi.Bp4Error):
        ctx.set_fused(True)
    n = rd.n_owned
    rng = np.random.default_rng(9)
    free = np.ones(n)
    free[rd.constrained] = 0.0
    prec3 = np.repeat(prec, 3)
    vp = ctx.vector(n // 3, data=prec)
    x, g, d, h = (rng.standard_normal(n) * free for _ in range(4))
    vx, vg, vd, vh = (ctx.vector(data=a) for a in (x, g, d, h))
    S = ctx.vmult_merged(vx, vg, vd, vh, vp, 0.7, 0.3, 0.4, 0.2)
    O.cg_update4b(h, x, g, d, prec3, 0.7, 0.3, 0.4, 0.2)
    h[:] = O.vmult_cells(rd, t, d)
    np.testing.assert_allclose(S, O.cg_update3b(g, d, h, prec3), rtol=1e-11)
    ctx.close()


def test_quadratic_geometry_through_the_host_mirror(bp4_lib):
    """mapping_degree = 2 in the C++ mirror: set-up, operator and the merged plugin on quadratic cells"""
    from mf_data_locality_b200 import host
    p, s = 4, 7
    rd = O.build_problem(p, s, quadratic=True)[0]
    t = O.make_tables(p)
    prob = host.Problem(p, s, plugin="merged", device=0, mapping_degree=2)
    v = np.random.default_rng(1).standard_normal(rd.n_owned)
    assert rel_l2(prob.vmult(v), O.vmult(rd, t, v)) <= 1e-12
    prec = O.finish_inverse_diagonal(O.inverse_diagonal(rd))
    assert rel_l2(prob.diagonal(), prec) <= 1e-12
    ctl = O.ReductionControl(100, 1e-15, 1e-8)
    xo = O.solver_cg_merged(lambda w: O.vmult_cells(rd, t, w), np.zeros(rd.n_owned), rd.rhs, prec, ctl)
    x, it = prob.run_cg_solver(rd.rhs)
    assert abs(it - ctl.last_step) <= 1
    assert rel_l2(x, xo) <= (1e-8 if ctl.last_step < 100 else 1e-6)
    prob.close()


# ---- parity at the sizes BASELINE.json quotes (configs[0..2]) ------------------------------
SCALE_CASES = [pytest.param(3, 15, False, id="config0-Q3-s15-plain"),
               pytest.param(4, 18, True, id="config1-Q4-s18-merged"),
               pytest.param(6, 17, True, id="config2-Q6-s17-merged")]


@pytest.mark.parametrize("p,s,merged", SCALE_CASES)
def test_benchmark_scale_parity(p, s, merged, bp4_lib, c_oracle_lib):
    """BASELINE.json's own sizes (Q6 at s=17, the largest the CPU oracle finishes in a minute):
    one operator apply (rel-L2 <= 1e-12) and the benchmark's 100-iteration CG through the C++
    plugin (iteration count +-1, solution <= 1e-6 at the iteration cap) against the C oracle.
    Exercises 32-bit index arithmetic at 5e7..1.7e8 entries, the persistent grid's tail and
    the drift of 100 iterations at full size."""
    from mf_data_locality_b200 import host
    rd, co = single(p, s)
    prob = host.Problem(p, s, plugin="merged" if merged else "plain", device=0)
    assert prob.n_owned == rd.n_owned
    v = np.random.default_rng(p * 100 + s).standard_normal(rd.n_owned)
    got = prob.vmult(v)
    want = co.vmult(v)
    assert rel_l2(got, want) <= 1e-12
    assert np.array_equal(got[rd.constrained], v[rd.constrained])
    del got, want, v
    prec = O.finish_inverse_diagonal(O.inverse_diagonal(rd))
    x, it = prob.run_cg_solver(rd.rhs)
    xo, ito, _ = co.cg(rd.rhs, prec, merged=merged)
    assert abs(it - ito) <= 1
    assert rel_l2(x, xo) <= (1e-8 if ito < 100 else 1e-6)
    prob.close()


@pytest.mark.parametrize("p,s", [(4, 16), (3, 16), (2, 16)])
def test_fused_cg_at_medium_scale(p, s, bp4_lib, c_oracle_lib):
    """the in-loop form of the merged iteration (vector updates staged through shared memory
    inside the cell kernel) with several units per thread block: descriptor ring wrap-around,
    unit claims, two jobs per run - the 100-iteration solve through the C++ plugin against the C
    oracle, and the same solve with the streamed form"""
    from mf_data_locality_b200 import capi, host
    rd, co = single(p, s)
    prob = host.Problem(p, s, plugin="merged", device=0)
    ctx = capi.Context.from_handle(prob.ctx_handle(), p, prob.n_cells, prob.n_owned, prob.n_ghost)
    _, n_private, n_units = ctx.fused_info()
    assert n_private == rd.group_sizes[0] and n_units > 2 * 296
    prec = O.finish_inverse_diagonal(O.inverse_diagonal(rd))
    xo, ito, _ = co.cg(rd.rhs, prec, merged=True)
    sols = []
    for fused in (True, False):
        ctx.set_fused(fused)
        assert ctx.fused_info()[0] is fused
        x, it = prob.run_cg_solver(rd.rhs)
        assert abs(it - ito) <= 1
        assert rel_l2(x, xo) <= (1e-8 if ito < 100 else 1e-6)
        sols.append(x)
    assert rel_l2(sols[0], sols[1]) <= 1e-6
    prob.close()


def test_default_form_of_the_merged_iteration(bp4_lib, monkeypatch):
    """which form vmult_with_merged_sums starts with follows the B200 measurements:
    vector updates inside the cell loop at Q4 on one rank, streamed at the other degrees;
    BP4_FUSED pins either"""
    monkeypatch.delenv("BP4_FUSED", raising=False)
    for p, want in ((4, True), (3, False), (2, False), (5, False)):
        rd, _ = single(p, 5)
        ctx = make_ctx(rd)
        assert ctx.fused_info()[0] is want
        ctx.close()
    monkeypatch.setenv("BP4_FUSED", "0")
    rd, _ = single(4, 5)
    ctx = make_ctx(rd)
    assert ctx.fused_info()[0] is False
    ctx.close()
    monkeypatch.setenv("BP4_FUSED", "1")
    rd, _ = single(3, 5)
    ctx = make_ctx(rd)
    assert ctx.fused_info()[0] is True
    ctx.close()
