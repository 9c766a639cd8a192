"""CPU tests of the C++ host mirror (Renumber, MatrixFree stand-in, LaplaceOperator::initialize
tables) against the numpy oracle: the DoF renumbering permutation, ghost index sets, entity
indices and constrained lists must be BIT-EXACT (north_star), for 1, 2, 4 and 8 virtual ranks."""
import numpy as np
import pytest

from oracle import bp4_oracle as O


@pytest.fixture(scope="module")
def host():
    from mf_data_locality_b200 import build, host
    build.build_all()
    return host


CASES = [(2, 5, 1), (3, 6, 1), (4, 7, 1), (5, 4, 1), (6, 6, 1), (7, 3, 1), (8, 4, 1), (4, 9, 1),
         (3, 6, 2), (4, 6, 4), (2, 9, 8), (4, 9, 2), (3, 7, 4), (5, 6, 2)]


@pytest.mark.parametrize("p,s,n_ranks", CASES)
def test_tables_bit_exact(host, p, s, n_ranks):
    rds = O.build_problem(p, s, n_ranks=n_ranks)
    for r, rd in enumerate(rds):
        pr = host.Problem(p, s, device=-1, n_ranks=n_ranks, rank=r)
        assert (pr.n_cells, pr.n_owned, pr.n_ghost) == (rd.n_cells, rd.n_owned, rd.n_ghost)
        assert pr.n_dofs == O.n_dofs_total(p, s)
        assert np.array_equal(pr.node_of_local(), rd.node_of_local.astype(np.uint64))   # permutation + ghosts
        assert np.array_equal(pr.entity_index(), rd.entity_index)
        assert np.array_equal(pr.constrained(), rd.constrained)
        assert np.allclose(pr.vertices(), rd.vertices, rtol=0, atol=1e-15)
        assert pr.n_batches == len(rd.batch_start) - 1 and pr.n_ranges == len(rd.range_start) - 1
        # cell-batch ranges and the DoF runs private to them (the pre/post ranges of the merged loop)
        rc, rp = pr.ranges()
        assert np.array_equal(rc, rd.range_cell_offset) and np.array_equal(rp, rd.range_private_offset)
        assert rp[-1] == rd.group_sizes[0]
        pr.close()


@pytest.mark.parametrize("lanes,bpr", [(4, 1), (8, 3), (2, 5)])
def test_batch_model_parameters(host, lanes, bpr):
    """the deal.II behaviours that are not in the reference tree are explicit parameters"""
    rd = O.build_problem(3, 7, lanes=lanes, batches_per_range=bpr)[0]
    pr = host.Problem(3, 7, device=-1, n_lanes=lanes, batches_per_range=bpr)
    assert np.array_equal(pr.node_of_local(), rd.node_of_local.astype(np.uint64))
    assert np.array_equal(pr.entity_index(), rd.entity_index)
    rc, rp = pr.ranges()
    assert np.array_equal(rc, rd.range_cell_offset) and np.array_equal(rp, rd.range_private_offset)
    pr.close()


def test_private_runs_follow_the_grouping(host):
    """only the cellbatch_range grouping (strategy g = 2) makes the DoFs private to a range one
    contiguous run per range starting at DoF 0; for the other numberings the host hands no
    range tables to the device (the vector updates are then streamed, never mis-hooked)"""
    for r in (1, 2):
        pr = host.Problem(3, 6, device=-1, renumber=(0, r, 2))
        rc, rp = pr.ranges()
        assert len(rc) == pr.n_ranges + 1 and rp[0] == 0 and rp[-1] > 0 and np.all(np.diff(rp.astype(np.int64)) >= 0)
        pr.close()
    pr = host.Problem(3, 6, device=-1, renumber=(0, 1, 0))
    assert len(pr.ranges()[0]) == 0
    pr.close()


STRATEGIES = [(a, r, g) for a in (0, 1) for r in (1, 2) for g in (0, 1, 2)]


@pytest.mark.parametrize("strategy", STRATEGIES)
@pytest.mark.parametrize("p,s,n_ranks", [(3, 6, 1), (2, 7, 2), (4, 6, 4)])
def test_all_renumber_strategies_bit_exact(host, strategy, p, s, n_ranks):
    """every (assembly, renumber, grouping) triple of Renumber's constructor
    (renumber_dofs_for_mf.h:17-21) gives the same permutation and ghost sets in the C++ host
    mirror and in the numpy oracle"""
    rds = O.build_problem(p, s, n_ranks=n_ranks, renumber=strategy)
    for r, rd in enumerate(rds):
        pr = host.Problem(p, s, device=-1, n_ranks=n_ranks, rank=r, renumber=strategy, numbering_only=True)
        assert np.array_equal(pr.node_of_local(), rd.node_of_local.astype(np.uint64))
        pr.close()


def test_neighbour_shortcut_equals_full_numbering(host, monkeypatch):
    """default strategy on several ranks: a process numbers its own rank completely and, of the
    other ranks, only the nodes shared between ranks (from the owner's cell loop order).  Same
    tables as numbering every rank completely (BP4_RENUMBER_ALL_RANKS=1)."""
    for (p, s, n_ranks) in [(3, 8, 8), (4, 7, 4), (2, 8, 2)]:
        for r in range(n_ranks):
            monkeypatch.delenv("BP4_RENUMBER_ALL_RANKS", raising=False)
            a = host.Problem(p, s, device=-1, n_ranks=n_ranks, rank=r)
            monkeypatch.setenv("BP4_RENUMBER_ALL_RANKS", "1")
            b = host.Problem(p, s, device=-1, n_ranks=n_ranks, rank=r)
            assert np.array_equal(a.node_of_local(), b.node_of_local())
            assert np.array_equal(a.entity_index(), b.entity_index())
            for x, y in zip(a.plan(), b.plan()):
                assert np.array_equal(x, y)
            a.close()
            b.close()
    monkeypatch.delenv("BP4_RENUMBER_ALL_RANKS", raising=False)


def test_renumber_strategies(host):
    """base numbering is rejected by the compressed operator, like the reference's AssertThrow
    "Expected contiguous numbering" (poisson_operator.h:198), and so is the cellbatch assembly
    for p >= 3 (it interleaves an entity's nodes over the cells of a batch); for p = 2 every
    entity is a single node and the cellbatch numbering is a valid operator input.  First/last
    touch and the three groupings all deliver contiguous entities."""
    with pytest.raises(host.HostError, match="contiguous"):
        host.Problem(3, 5, device=-1, renumber=(0, 0, 0))
    with pytest.raises(host.HostError, match="contiguous"):
        host.Problem(3, 5, device=-1, renumber=(1, 1, 2))
    pr = host.Problem(2, 6, device=-1, renumber=(1, 1, 2))
    assert pr.n_owned == O.n_dofs_total(2, 6)
    pr.close()
    seen = set()
    for r in (1, 2):
        for g in (0, 1, 2):
            pr = host.Problem(3, 6, device=-1, renumber=(0, r, g))
            nl = pr.node_of_local()
            assert len(np.unique(nl)) == len(nl) == pr.n_owned // 3
            seen.add(nl.tobytes())
            pr.close()
    assert len(seen) >= 4          # the strategies really are different numberings


@pytest.mark.parametrize("p,s,n_ranks", [(3, 6, 1), (2, 7, 2), (5, 4, 1)])
def test_quadratic_mapping_coefficients(host, p, s, n_ranks):
    """mapping_degree = 2: LaplaceOperator::initialize fills all 27 coefficient vectors of every
    cell (the array local_apply evaluates, poisson_operator.h:577-602, :690) -- equal to the
    oracle's to rounding; the tri-linear mapping hands none"""
    for r, rd in enumerate(O.build_problem(p, s, n_ranks=n_ranks, quadratic=True)):
        pr = host.Problem(p, s, device=-1, n_ranks=n_ranks, rank=r, mapping_degree=2)
        c = pr.coefficients()
        assert c.shape == rd.coefficients.shape and np.allclose(c, rd.coefficients, rtol=0, atol=1e-13)
        assert np.array_equal(pr.entity_index(), rd.entity_index)
        pr.close()
    pr = host.Problem(p, s, device=-1)
    assert pr.coefficients() is None
    pr.close()


def test_unsupported_degree(host):
    with pytest.raises(host.HostError, match="degrees 2 to 8"):
        host.Problem(9, 3, device=-1)
