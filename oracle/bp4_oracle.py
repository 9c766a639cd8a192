"""CPU ORACLE (test infrastructure, NOT product code) for the CEED BP4 hot path of
peterrum/mf_data_locality.

This file is a numpy restatement of the reference algorithm.  It may only be
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs -- never by mf_data_locality_b200/ (the product).

PARITY UNPINNED: the reference tree holds no tests, golden vectors or fixtures,
and cannot be compiled here (it needs deal.II >= 9.3 + p4est + MPI, all absent).
The oracle is therefore pinned by independent mathematics instead (dense
brute-force assembly of the same bilinear form, symmetry, null space, merged ==
plain CG iterates) -- see tests/test_oracle_*.py and DESIGN.md.

Every function cites the reference file:line it follows (paths relative to
/root/reference/).  The arithmetic that lives in deal.II (not vendored in the
reference; `FIND_PACKAGE(deal.II 9.3.0 ...)`, CMakeLists.txt:3) is restated from
its published algorithms: Gauss/Gauss-Lobatto quadrature, Lagrange bases,
sum-factorised evaluation, SolverCG, MatrixFree cell batching.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

INVALID = np.uint32(0xFFFFFFFF)

# --------------------------------------------------------------------------- #
# 1-D tables (deal.II QGauss<1>, QGaussLobatto<1>, FE_Q Lagrange basis)
# --------------------------------------------------------------------------- #


def gauss_01(n: int):
    """Gauss-Legendre points/weights on [0,1] (deal.II QGauss<1>(n);
    used at poisson_operator.h:107 with n = n_q_points_1d)."""
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def gauss_lobatto_01(n: int):
    """Gauss-Lobatto points/weights on [0,1] (deal.II QGaussLobatto<1>(n);
    FE_Q(p) support points are GLL(p+1); benchmark.h:129 uses it as quadrature)."""
    assert n >= 2
    if n == 2:
        return np.array([0.0, 1.0]), np.array([0.5, 0.5])
    # interior points = roots of P'_{n-1}
    c = np.zeros(n)
    c[-1] = 1.0
    dc = np.polynomial.legendre.legder(c)
    xi = np.sort(np.real(np.polynomial.legendre.legroots(dc)))
    # Newton polish on P'_{n-1}
    ddc = np.polynomial.legendre.legder(dc)
    for _ in range(3):
        xi = xi - np.polynomial.legendre.legval(xi, dc) / np.polynomial.legendre.legval(xi, ddc)
    x = np.concatenate(([-1.0], xi, [1.0]))
    # symmetrise
    x = 0.5 * (x - x[::-1])
    pn = np.polynomial.legendre.legval(x, c)
    w = 2.0 / (n * (n - 1) * pn * pn)
    return 0.5 * (x + 1.0), 0.5 * w


def lagrange_values(nodes: np.ndarray, pts: np.ndarray) -> np.ndarray:
    """M[i][q] = l_i(pts[q]) for the Lagrange basis on `nodes`."""
    n = len(nodes)
    M = np.ones((n, len(pts)))
    for i in range(n):
        for j in range(n):
            if j != i:
                M[i] *= (pts - nodes[j]) / (nodes[i] - nodes[j])
    return M


def lagrange_derivs(nodes: np.ndarray, pts: np.ndarray) -> np.ndarray:
    """M[i][q] = l_i'(pts[q])."""
    n = len(nodes)
    M = np.zeros((n, len(pts)))
    for i in range(n):
        for k in range(n):
            if k == i:
                continue
            term = np.ones(len(pts)) / (nodes[i] - nodes[k])
            for j in range(n):
                if j != i and j != k:
                    term *= (pts - nodes[j]) / (nodes[i] - nodes[j])
            M[i] += term
    return M


@dataclass
class Tables:
    """1-D matrices used by the operator.
    S[i][q]: GLL(p+1) Lagrange basis at Gauss(q) points (shape_values_eo packing of
             the same numbers, poisson_operator.h:461, :549).
    D[i][q]: derivative of the Gauss-point Lagrange basis at the Gauss points
             (shape_gradients_collocation_eo, poisson_operator.h:553)."""
    degree: int
    n_q: int
    xq: np.ndarray
    wq: np.ndarray
    xn: np.ndarray
    S: np.ndarray
    D: np.ndarray


def make_tables(degree: int, n_q: int | None = None, quad: str = "gauss") -> Tables:
    n_q = degree + 2 if n_q is None else n_q
    xn, _ = gauss_lobatto_01(degree + 1)
    if quad == "gauss":
        xq, wq = gauss_01(n_q)
    else:
        xq, wq = gauss_lobatto_01(n_q)
    S = lagrange_values(xn, xq)
    D = lagrange_derivs(xq, xq)
    return Tables(degree, n_q, xq, wq, xn, S, D)


# --------------------------------------------------------------------------- #
# mesh (benchmark.h:66-88, curved_manifold.h:25-35)
# --------------------------------------------------------------------------- #


def push_forward(p: np.ndarray) -> np.ndarray:
    """MyManifold::push_forward, curved_manifold.h:25-35: x + 0.1*prod_d sin(pi x_d)."""
    sinval = 0.1 * np.prod(np.sin(math.pi * p), axis=-1, keepdims=True)
    return p + sinval


def mesh_dims(s: int):
    """benchmark.h:67-88: n_refine = s/3, remainder = s%3; subdivisions 2 in the
    first `remainder` directions."""
    n_refine, rem = divmod(s, 3)
    sub = [2 if d < rem else 1 for d in range(3)]
    n = [sub[d] << n_refine for d in range(3)]
    return n_refine, sub, n


def active_cell_order(s: int) -> np.ndarray:
    """Lattice coordinates (cx,cy,cz) of the active cells in deal.II's traversal
    order after refine_global: coarse cells x-fastest, children
    recursively by child index cx + 2 cy + 4 cz (Morton order inside a coarse cell)."""
    n_refine, sub, _ = mesh_dims(s)
    m = np.arange(8 ** n_refine, dtype=np.int64)
    lx = np.zeros_like(m)
    ly = np.zeros_like(m)
    lz = np.zeros_like(m)
    for b in range(n_refine):
        lx |= ((m >> (3 * b)) & 1) << b
        ly |= ((m >> (3 * b + 1)) & 1) << b
        lz |= ((m >> (3 * b + 2)) & 1) << b
    out = []
    side = 1 << n_refine
    for kz in range(sub[2]):
        for ky in range(sub[1]):
            for kx in range(sub[0]):
                out.append(np.stack([lx + kx * side, ly + ky * side, lz + kz * side], axis=1))
    return np.concatenate(out, axis=0)


def cell_vertices(s: int, cells: np.ndarray) -> np.ndarray:
    """[n_cells][8][3] vertex coordinates, deal.II vertex order v = x + 2y + 4z
    (poisson_operator.h:153-160).  Vertices are push_forward(lattice point)
    (GridTools::transform at benchmark.h:84 for the coarse mesh; refinement through
    the ChartManifold reproduces the lattice up to its Newton tolerance)."""
    n_refine, _, _ = mesh_dims(s)
    h = 1.0 / (1 << n_refine)
    off = np.array([[(v >> 0) & 1, (v >> 1) & 1, (v >> 2) & 1] for v in range(8)], dtype=np.float64)
    lat = (cells[:, None, :].astype(np.float64) + off[None, :, :]) * h
    return push_forward(lat)


def trilinear_coefficients(verts: np.ndarray) -> np.ndarray:
    """[n_cells][8][3]: the eight non-zero cell_quadratic_coefficients in the order
    m = 0,1,3,4,9,10,12,13 (poisson_operator.h:165-177)."""
    v = verts
    c = np.empty_like(v)
    c[:, 0] = v[:, 0]
    c[:, 1] = v[:, 1] - v[:, 0]
    c[:, 2] = v[:, 2] - v[:, 0]
    c[:, 3] = v[:, 3] - v[:, 2] - (v[:, 1] - v[:, 0])
    c[:, 4] = v[:, 4] - v[:, 0]
    c[:, 5] = v[:, 5] - v[:, 4] - (v[:, 1] - v[:, 0])
    c[:, 6] = v[:, 6] - v[:, 4] - (v[:, 2] - v[:, 0])
    c[:, 7] = (v[:, 7] - v[:, 6] - (v[:, 5] - v[:, 4]) - (v[:, 3] - v[:, 2] - (v[:, 1] - v[:, 0])))
    return c


def cell_points27(s: int, cells: np.ndarray) -> np.ndarray:
    """[n_cells][27][3] geometry nodes of genuinely quadratic cells: push_forward of the lattice
    points (i, j, k)/2 of every cell, index i + 3j + 9k -- what a MappingQ(2) would place on the
    manifold (the TODO at benchmark.h:75-77)."""
    n_refine, _, _ = mesh_dims(s)
    h = 1.0 / (1 << n_refine)
    off = np.array([[i, j, k] for k in range(3) for j in range(3) for i in range(3)], dtype=np.float64) * 0.5
    lat = (cells[:, None, :].astype(np.float64) + off[None, :, :]) * h
    return push_forward(lat)


def quadratic_coefficients(points27: np.ndarray) -> np.ndarray:
    """[n_cells][27][3] monomial coefficients v_m, m = a + 3b + 9c, of the tri-quadratic
    interpolant X(xi) = sum_m v_m xi^a eta^b zeta^c through the 27 nodes (all entries of
    cell_quadratic_coefficients, poisson_operator.h:690).  1-D: f(t) through t = 0, 1/2, 1 is
    f0 + (-3 f0 + 4 f1 - f2) t + (2 f0 - 4 f1 + 2 f2) t^2."""
    T = np.array([[1.0, 0.0, 0.0], [-3.0, 4.0, -1.0], [2.0, -4.0, 2.0]])     # [power][node]
    pts = points27.reshape(-1, 3, 3, 3, 3)                                     # [n][k][j][i][d]
    c = np.einsum("ck,bj,ai,nkjid->ncbad", T, T, T, pts, optimize=True)
    return c.reshape(-1, 27, 3)


def cell_coefficients(rd) -> np.ndarray:
    """geometry coefficients of a rank's cells: all 27 vectors when the problem was built with
    quadratic cells, else the eight tri-linear ones"""
    if rd.coefficients is not None:
        return rd.coefficients
    return trilinear_coefficients(rd.vertices)


# --------------------------------------------------------------------------- #
# DoF lattice, entity walk, renumbering (renumber_dofs_for_mf.h, strategy 0,1,2)
# --------------------------------------------------------------------------- #


def entity_ranges(p: int, e: int):
    """node offsets inside a cell along one direction for entity code e in {0,1,2}."""
    if e == 0:
        return [0]
    if e == 1:
        return list(range(1, p))
    return [p]


def entity_walk(p: int):
    """Cell-local nodes (i,j,k) in the first-touch visiting order of
    Renumber::cell_assembly (renumber_dofs_for_mf.h:333-357): entities
    a = ex + 3 ey + 9 ez ascending (object table :289-316), nodes lexicographic
    (x fastest) inside each entity -- for a in {10,16} the i1-outer/i0-inner loop over
    `i1 + i0*nn` (:336-347) is exactly the lexicographic walk of the y-face.
    Returns (walk[(p+1)^3][3], entity_of_walk[(p+1)^3], first_walk_pos[27])."""
    walk, ent, first = [], [], []
    for a in range(27):
        ex, ey, ez = a % 3, (a // 3) % 3, a // 9
        first.append(len(walk))
        for k in entity_ranges(p, ez):
            for j in entity_ranges(p, ey):
                for i in entity_ranges(p, ex):
                    walk.append((i, j, k))
                    ent.append(a)
    return np.array(walk, dtype=np.int64).reshape(-1, 3), np.array(ent), np.array(first)


@dataclass
class RankData:
    rank: int
    degree: int
    s: int
    n_cells: int                      # local cells
    cells: np.ndarray                 # [n_cells][3] lattice coords, matrix-free order
    n_owned: int                      # owned DoFs (3 * owned nodes)
    n_ghost: int
    global_offset: int                # first global (renumbered) DoF index of this rank
    node_of_local: np.ndarray         # [n_owned/3 + n_ghost/3] lattice node id per local node
    entity_index: np.ndarray          # [n_cells][27] uint32, first local DoF of entity / INVALID
    vertices: np.ndarray              # [n_cells][8][3]
    constrained: np.ndarray           # local DoF indices of owned Dirichlet DoFs (ascending)
    rhs: np.ndarray                   # [n_owned] b[i] = i % 8, 0 on constrained (benchmark.h:174-176)
    batch_start: np.ndarray           # cell index where each cell batch starts (+ sentinel)
    range_start: np.ndarray           # batch index where each range starts (+ sentinel)
    part_start: np.ndarray            # range index where each partition starts (+ sentinel)
    ghost_owner: np.ndarray           # [n_ghost/3] owning rank of each ghost node
    ghost_remote_local: np.ndarray    # [n_ghost/3] local node index on the owner
    group_sizes: tuple = (0, 0, 0)    # owned DoFs in (single-range, multi/zero-range, multi-rank)
    # cell-batch ranges as cell offsets, and the run of owned DoFs private to each range: the
    # first group of touch_count_grouping (renumber_dofs_for_mf.h:556-590) is ordered by first
    # touch, hence range by range -- the DoF ranges around which MatrixFree::cell_loop runs the
    # pre/post hooks of vmult_with_merged_sums (poisson_operator.h:339-364)
    range_cell_offset: np.ndarray = None
    range_private_offset: np.ndarray = None
    # genuinely quadratic cells (build_problem(quadratic=True)): the 27 geometry nodes and all 27
    # coefficient vectors of the form local_apply evaluates (poisson_operator.h:577-602)
    points27: np.ndarray = None
    coefficients: np.ndarray = None


def _lattice(p, n):
    return [n[d] * p + 1 for d in range(3)]


def fe_q_walk(p: int) -> np.ndarray:
    """Cell-local nodes (i,j,k) in deal.II's FE_Q<3>(p) numbering -- 8 vertices (x + 2y + 4z),
    12 lines (0-3 bottom: x=0, x=1 along y, y=0, y=1 along x; 4-7 the same on top; 8-11 vertical at
    (0,0), (1,0), (0,1), (1,1)), 6 quads (x=0, x=1, y=0, y=1, z=0, z=1; local coordinates (y,z),
    (z,x), (x,y), first one fastest), interior lexicographic -- the order `cf` runs through in
    Renumber::cellbatch_assembly (renumber_dofs_for_mf.h:394-456)."""
    r = range(1, p)
    w = [((v & 1) * p, ((v >> 1) & 1) * p, ((v >> 2) & 1) * p) for v in range(8)]
    for z in (0, p):
        w += [(0, t, z) for t in r] + [(p, t, z) for t in r] + [(t, 0, z) for t in r] + [(t, p, z) for t in r]
    for (x, y) in ((0, 0), (p, 0), (0, p), (p, p)):
        w += [(x, y, t) for t in r]
    for x in (0, p):
        w += [(x, a, b) for b in r for a in r]            # y fastest, z slow
    for y in (0, p):
        w += [(a, y, b) for a in r for b in r]            # z fastest, x slow
    for z in (0, p):
        w += [(a, b, z) for b in r for a in r]            # x fastest, y slow
    w += [(i, j, k) for k in r for j in r for i in r]
    assert len(w) == (p + 1) ** 3 and len(set(w)) == len(w)
    return np.array(w, dtype=np.int64)


def build_problem(degree: int, s: int, n_ranks: int = 1, lanes: int = 8,
                  batches_per_range: int = 1, renumber=(0, 1, 2), quadratic: bool = False) -> list[RankData]:
    """Problem definition of run_templated (benchmark.h:66-176) for `n_ranks` virtual
    MPI ranks: mesh, Q_p^3 DoFs, Dirichlet set, Renumber(0,1,2) numbering,
    LaplaceOperator::initialize data (poisson_operator.h:161-267) and the RHS.

    deal.II behaviours that are not in the reference tree are explicit parameters
    (SURVEY App. B1/B4): `lanes` = VectorizedArray<double>::size(), `batches_per_range`
    = cell batches per cell_partition_data range.  Ranks own equal contiguous chunks
    of the active-cell order (p4est), an interface node belongs to the lowest rank.

    renumber = (assembly, renumber, grouping) strategy triple of Renumber's constructor
    (renumber_dofs_for_mf.h:17-21): assembly 0 cell / 1 cellbatch (:247-361 / :363-459), renumber
    1 first touch / 2 last touch (:461-490; the by-value set copy at :481 makes EVERY touch draw
    a new number, i.e. the order of last touches), grouping 0 base / 1 cellbatch / 2
    cellbatch_range (:537-671).  The benchmark uses (0, 1, 2) (benchmark.h:112).  The entity
    indices are only meaningful for numberings that keep every entity contiguous; for the others
    `contiguous` is False (LaplaceOperator::initialize would throw, poisson_operator.h:198)."""
    p = degree
    assembly_strat, renumber_strat, grouping_strat = renumber
    assert assembly_strat in (0, 1) and renumber_strat in (1, 2) and grouping_strat in (0, 1, 2)
    _, _, n = mesh_dims(s)
    NI, NJ, NK = _lattice(p, n)
    n_nodes = NI * NJ * NK
    cells_all = active_cell_order(s)
    n_cells_all = len(cells_all)
    assert n_cells_all % n_ranks == 0 or n_ranks == 1
    walk, ent_of_walk, first_walk = entity_walk(p)
    npc = (p + 1) ** 3

    def cell_nodes(cells):
        """lattice node ids [n][npc] in entity-walk order"""
        I = cells[:, 0:1] * p + walk[None, :, 0]
        J = cells[:, 1:2] * p + walk[None, :, 1]
        K = cells[:, 2:3] * p + walk[None, :, 2]
        return (K * NJ + J) * NI + I

    # boundary (Dirichlet) nodes, benchmark.h:96-102
    ii = np.arange(n_nodes)
    I, J, K = ii % NI, (ii // NI) % NJ, ii // (NI * NJ)
    on_bnd = (I == 0) | (I == NI - 1) | (J == 0) | (J == NJ - 1) | (K == 0) | (K == NK - 1)
    del ii, I, J, K

    # cell -> rank (contiguous equal chunks), node owner = lowest touching rank
    chunk = n_cells_all // n_ranks
    cell_rank = np.minimum(np.arange(n_cells_all) // chunk, n_ranks - 1)
    owner = np.full(n_nodes, n_ranks, dtype=np.int64)
    multi = np.zeros(n_nodes, dtype=bool)
    if n_ranks > 1:
        nodes_all = cell_nodes(cells_all)
        np.minimum.at(owner, nodes_all.ravel(), np.repeat(cell_rank, npc))
        # node touched by a cell of another rank than its owner -> "multi-rank"
        # (domain_dof_mapping, renumber_dofs_for_mf.h:673-730: owned DoFs on ghost cells)
        other = np.repeat(cell_rank, npc) != owner[nodes_all.ravel()]
        multi[nodes_all.ravel()[other]] = True
        del nodes_all
    else:
        owner[:] = 0

    # ---- per rank: matrix-free cell order, first touch, grouping ------------
    per_rank = []
    new_local = np.full(n_nodes, -1, dtype=np.int64)   # local node index on the owner
    for r in range(n_ranks):
        cells_r = cells_all[cell_rank == r]
        nodes_r = cell_nodes(cells_r)
        if n_ranks > 1:
            comm = (owner[nodes_r] != r).any(axis=1)
            nc_idx = np.nonzero(~comm)[0]
            c_idx = np.nonzero(comm)[0]
            n_before = ((len(nc_idx) // lanes) // 2) * lanes
            parts = [nc_idx[:n_before], c_idx, nc_idx[n_before:]]
        else:
            parts = [np.arange(len(cells_r))]
        order, batch_start, range_start, part_start = [], [0], [0], [0]
        for part in parts:
            nb0 = len(batch_start) - 1
            for b0 in range(0, len(part), lanes):
                order.append(part[b0:b0 + lanes])
                batch_start.append(batch_start[-1] + len(order[-1]))
            nb1 = len(batch_start) - 1
            for rb in range(nb0, nb1, batches_per_range):
                range_start.append(min(rb + batches_per_range, nb1))
            part_start.append(len(range_start) - 1)
        order = np.concatenate(order) if order else np.zeros(0, dtype=np.int64)
        cells_r = cells_r[order]
        nodes_r = nodes_r[order]
        batch_start = np.array(batch_start)
        range_start = np.array(range_start)
        part_start = np.array(part_start)

        # traversal of the numbering pass: cell by cell in entity-walk order (cell_assembly,
        # :320-358), or batch by batch, FE_Q slot by slot, lane by lane (cellbatch_assembly, :378-456)
        flat = nodes_r.ravel()
        if assembly_strat == 1:
            fw = fe_q_walk(p)
            fe_nodes = ((cells_r[:, 2:3] * p + fw[None, :, 2]) * NJ + cells_r[:, 1:2] * p + fw[None, :, 1]) * NI \
                + cells_r[:, 0:1] * p + fw[None, :, 0]
            trav = np.concatenate([fe_nodes[batch_start[b]:batch_start[b + 1]].T.ravel()
                                   for b in range(len(batch_start) - 1)]) if len(cells_r) else flat
        else:
            trav = flat
        # first touch (first_touch_renumber, :461-474) / last touch (:476-490) key per owned node
        if renumber_strat == 1:
            uniq_t, pos_t = np.unique(trav, return_index=True)
        else:
            uniq_t, pos_r = np.unique(trav[::-1], return_index=True)
            pos_t = len(trav) - 1 - pos_r
        uniq, first_pos = np.unique(flat, return_index=True)
        assert np.array_equal(uniq, uniq_t)
        owned_mask = owner[uniq] == r
        owned_nodes = uniq[owned_mask]
        ft = pos_t[owned_mask]                           # ordering key (unique)
        first_cell = first_pos[owned_mask] // npc        # first cell touching the node (loop order)
        # touch count over cell-batch ranges (touch_count_cellbatch_range, :622-671);
        # constrained DoFs are not in MatrixFree's index lists -> count 0
        cell_batch = np.searchsorted(batch_start[1:], np.arange(len(cells_r)), side="right")
        cell_range = np.searchsorted(range_start[1:], cell_batch, side="right")
        # grouping 1 counts cell batches (touch_count_cellbatch, :592-620), 2 counts ranges
        cell_group = cell_batch if grouping_strat == 1 else cell_range
        n_grp = len(batch_start) if grouping_strat == 1 else len(range_start)
        pair = np.unique(flat.astype(np.int64) * n_grp + np.repeat(cell_group, npc))
        cnt_nodes, cnt = np.unique(pair // n_grp, return_counts=True)
        tc = np.zeros(n_nodes, dtype=np.int64)
        tc[cnt_nodes] = cnt
        tc[on_bnd] = 0
        tco = tc[owned_nodes]
        mo = multi[owned_nodes]
        if grouping_strat == 0:                          # base_grouping, :537-554
            g1 = ~mo
            g2 = np.zeros_like(mo)
        else:
            g1 = (~mo) & (tco == 1)
            g2 = (~mo) & (tco != 1)
        g3 = mo
        seq = []
        for g in (g1, g2, g3):                           # grouping, :492-535, :556-590
            idx = np.nonzero(g)[0]
            seq.append(idx[np.argsort(ft[idx], kind="stable")])
        # group-1 nodes per range: the (only) range touching such a node is that of its first touch
        n_rng = len(range_start) - 1
        # (only the cellbatch_range grouping of cell-wise numberings makes these runs contiguous)
        if renumber == (0, 1, 2) or renumber == (0, 2, 2):
            priv_cnt = np.bincount(cell_range[first_cell[g1]], minlength=n_rng) if n_rng else np.zeros(0, dtype=np.int64)
            range_private_offset = 3 * np.concatenate(([0], np.cumsum(priv_cnt)))
            range_cell_offset = batch_start[range_start]
        else:
            range_private_offset = range_cell_offset = np.zeros(0, dtype=np.int64)
        seq = np.concatenate(seq)
        new_nodes = owned_nodes[seq]                     # lattice node at each new local position
        new_local[new_nodes] = np.arange(len(new_nodes))
        per_rank.append(dict(cells=cells_r, nodes=nodes_r, owned=new_nodes,
                             batch_start=batch_start, range_start=range_start,
                             part_start=part_start, uniq=uniq,
                             range_cell_offset=range_cell_offset.astype(np.uint64),
                             range_private_offset=range_private_offset.astype(np.uint64),
                             groups=(3 * int(g1.sum()), 3 * int(g2.sum()), 3 * int(g3.sum()))))

    node_offset = np.concatenate(([0], np.cumsum([len(d["owned"]) for d in per_rank])))
    global_new = node_offset[np.minimum(owner, n_ranks - 1)] + new_local

    out = []
    for r, d in enumerate(per_rank):
        uniq = d["uniq"]
        ghosts = uniq[owner[uniq] != r]
        ghosts = ghosts[np.argsort(global_new[ghosts], kind="stable")]   # sorted by global index
        n_on = len(d["owned"])
        local_of_node = np.full(n_nodes, -1, dtype=np.int64)
        local_of_node[d["owned"]] = np.arange(n_on)
        local_of_node[ghosts] = n_on + np.arange(len(ghosts))
        first_nodes = d["nodes"][:, first_walk]                            # [n_cells][27]
        ei = (3 * local_of_node[first_nodes]).astype(np.int64)
        ei[on_bnd[first_nodes]] = int(INVALID)
        verts = cell_vertices(s, d["cells"])
        con_nodes = np.nonzero(on_bnd[d["owned"]])[0]
        constrained = (3 * con_nodes[:, None] + np.arange(3)[None, :]).ravel()
        rhs = (np.arange(3 * n_on) % 8).astype(np.float64)
        rhs[constrained] = 0.0
        out.append(RankData(
            rank=r, degree=p, s=s, n_cells=len(d["cells"]), cells=d["cells"],
            n_owned=3 * n_on, n_ghost=3 * len(ghosts), global_offset=3 * int(node_offset[r]),
            node_of_local=np.concatenate([d["owned"], ghosts]),
            entity_index=ei.astype(np.uint32), vertices=verts,
            constrained=constrained.astype(np.uint32), rhs=rhs,
            batch_start=d["batch_start"], range_start=d["range_start"], part_start=d["part_start"],
            ghost_owner=owner[ghosts], ghost_remote_local=new_local[ghosts],
            group_sizes=d["groups"], range_cell_offset=d["range_cell_offset"],
            range_private_offset=d["range_private_offset"]))
        if quadratic:
            out[-1].points27 = cell_points27(s, d["cells"])
            out[-1].coefficients = quadratic_coefficients(out[-1].points27)
    return out


def n_dofs_total(degree: int, s: int) -> int:
    _, _, n = mesh_dims(s)
    return 3 * int(np.prod(_lattice(degree, n)))


# --------------------------------------------------------------------------- #
# operator (poisson_operator.h:429-685, vector_access_reduced.h:175-283, :437-531)
# --------------------------------------------------------------------------- #


def local_dof_map(p: int, entity_index: np.ndarray) -> np.ndarray:
    """[n_cells][(p+1)^3] local index of component 0 of every cell node in
    lexicographic (x fastest) order, or -1 where the entity is constrained; node
    (i,j,k) of entity a sits at idx[a] + 3*(lexicographic position in a)
    (vector_access_reduced.h:176-258)."""
    n1 = p + 1
    ent = np.zeros((n1, n1, n1), dtype=np.int64)
    pos = np.zeros((n1, n1, n1), dtype=np.int64)

    def code(i):
        return 0 if i == 0 else (2 if i == p else 1)

    def off(i):
        return i - 1 if 0 < i < p else 0

    def size(e):
        return p - 1 if e == 1 else 1
    for k in range(n1):
        for j in range(n1):
            for i in range(n1):
                ex, ey, ez = code(i), code(j), code(k)
                ent[k, j, i] = ex + 3 * ey + 9 * ez
                pos[k, j, i] = off(i) + size(ex) * (off(j) + size(ey) * off(k))
    ent, pos = ent.ravel(), pos.ravel()
    base = entity_index.astype(np.int64)[:, ent]
    m = base + 3 * pos[None, :]
    m[base == int(INVALID)] = -1
    return m


def do_invert(J: np.ndarray):
    """3x3 inverse through cofactors, returns (inverse, det) -- poisson_operator.h:41-63."""
    t = J
    tr00 = t[..., 1, 1] * t[..., 2, 2] - t[..., 1, 2] * t[..., 2, 1]
    tr10 = t[..., 1, 2] * t[..., 2, 0] - t[..., 1, 0] * t[..., 2, 2]
    tr20 = t[..., 1, 0] * t[..., 2, 1] - t[..., 1, 1] * t[..., 2, 0]
    det = t[..., 0, 0] * tr00 + t[..., 0, 1] * tr10 + t[..., 0, 2] * tr20
    inv_det = 1.0 / det
    out = np.empty_like(t)
    out[..., 0, 0] = inv_det * tr00
    out[..., 0, 1] = inv_det * (t[..., 0, 2] * t[..., 2, 1] - t[..., 0, 1] * t[..., 2, 2])
    out[..., 0, 2] = inv_det * (t[..., 0, 1] * t[..., 1, 2] - t[..., 0, 2] * t[..., 1, 1])
    out[..., 1, 0] = inv_det * tr10
    out[..., 1, 1] = inv_det * (t[..., 0, 0] * t[..., 2, 2] - t[..., 0, 2] * t[..., 2, 0])
    out[..., 1, 2] = inv_det * (t[..., 0, 2] * t[..., 1, 0] - t[..., 0, 0] * t[..., 1, 2])
    out[..., 2, 0] = inv_det * tr20
    out[..., 2, 1] = inv_det * (t[..., 0, 1] * t[..., 2, 0] - t[..., 0, 0] * t[..., 2, 1])
    out[..., 2, 2] = inv_det * (t[..., 0, 0] * t[..., 1, 1] - t[..., 0, 1] * t[..., 1, 0])
    return out, det


def jacobians(coef: np.ndarray, x1d: np.ndarray) -> np.ndarray:
    """jac[cell][qz][qy][qx][row][d], rows = dX/dxi, dX/deta, dX/dzeta evaluated from
    the tri-linear coefficients exactly as poisson_operator.h:577-602 (the 19
    quadratic coefficients are identically zero, :161-178)."""
    if coef.shape[1] == 27:
        # all 27 coefficients: d/dxi_e of sum_m v_m x^a y^b z^c (poisson_operator.h:577-602)
        P = np.stack([np.ones_like(x1d), x1d, x1d * x1d])                     # [power][q]
        dP = np.stack([np.zeros_like(x1d), np.ones_like(x1d), 2.0 * x1d])
        v = coef.reshape(-1, 3, 3, 3, 3)                                       # [n][c][b][a][d]
        nq = len(x1d)
        jac = np.empty((coef.shape[0], nq, nq, nq, 3, 3))
        jac[..., 0, :] = np.einsum("ncbad,ax,by,cz->nzyxd", v, dP, P, P, optimize=True)
        jac[..., 1, :] = np.einsum("ncbad,ax,by,cz->nzyxd", v, P, dP, P, optimize=True)
        jac[..., 2, :] = np.einsum("ncbad,ax,by,cz->nzyxd", v, P, P, dP, optimize=True)
        return jac
    v1, v3, v4, v9, v10, v12, v13 = (coef[:, i] for i in (1, 2, 3, 4, 5, 6, 7))
    x = x1d[None, None, None, :, None]
    y = x1d[None, None, :, None, None]
    z = x1d[None, :, None, None, None]

    def b(a):
        return a[:, None, None, None, :]
    nq = len(x1d)
    jac = np.empty((coef.shape[0], nq, nq, nq, 3, 3))
    jac[..., 0, :] = (b(v1) + z * b(v10)) + y * (b(v4) + z * b(v13)) + 0 * x
    jac[..., 1, :] = (b(v3) + z * b(v12)) + x * (b(v4) + z * b(v13)) + 0 * y
    jac[..., 2, :] = (b(v9) + y * b(v12)) + x * (b(v10) + y * b(v13)) + 0 * z
    return jac


def apply_cells(t: Tables, coef: np.ndarray, u: np.ndarray) -> np.ndarray:
    """Per-cell kernel of LaplaceOperator::local_apply (3-D branch,
    poisson_operator.h:534-666) for cell-local values u[cell][comp][k][j][i]
    -> result in the same layout.  SURVEY App. A7."""
    S, D = t.S, t.D
    uq = np.einsum("kc,jb,ia,nmkji->nmcba", S, S, S, u, optimize=True)          # values at q points
    gx = np.einsum("ax,nmcba->nmcbx", D, uq, optimize=True)                     # d/dxi
    gy = np.einsum("by,nmcba->nmcya", D, uq, optimize=True)
    gz = np.einsum("cz,nmcba->nmzba", D, uq, optimize=True)
    g = np.stack([gx, gy, gz], axis=-1)                                         # [n][m][z][y][x][e]
    jinv, det = do_invert(jacobians(coef, t.xq))                                # jinv = J^{-T}
    w = t.wq[:, None, None] * t.wq[None, :, None] * t.wq[None, None, :]
    det = det * w[None]
    tmp = np.einsum("nzyxde,nmzyxe->nmzyxd", jinv, g, optimize=True) * det[:, None, ..., None]
    out = np.einsum("nzyxed,nmzyxe->nmzyxd", jinv, tmp, optimize=True)
    r = np.einsum("ax,nmcbx->nmcba", D, out[..., 0], optimize=True)
    r += np.einsum("by,nmcya->nmcba", D, out[..., 1], optimize=True)
    r += np.einsum("cz,nmzba->nmcba", D, out[..., 2], optimize=True)
    return np.einsum("kc,jb,ia,nmcba->nmkji", S, S, S, r, optimize=True)


def vmult_cells(rd: RankData, t: Tables, src: np.ndarray, chunk: int = 4096) -> np.ndarray:
    """cell-loop part of vmult (gather, apply, scatter-add) on a local vector of
    length n_owned+n_ghost; constrained entries read as 0 and are not written."""
    p = rd.degree
    n1 = p + 1
    dmap = local_dof_map(p, rd.entity_index)
    coef = cell_coefficients(rd)
    dst = np.zeros(rd.n_owned + rd.n_ghost)
    for c0 in range(0, rd.n_cells, chunk):
        m = dmap[c0:c0 + chunk]
        valid = m >= 0
        idx = np.where(valid, m, 0)
        u = np.stack([np.where(valid, src[idx + c], 0.0) for c in range(3)], axis=1)
        u = u.reshape(-1, 3, n1, n1, n1)
        r = apply_cells(t, coef[c0:c0 + chunk], u).reshape(-1, 3, n1 ** 3)
        for c in range(3):
            np.add.at(dst, (idx + c)[valid], r[:, c][valid])
    return dst


def vmult(rd: RankData, t: Tables, src: np.ndarray) -> np.ndarray:
    """LaplaceOperator::vmult on one rank (poisson_operator.h:307-313): cell loop,
    then dst = src on constrained rows."""
    assert rd.n_ghost == 0
    dst = vmult_cells(rd, t, src)
    dst[rd.constrained] = src[rd.constrained]
    return dst


def inverse_diagonal(rd: RankData) -> np.ndarray:
    """compute_inverse_diagonal + extraction (poisson_operator.h:392-426,
    benchmark.h:141-147): diagonal of the scalar Laplacian with GLL(p+1) quadrature,
    assembled over cells skipping constrained DoFs, 1/d (1 where d == 0), one entry
    per node.  Returns [(n_owned+n_ghost)/3]; only owned entries are meaningful on a
    multi-rank partition until contributions are exchanged (see virtual ranks)."""
    p = rd.degree
    t = make_tables(p, p + 1, quad="gll")
    coef = cell_coefficients(rd)
    jinv, det = do_invert(jacobians(coef, t.xq))
    w = t.wq[:, None, None] * t.wq[None, :, None] * t.wq[None, None, :]
    G = np.einsum("nzyxde,nzyxdf->nzyxef", jinv, jinv) * (det * w[None])[..., None, None]
    # unit vector at node (k,j,i): gradient non-zero along the three grid lines
    # through it.  S = identity (collocation); D[i][q] = l_i'(x_q).
    D = t.D
    n1 = p + 1
    d = np.einsum("iq,nkjq->nkji", D * D, G[..., 0, 0])
    d += np.einsum("jq,nkqi->nkji", D * D, G[..., 1, 1])
    d += np.einsum("kq,nqji->nkji", D * D, G[..., 2, 2])
    # cross terms only where both derivative lines pass through the same point = the node
    dd = np.diag(D)
    d += 2 * dd[None, None, None, :] * dd[None, None, :, None] * G[..., 0, 1]
    d += 2 * dd[None, None, None, :] * dd[None, :, None, None] * G[..., 0, 2]
    d += 2 * dd[None, None, :, None] * dd[None, :, None, None] * G[..., 1, 2]
    dmap = local_dof_map(p, rd.entity_index)
    valid = dmap >= 0
    diag = np.zeros((rd.n_owned + rd.n_ghost) // 3)
    np.add.at(diag, (dmap // 3)[valid], d.reshape(-1, n1 ** 3)[valid])
    return diag


def finish_inverse_diagonal(diag: np.ndarray) -> np.ndarray:
    """0 -> 1 else 1/x, poisson_operator.h:420-424."""
    out = np.ones_like(diag)
    nz = diag != 0.0
    out[nz] = 1.0 / diag[nz]
    return out


# --------------------------------------------------------------------------- #
# solvers
# --------------------------------------------------------------------------- #


class ReductionControl:
    """deal.II ReductionControl(max_steps, tol, reduce) as used at bench.cc:11
    (SURVEY App. B3)."""

    def __init__(self, max_steps=100, tol=1e-15, reduce=1e-8):
        self.max_steps, self.tol, self.reduce = max_steps, tol, reduce
        self.last_step, self.last_value, self.initial = 0, 0.0, 0.0
        self.history = []

    def check(self, step, value):
        if step == 0:
            self.initial = value
            self.reduced_tol = value * self.reduce
        self.last_step, self.last_value = step, value
        self.history.append(value)
        if value < self.reduced_tol or value <= self.tol:      # strict for the reduced tolerance (deal.II)
            return "success"
        if step >= self.max_steps or math.isnan(value):
            return "failure"
        return "iterate"


def jacobi_vmult(diag: np.ndarray, src: np.ndarray) -> np.ndarray:
    """DiagonalMatrixBlocked::vmult, diagonal_matrix_blocked.h:13-27."""
    return np.repeat(diag, 3)[: len(src)] * src


def solver_cg_plain(A, x, b, diag, control: ReductionControl, dot=np.dot):
    """deal.II 9.3 SolverCG::solve with a preconditioner, as instantiated by
    benchmark_precond/bench.cc:11-16.  `A` maps a vector to A*v."""
    g = -b.copy() if not x.any() else A(x) - b
    res = math.sqrt(dot(g, g))
    if control.check(0, res) != "iterate":
        return x
    h = jacobi_vmult(diag, g)
    d = -h
    gh = dot(g, h)
    it = 0
    while True:
        it += 1
        h = A(d)
        alpha = gh / dot(d, h)
        x += alpha * d
        g += alpha * h
        res = math.sqrt(abs(dot(g, g)))
        if control.check(it, res) != "iterate":
            break
        h = jacobi_vmult(diag, g)
        beta_den = gh
        gh = dot(g, h)
        beta = gh / beta_den
        d = beta * d - h
    return x


def cg_update4b(h, x, r, p, prec3, alpha, beta, alpha_old, beta_old):
    """do_cg_update4b<3,double,true>, solver_cg_optimized.h:65-161 (whole range)."""
    if alpha == 0.0:
        p[:] = -prec3 * r
    elif alpha_old == 0.0:
        r += alpha * h
        p[:] = beta * p - prec3 * r
    else:
        x += (alpha + alpha_old / beta_old) * p + (alpha_old / beta_old) * prec3 * r
        r += alpha * h
        p[:] = beta * p - prec3 * r
    h[:] = 0.0


def cg_update3b(r, d, h, prec3, dot=np.dot):
    """do_cg_update3b, solver_cg_optimized.h:12-61: the seven merged sums."""
    zi = prec3 * h
    return np.array([dot(d, h), dot(h, h), dot(r, h), dot(r, r), dot(r, zi), dot(h, zi),
                     dot(r, prec3 * r)])


def solver_cg_merged(A_cells, x, b, diag, control: ReductionControl, reduce=lambda s: s):
    """SolverCGFullMerge::solve (solver_cg_optimized.h:192-302) around
    vmult_with_merged_sums (poisson_operator.h:327-377).  `A_cells(d)` is the
    cell-loop part of the operator (no constrained-row fix-up)."""
    n = len(b)
    prec3 = np.repeat(diag, 3)[:n]
    g = -b.copy()
    assert not x.any()
    d = np.zeros(n)
    h = np.zeros(n)
    res = math.sqrt(reduce(np.dot(g, g)))
    if control.check(0, res) != "iterate":
        return x
    alpha = beta = alpha_old = beta_old = 0.0
    it = 0
    while True:
        it += 1
        cg_update4b(h, x, g, d, prec3, alpha, beta, alpha_old if it % 2 == 1 else 0.0, beta_old)
        h[:] = A_cells(d)[:n]
        S = reduce(cg_update3b(g, d, h, prec3))
        alpha_old, beta_old = alpha, beta
        alpha = S[6] / S[0]
        res = math.sqrt(S[3] + 2 * alpha * S[2] + alpha * alpha * S[1])
        if control.check(it, res) != "iterate":
            if it % 2 == 1:
                x += alpha * d
            else:
                x += (alpha + alpha_old / beta_old) * d + (alpha_old / beta_old) * prec3 * g
            break
        beta = alpha * (S[4] + alpha * S[5]) / S[6]
    return x


# --------------------------------------------------------------------------- #
# virtual MPI ranks in one process: ghost exchange emulated with numpy
# --------------------------------------------------------------------------- #


def multi_update_ghosts(rds, vecs):
    """LA::distributed::Vector::update_ghost_values: owners -> ghost copies"""
    for rd, v in zip(rds, vecs):
        for g in range(rd.n_ghost // 3):
            o, l = int(rd.ghost_owner[g]), int(rd.ghost_remote_local[g])
            v[rd.n_owned + 3 * g: rd.n_owned + 3 * g + 3] = vecs[o][3 * l: 3 * l + 3]


def multi_compress_add(rds, vecs):
    """compress(VectorOperation::add): ghost contributions added into the owners, ghosts zeroed"""
    for rd, v in zip(rds, vecs):
        for g in range(rd.n_ghost // 3):
            o, l = int(rd.ghost_owner[g]), int(rd.ghost_remote_local[g])
            vecs[o][3 * l: 3 * l + 3] += v[rd.n_owned + 3 * g: rd.n_owned + 3 * g + 3]
        v[rd.n_owned:] = 0.0


def multi_vmult_cells(rds, t, srcs, cell_op=None):
    """cell loop of every rank with the two exchanges of MatrixFree::cell_loop around it;
    srcs/dsts are local vectors [owned | ghost]"""
    srcs = [s.copy() for s in srcs]
    multi_update_ghosts(rds, srcs)
    dsts = [(cell_op(rd, s) if cell_op else vmult_cells(rd, t, s)) for rd, s in zip(rds, srcs)]
    multi_compress_add(rds, dsts)
    return dsts


def multi_vmult(rds, t, srcs, cell_op=None):
    dsts = multi_vmult_cells(rds, t, srcs, cell_op)
    for rd, d, s in zip(rds, dsts, srcs):
        d[rd.constrained] = s[rd.constrained]
    return dsts


def multi_inverse_diagonal(rds):
    diags = []
    for rd in rds:
        d = np.zeros(rd.n_owned + rd.n_ghost)
        d[0::3] = inverse_diagonal(rd)
        diags.append(d)
    multi_compress_add(rds, diags)
    return [finish_inverse_diagonal(d[: rd.n_owned: 3]) for rd, d in zip(rds, diags)]


def multi_cg_merged(rds, t, bs, diags, control: ReductionControl, cell_op=None):
    """SolverCGFullMerge over virtual ranks: per-rank pre/post sweeps, exchanged cell loop,
    sums reduced over the ranks (the MPI_Allreduce at poisson_operator.h:373)"""
    n = [rd.n_owned for rd in rds]
    prec3 = [np.repeat(d, 3)[:k] for d, k in zip(diags, n)]
    xs = [np.zeros(rd.n_owned + rd.n_ghost) for rd in rds]
    gs = [np.concatenate([-b[:k], np.zeros(rd.n_ghost)]) for b, k, rd in zip(bs, n, rds)]
    ds = [np.zeros(rd.n_owned + rd.n_ghost) for rd in rds]
    hs = [np.zeros(rd.n_owned + rd.n_ghost) for rd in rds]
    res = math.sqrt(sum(float(g[:k] @ g[:k]) for g, k in zip(gs, n)))
    if control.check(0, res) != "iterate":
        return xs
    alpha = beta = alpha_old = beta_old = 0.0
    it = 0
    while True:
        it += 1
        for r in range(len(rds)):
            k = n[r]
            cg_update4b(hs[r][:k], xs[r][:k], gs[r][:k], ds[r][:k], prec3[r], alpha, beta,
                        alpha_old if it % 2 == 1 else 0.0, beta_old)
        hs = multi_vmult_cells(rds, t, ds, cell_op)
        S = sum(cg_update3b(gs[r][:n[r]], ds[r][:n[r]], hs[r][:n[r]], prec3[r]) for r in range(len(rds)))
        alpha_old, beta_old = alpha, beta
        alpha = S[6] / S[0]
        res = math.sqrt(S[3] + 2 * alpha * S[2] + alpha * alpha * S[1])
        if control.check(it, res) != "iterate":
            for r in range(len(rds)):
                k = n[r]
                if it % 2 == 1:
                    xs[r][:k] += alpha * ds[r][:k]
                else:
                    xs[r][:k] += (alpha + alpha_old / beta_old) * ds[r][:k] + (alpha_old / beta_old) * prec3[r] * gs[r][:k]
            break
        beta = alpha * (S[4] + alpha * S[5]) / S[6]
    return xs


# --------------------------------------------------------------------------- #
# independent pin: dense brute-force assembly (no sum factorisation, no entity
# compression) of the same bilinear form on a handful of cells
# --------------------------------------------------------------------------- #


def dense_matrix(rd: RankData, t: Tables) -> np.ndarray:
    """A[i][j] = sum_cells sum_q grad(phi_i).grad(phi_j) det(J) w_q, vector-valued
    (block diagonal over components), Dirichlet rows/cols replaced by identity.
    Built from full 3-D Lagrange gradients at every quadrature point and a
    numpy.linalg inverse of the Jacobian -- shares no code path with apply_cells."""
    p = rd.degree
    n1, nq = p + 1, t.n_q
    n = rd.n_owned + rd.n_ghost
    A = np.zeros((n, n))
    V = lagrange_values(t.xn, t.xq)         # [i][q]
    G = lagrange_derivs(t.xn, t.xq)         # d/dx of the nodal basis at the q points
    dmap = local_dof_map(p, rd.entity_index)
    for c in range(rd.n_cells):
        X = rd.vertices[c]
        X27 = rd.points27[c] if rd.points27 is not None else None
        # reference gradients of all (p+1)^3 basis functions at all q^3 points
        gr = np.zeros((n1, n1, n1, nq, nq, nq, 3))
        for k in range(n1):
            for j in range(n1):
                for i in range(n1):
                    gr[k, j, i, :, :, :, 0] = V[k][:, None, None] * V[j][None, :, None] * G[i][None, None, :]
                    gr[k, j, i, :, :, :, 1] = V[k][:, None, None] * G[j][None, :, None] * V[i][None, None, :]
                    gr[k, j, i, :, :, :, 2] = G[k][:, None, None] * V[j][None, :, None] * V[i][None, None, :]
        gr = gr.reshape(n1 ** 3, nq ** 3, 3)
        Ke = np.zeros((n1 ** 3, n1 ** 3))
        q = 0
        for qz in range(nq):
            for qy in range(nq):
                for qx in range(nq):
                    xi = np.array([t.xq[qx], t.xq[qy], t.xq[qz]])
                    # dX/dxi_e from the trilinear vertex interpolation
                    J = np.zeros((3, 3))      # J[d][e] = dX_d / dxi_e
                    if X27 is not None:
                        # tri-quadratic Lagrange interpolation of the 27 geometry nodes (nodes 0, 1/2, 1)
                        nodes = np.array([0.0, 0.5, 1.0])
                        L = [lagrange_values(nodes, np.array([xi[e]]))[:, 0] for e in range(3)]
                        dL = [lagrange_derivs(nodes, np.array([xi[e]]))[:, 0] for e in range(3)]
                        for k2 in range(3):
                            for j2 in range(3):
                                for i2 in range(3):
                                    Xn = X27[i2 + 3 * j2 + 9 * k2]
                                    J[:, 0] += Xn * dL[0][i2] * L[1][j2] * L[2][k2]
                                    J[:, 1] += Xn * L[0][i2] * dL[1][j2] * L[2][k2]
                                    J[:, 2] += Xn * L[0][i2] * L[1][j2] * dL[2][k2]
                    for v in (range(8) if X27 is None else ()):
                        b = [(v >> e) & 1 for e in range(3)]
                        for e in range(3):
                            f = 1.0
                            for e2 in range(3):
                                if e2 == e:
                                    f *= 1.0 if b[e2] else -1.0
                                else:
                                    f *= xi[e2] if b[e2] else 1.0 - xi[e2]
                            J[:, e] += X[v] * f
                    Ji = np.linalg.inv(J)
                    wdet = np.linalg.det(J) * t.wq[qx] * t.wq[qy] * t.wq[qz]
                    pg = gr[:, q, :] @ Ji                  # physical gradients [node][d]
                    Ke += wdet * (pg @ pg.T)
                    q += 1
        m = dmap[c]
        ok = m >= 0
        for comp in range(3):
            rows = m[ok] + comp
            A[np.ix_(rows, rows)] += Ke[np.ix_(ok, ok)]
    A[rd.constrained, rd.constrained] = 1.0
    return A
