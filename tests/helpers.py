"""Shared helpers of the parity tests: problem cache and a thin driver that runs the
reference's solver recurrences through the C ABI (used until/alongside the C++ host)."""
from __future__ import annotations

import functools
import math

import numpy as np

from oracle import bp4_oracle as O
from oracle.c_oracle import COracle


@functools.lru_cache(maxsize=32)
def problem(p, s, n_ranks=1):
    rds = O.build_problem(p, s, n_ranks=n_ranks)
    return rds


@functools.lru_cache(maxsize=32)
def single(p, s):
    rd = problem(p, s)[0]
    return rd, COracle(rd)


def make_ctx(rd, device=0, ranges=True):
    """ranges=True hands the oracle's cell-batch ranges + private DoF runs to the context, which
    switches the in-loop (fused) vector updates of vmult_with_merged_sums on"""
    from mf_data_locality_b200 import capi
    return capi.Context(rd.degree, rd.entity_index, rd.vertices, rd.n_owned, rd.n_ghost,
                        rd.constrained, device=device,
                        ranges=(rd.range_cell_offset, rd.range_private_offset) if ranges else None,
                        coefficients=rd.coefficients)


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def gpu_cg_merged(ctx, b_host, prec, control: O.ReductionControl):
    """SolverCGFullMerge::solve (solver_cg_optimized.h:192-302) driven through the C ABI."""
    x = ctx.vector()
    g = ctx.vector()
    d = ctx.vector()
    h = ctx.vector()
    b = ctx.vector(data=b_host)
    ctx.equ(g, -1.0, b)
    res = ctx.l2_norm(g)
    if control.check(0, res) != "iterate":
        return x
    alpha = beta = alpha_old = beta_old = 0.0
    it = 0
    while True:
        it += 1
        S = ctx.vmult_merged(x, g, d, h, prec, alpha, beta, alpha_old if it % 2 == 1 else 0.0, beta_old)
        alpha_old, beta_old = alpha, beta
        alpha = S[6] / S[0]
        res = math.sqrt(S[3] + 2 * alpha * S[2] + alpha * alpha * S[1])
        if control.check(it, res) != "iterate":
            if it % 2 == 1:
                ctx.add(x, alpha, d)
            else:
                ctx.x_finalize_even(x, d, g, prec, alpha + alpha_old / beta_old, alpha_old / beta_old)
            break
        beta = alpha * (S[4] + alpha * S[5]) / S[6]
    return x


def gpu_cg_plain(ctx, b_host, prec, control: O.ReductionControl):
    """deal.II SolverCG::solve as used by benchmark_precond/bench.cc:11-16, through the C ABI."""
    x = ctx.vector()
    g = ctx.vector()
    d = ctx.vector()
    h = ctx.vector()
    b = ctx.vector(data=b_host)
    ctx.equ(g, -1.0, b)
    res = ctx.l2_norm(g)
    if control.check(0, res) != "iterate":
        return x
    ctx.jacobi_vmult(h, g, prec)
    ctx.equ(d, -1.0, h)
    gh = ctx.dot(g, h)
    it = 0
    while True:
        it += 1
        ctx.vmult(h, d)
        alpha = gh / ctx.dot(d, h)
        ctx.add(x, alpha, d)
        res = math.sqrt(abs(ctx.add_and_dot(g, alpha, h, g)))
        if control.check(it, res) != "iterate":
            break
        ctx.jacobi_vmult(h, g, prec)
        beta_den = gh
        gh = ctx.dot(g, h)
        ctx.sadd(d, gh / beta_den, -1.0, h)
    return x
