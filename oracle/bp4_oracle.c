/* CPU ORACLE (test infrastructure, NOT product code) -- C/OpenMP restatement of the
 * BP4 hot path of peterrum/mf_data_locality for sizes the numpy oracle is too slow
 * for, and the `cpu_baseline` / `--impl reference` legs of bench.py.
 * PARITY UNPINNED by the reference (it ships no tests/golden vectors and cannot be
 * built here: deal.II + p4est + MPI are absent); pinned instead against the numpy
 * oracle, which is pinned against dense brute-force assembly (tests/).
 * Only tests/, __graft_entry__.smoke() and bench.py may load this library.
 *
 * Setup data (entity indices, tri-linear coefficients, 1-D tables) comes from
 * oracle/bp4_oracle.py.  Citations are relative to /root/reference/common_code/. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#  include <omp.h>
#endif

typedef struct
{
  int           degree;
  const double *S;   /* [N][Q] */
  const double *D;   /* [Q][Q] */
  const double *xq;  /* [Q] */
  const double *wq;  /* [Q] */
  const int    *ent; /* [N^3] entity of every cell node (lexicographic) */
  const int    *pos; /* [N^3] lexicographic position inside its entity */
  /* optional colouring: cells of one colour share no DoF, so they can be scattered without
   * atomics (plays the role of the rank-private data of the reference's MPI run) */
  int         n_colors;
  const long *color_start; /* [n_colors + 1] */
  const long *color_cells; /* cell indices grouped by colour */
} oracle_tables;

#define VL 8
typedef double vd __attribute__((vector_size(8 * VL)));

#define P 2
#include "bp4_oracle_kernel.inc"
#include "bp4_oracle_fast.inc"
#undef P
#define P 3
#include "bp4_oracle_kernel.inc"
#include "bp4_oracle_fast.inc"
#undef P
#define P 4
#include "bp4_oracle_kernel.inc"
#include "bp4_oracle_fast.inc"
#undef P
#define P 5
#include "bp4_oracle_kernel.inc"
#include "bp4_oracle_fast.inc"
#undef P
#define P 6
#include "bp4_oracle_kernel.inc"
#include "bp4_oracle_fast.inc"
#undef P
#define P 7
#include "bp4_oracle_kernel.inc"
#include "bp4_oracle_fast.inc"
#undef P
#define P 8
#include "bp4_oracle_kernel.inc"
#include "bp4_oracle_fast.inc"
#undef P

typedef void (*cell_fn)(const oracle_tables *, const long *, int, const uint32_t *, const double *,
                        const double *, double *, int);
/* kernel used by every entry point below: 1 = even-odd, layer-wise (bp4_oracle_fast.inc, default),
 * 0 = plain dense restatement (bp4_oracle_kernel.inc); the two are tested against each other */
static int g_fast = 1;
void oracle_set_fast(int on) { g_fast = on; }

void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
  if (n > 0)
    omp_set_num_threads(n);
#else
  (void)n;
#endif
}

static cell_fn pick(int p)
{
  if (g_fast)
    switch (p)
      {
        case 2: return apply_cell_fast_2;
        case 3: return apply_cell_fast_3;
        case 4: return apply_cell_fast_4;
        case 5: return apply_cell_fast_5;
        case 6: return apply_cell_fast_6;
        case 7: return apply_cell_fast_7;
        case 8: return apply_cell_fast_8;
      }
  switch (p)
    {
      case 2: return apply_cell_2;
      case 3: return apply_cell_3;
      case 4: return apply_cell_4;
      case 5: return apply_cell_5;
      case 6: return apply_cell_6;
      case 7: return apply_cell_7;
      case 8: return apply_cell_8;
    }
  return 0;
}

static void prepare(const oracle_tables *t)
{
  switch (t->degree)
    {
      case 2: eo_prepare_2(t); break;
      case 3: eo_prepare_3(t); break;
      case 4: eo_prepare_4(t); break;
      case 5: eo_prepare_5(t); break;
      case 6: eo_prepare_6(t); break;
      case 7: eo_prepare_7(t); break;
      case 8: eo_prepare_8(t); break;
    }
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* cell-loop part of vmult (poisson_operator.h:310): dst = sum_cells scatter(apply(gather src)) */
int oracle_vmult_cells(const oracle_tables *t, long n_cells, long n_local, const uint32_t *eidx,
                       const double *coef, const double *src, double *dst)
{
  cell_fn f = pick(t->degree);
  if (!f)
    return -1;
  prepare(t);
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n_local; ++i)
    dst[i] = 0.;
  if (t->n_colors > 0)
    {
      for (int col = 0; col < t->n_colors; ++col)
        {
          const long k0 = t->color_start[col], k1 = t->color_start[col + 1];
#pragma omp parallel for schedule(static)
          for (long k = k0; k < k1; k += VL)
            f(t, t->color_cells + k, (int)(k1 - k < VL ? k1 - k : VL), eidx, coef, src, dst, 0);
        }
      return 0;
    }
#pragma omp parallel for schedule(dynamic, 4)
  for (long c = 0; c < n_cells; c += VL)
    {
      long cells[VL];
      for (int v = 0; v < VL; ++v)
        cells[v] = c + v < n_cells ? c + v : c;
      f(t, cells, (int)(n_cells - c < VL ? n_cells - c : VL), eidx, coef, src, dst, 1);
    }
  return 0;
}

/* LaplaceOperator::vmult, poisson_operator.h:307-313 */
int oracle_vmult(const oracle_tables *t, long n_cells, long n_local, const uint32_t *eidx,
                 const double *coef, long n_con, const uint32_t *con, const double *src, double *dst)
{
  int e = oracle_vmult_cells(t, n_cells, n_local, eidx, coef, src, dst);
  for (long i = 0; i < n_con; ++i)
    dst[con[i]] = src[con[i]];
  return e;
}

static double dot(long n, const double *a, const double *b)
{
  double s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (long i = 0; i < n; ++i)
    s += a[i] * b[i];
  return s;
}

/* deal.II ReductionControl::check: 0 iterate, 1 success, 2 failure */
static int check(int step, double v, int max_steps, double tol, double reduce, double *reduced_tol)
{
  if (step == 0)
    *reduced_tol = v * reduce;
  if (v < *reduced_tol || v <= tol) /* strict for the reduced tolerance, as in deal.II */
    return 1;
  if (step >= max_steps || isnan(v))
    return 2;
  return 0;
}

/* deal.II 9.3 SolverCG::solve + DiagonalMatrixBlocked (benchmark_precond/bench.cc:11-16,
 * diagonal_matrix_blocked.h:13-27).  x must be zero on entry.  Returns last_step. */
int oracle_cg_plain(const oracle_tables *t, long n_cells, long n, const uint32_t *eidx,
                    const double *coef, long n_con, const uint32_t *con, const double *diag,
                    const double *b, double *x, int max_steps, double tol, double reduce,
                    double *history)
{
  double *g = malloc(sizeof(double) * n), *d = malloc(sizeof(double) * n),
         *h = malloc(sizeof(double) * n);
  double reduced_tol = 0;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i)
    g[i] = -b[i];
  double res = sqrt(dot(n, g, g));
  int    it  = 0;
  if (history)
    history[0] = res;
  if (check(0, res, max_steps, tol, reduce, &reduced_tol) == 0)
    {
#pragma omp parallel for schedule(static)
      for (long i = 0; i < n; ++i)
        {
          h[i] = diag[i / 3] * g[i];
          d[i] = -h[i];
        }
      double gh = dot(n, g, h);
      for (;;)
        {
          ++it;
          oracle_vmult(t, n_cells, n, eidx, coef, n_con, con, d, h);
          const double alpha = gh / dot(n, d, h);
          double       gg    = 0;
#pragma omp parallel for reduction(+ : gg) schedule(static)
          for (long i = 0; i < n; ++i)
            {
              x[i] += alpha * d[i];
              g[i] += alpha * h[i];
              gg += g[i] * g[i];
            }
          res = sqrt(fabs(gg));
          if (history)
            history[it] = res;
          if (check(it, res, max_steps, tol, reduce, &reduced_tol) != 0)
            break;
          double ghn = 0;
#pragma omp parallel for reduction(+ : ghn) schedule(static)
          for (long i = 0; i < n; ++i)
            {
              h[i] = diag[i / 3] * g[i];
              ghn += g[i] * h[i];
            }
          const double beta = ghn / gh;
          gh                = ghn;
#pragma omp parallel for schedule(static)
          for (long i = 0; i < n; ++i)
            d[i] = beta * d[i] - h[i];
        }
    }
  free(g);
  free(d);
  free(h);
  return it;
}

/* SolverCGFullMerge::solve (solver_cg_optimized.h:192-302) with do_cg_update4b (:65-161) as
 * the pre sweep, the cell loop, and do_cg_update3b (:12-61) as the post sweep
 * (poisson_operator.h:327-377).  x must be zero on entry.  Returns last_step. */
int oracle_cg_merged(const oracle_tables *t, long n_cells, long n, const uint32_t *eidx,
                     const double *coef, const double *diag, const double *b, double *x,
                     int max_steps, double tol, double reduce, double *history)
{
  double *g = malloc(sizeof(double) * n), *d = calloc(n, sizeof(double)),
         *h = calloc(n, sizeof(double));
  double reduced_tol = 0;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i)
    g[i] = -b[i];
  double res = sqrt(dot(n, g, g));
  int    it  = 0;
  if (history)
    history[0] = res;
  int    conv  = check(0, res, max_steps, tol, reduce, &reduced_tol);
  double alpha = 0, beta = 0, alpha_old = 0, beta_old = 0;
  while (conv == 0)
    {
      ++it;
      const double ao = (it % 2 == 1) ? alpha_old : 0.;
      if (alpha == 0.)
        {
#pragma omp parallel for schedule(static)
          for (long i = 0; i < n; ++i)
            d[i] = -diag[i / 3] * g[i];
        }
      else if (ao == 0.)
        {
#pragma omp parallel for schedule(static)
          for (long i = 0; i < n; ++i)
            {
              g[i] += alpha * h[i];
              d[i] = beta * d[i] - diag[i / 3] * g[i];
            }
        }
      else
        {
          const double c1 = alpha + ao / beta_old, c2 = ao / beta_old;
#pragma omp parallel for schedule(static)
          for (long i = 0; i < n; ++i)
            {
              x[i] += c1 * d[i] + c2 * diag[i / 3] * g[i];
              g[i] += alpha * h[i];
              d[i] = beta * d[i] - diag[i / 3] * g[i];
            }
        }
      oracle_vmult_cells(t, n_cells, n, eidx, coef, d, h);
      double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0;
#pragma omp parallel for reduction(+ : s0, s1, s2, s3, s4, s5, s6) schedule(static)
      for (long i = 0; i < n; ++i)
        {
          const double pr = diag[i / 3], zi = pr * h[i];
          s0 += d[i] * h[i];
          s1 += h[i] * h[i];
          s2 += g[i] * h[i];
          s3 += g[i] * g[i];
          s6 += g[i] * pr * g[i];
          s4 += g[i] * zi;
          s5 += h[i] * zi;
        }
      alpha_old = alpha;
      beta_old  = beta;
      alpha     = s6 / s0;
      res       = sqrt(s3 + 2 * alpha * s2 + alpha * alpha * s1);
      if (history)
        history[it] = res;
      conv = check(it, res, max_steps, tol, reduce, &reduced_tol);
      if (conv != 0)
        {
          if (it % 2 == 1)
            {
#pragma omp parallel for schedule(static)
              for (long i = 0; i < n; ++i)
                x[i] += alpha * d[i];
            }
          else
            {
              const double c1 = alpha + alpha_old / beta_old, c2 = alpha_old / beta_old;
#pragma omp parallel for schedule(static)
              for (long i = 0; i < n; ++i)
                x[i] += c1 * d[i] + c2 * diag[i / 3] * g[i];
            }
          break;
        }
      beta = alpha * (s4 + alpha * s5) / s6;
    }
  free(g);
  free(d);
  free(h);
  return it;
}


/* ---------------------------------------------------------------------------------------------
 * Cache-blocked merged CG: the reference's MatrixFree::cell_loop with pre/post hooks
 * (poisson_operator.h:339-364).  Every thread owns a contiguous chunk of cell-batch ranges of the
 * (Morton-ordered) loop -- the role of an MPI rank's subdomain -- and walks it in order:
 *   do_cg_update4b on the DoFs private to the range (touched by no other range: the first group
 *   of Renumber's cellbatch_range grouping, a contiguous run), the range's cells, do_cg_update3b
 *   on the same run while it is still in cache.  The remaining DoFs (shared between ranges,
 *   Dirichlet) get the two vector kernels as sweeps before / after the loop.
 * Cells whose DoFs are all touched by one thread only are scattered with plain adds, the others
 * (on the chunk surfaces) with atomics.  Same arithmetic as oracle_cg_merged up to summation order.
 * ------------------------------------------------------------------------------------------- */
typedef struct
{
  long           n_ranges;
  const long    *range_cell; /* [n_ranges + 1] */
  const long    *range_priv; /* [n_ranges + 1] private DoF runs */
  int            n_threads;
  long          *chunk;      /* [n_threads + 1] first range of every thread */
  unsigned char *unsafe;     /* [n_ranges] some DoF of the range is touched by another thread */
} blocked_plan;

static void plan_build(blocked_plan *pl, const oracle_tables *t, long n_local, const uint32_t *eidx,
                       long n_ranges, const long *range_cell, const long *range_priv)
{
  pl->n_ranges   = n_ranges;
  pl->range_cell = range_cell;
  pl->range_priv = range_priv;
  pl->n_threads  = oracle_num_threads();
  pl->chunk      = malloc(sizeof(long) * (pl->n_threads + 1));
  pl->unsafe     = calloc(n_ranges > 0 ? n_ranges : 1, 1);
  for (int k = 0; k <= pl->n_threads; ++k)
    pl->chunk[k] = n_ranges * k / pl->n_threads;
  /* first thread to touch every entity (by its first node), then mark the ranges that meet
   * an entity first touched by another thread */
  const long     n_nodes = n_local / 3 + 1;
  unsigned char *owner   = malloc(n_nodes);
  memset(owner, 255, n_nodes);
  for (int k = 0; k < pl->n_threads; ++k)
    for (long r = pl->chunk[k]; r < pl->chunk[k + 1]; ++r)
      for (long c = range_cell[r]; c < range_cell[r + 1]; ++c)
        for (int e = 0; e < 27; ++e)
          {
            const uint32_t base = eidx[27 * c + e];
            if (base == 0xFFFFFFFFu)
              continue;
            if (owner[base / 3] == 255)
              owner[base / 3] = (unsigned char)k;
            else if (owner[base / 3] != (unsigned char)k)
              owner[base / 3] = 254; /* shared between threads */
          }
#pragma omp parallel for schedule(static)
  for (long r = 0; r < n_ranges; ++r)
    for (long c = range_cell[r]; c < range_cell[r + 1]; ++c)
      for (int e = 0; e < 27; ++e)
        {
          const uint32_t base = eidx[27 * c + e];
          if (base != 0xFFFFFFFFu && owner[base / 3] == 254)
            pl->unsafe[r] = 1;
        }
  free(owner);
  (void)t;
}

static void plan_free(blocked_plan *pl)
{
  free(pl->chunk);
  free(pl->unsafe);
}

static inline void upd4b(long i0, long i1, double *h, double *x, double *g, double *d, const double *diag,
                         double alpha, double beta, double ao, double beta_old)
{
  if (alpha == 0.)
    for (long i = i0; i < i1; ++i)
      {
        d[i] = -diag[i / 3] * g[i];
        h[i] = 0.;
      }
  else if (ao == 0.)
    for (long i = i0; i < i1; ++i)
      {
        g[i] += alpha * h[i];
        d[i] = beta * d[i] - diag[i / 3] * g[i];
        h[i] = 0.;
      }
  else
    {
      const double c1 = alpha + ao / beta_old, c2 = ao / beta_old;
      for (long i = i0; i < i1; ++i)
        {
          x[i] += c1 * d[i] + c2 * diag[i / 3] * g[i];
          g[i] += alpha * h[i];
          d[i] = beta * d[i] - diag[i / 3] * g[i];
          h[i] = 0.;
        }
    }
}

static inline void upd3b(long i0, long i1, const double *h, const double *g, const double *d,
                         const double *diag, double *s)
{
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0;
  for (long i = i0; i < i1; ++i)
    {
      const double pr = diag[i / 3], zi = pr * h[i];
      s0 += d[i] * h[i];
      s1 += h[i] * h[i];
      s2 += g[i] * h[i];
      s3 += g[i] * g[i];
      s4 += g[i] * zi;
      s5 += h[i] * zi;
      s6 += g[i] * pr * g[i];
    }
  s[0] += s0, s[1] += s1, s[2] += s2, s[3] += s3, s[4] += s4, s[5] += s5, s[6] += s6;
}

int oracle_cg_merged_blocked(const oracle_tables *t, long n_cells, long n, const uint32_t *eidx,
                             const double *coef, const double *diag, const double *b, double *x,
                             int max_steps, double tol, double reduce, double *history, long n_ranges,
                             const long *range_cell, const long *range_priv)
{
  cell_fn f = pick(t->degree);
  if (!f || n_ranges <= 0 || range_cell[n_ranges] != n_cells)
    return -1;
  prepare(t);
  blocked_plan pl;
  plan_build(&pl, t, n, eidx, n_ranges, range_cell, range_priv);
  const long tail = range_priv[n_ranges];
  double    *g = malloc(sizeof(double) * n), *d = calloc(n, sizeof(double)), *h = calloc(n, sizeof(double));
  double     reduced_tol = 0;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i)
    g[i] = -b[i];
  double res = sqrt(dot(n, g, g));
  int    it  = 0;
  if (history)
    history[0] = res;
  int    conv  = check(0, res, max_steps, tol, reduce, &reduced_tol);
  double alpha = 0, beta = 0, alpha_old = 0, beta_old = 0;
  while (conv == 0)
    {
      ++it;
      const double ao = (it % 2 == 1) ? alpha_old : 0.;
      double       S[7] = {0, 0, 0, 0, 0, 0, 0};
#pragma omp parallel
      {
        double s[7] = {0, 0, 0, 0, 0, 0, 0};
#ifdef _OPENMP
        const int k = omp_get_thread_num(), nt = omp_get_num_threads();
#else
        const int k = 0, nt = 1;
#endif
        /* shared DoFs first: every thread a slice of the tail */
        {
          const long len = n - tail, i0 = tail + len * k / nt, i1 = tail + len * (k + 1) / nt;
          upd4b(i0, i1, h, x, g, d, diag, alpha, beta, ao, beta_old);
        }
#pragma omp barrier
        if (k < pl.n_threads)
          for (long r = pl.chunk[k]; r < pl.chunk[k + 1]; ++r)
            {
              upd4b(range_priv[r], range_priv[r + 1], h, x, g, d, diag, alpha, beta, ao, beta_old);
              for (long c = range_cell[r]; c < range_cell[r + 1]; c += VL)
                {
                  long      cells[VL];
                  const int nl = (int)(range_cell[r + 1] - c < VL ? range_cell[r + 1] - c : VL);
                  for (int v = 0; v < VL; ++v)
                    cells[v] = c + (v < nl ? v : 0);
                  f(t, cells, nl, eidx, coef, d, h, pl.unsafe[r]);
                }
              if (!pl.unsafe[r]) /* nobody else adds into this range's private run */
                upd3b(range_priv[r], range_priv[r + 1], h, g, d, diag, s);
            }
#pragma omp barrier
        /* private runs of the ranges on chunk surfaces (other threads' atomics had to finish) */
        if (k < pl.n_threads)
          for (long r = pl.chunk[k]; r < pl.chunk[k + 1]; ++r)
            if (pl.unsafe[r])
              upd3b(range_priv[r], range_priv[r + 1], h, g, d, diag, s);
        {
          const long len = n - tail, i0 = tail + len * k / nt, i1 = tail + len * (k + 1) / nt;
          upd3b(i0, i1, h, g, d, diag, s);
        }
#pragma omp critical
        for (int q = 0; q < 7; ++q)
          S[q] += s[q];
      }
      alpha_old = alpha;
      beta_old  = beta;
      alpha     = S[6] / S[0];
      res       = sqrt(S[3] + 2 * alpha * S[2] + alpha * alpha * S[1]);
      if (history)
        history[it] = res;
      conv = check(it, res, max_steps, tol, reduce, &reduced_tol);
      if (conv != 0)
        {
          if (it % 2 == 1)
            {
#pragma omp parallel for schedule(static)
              for (long i = 0; i < n; ++i)
                x[i] += alpha * d[i];
            }
          else
            {
              const double c1 = alpha + alpha_old / beta_old, c2 = alpha_old / beta_old;
#pragma omp parallel for schedule(static)
              for (long i = 0; i < n; ++i)
                x[i] += c1 * d[i] + c2 * diag[i / 3] * g[i];
            }
          break;
        }
      beta = alpha * (S[4] + alpha * S[5]) / S[6];
    }
  plan_free(&pl);
  free(g);
  free(d);
  free(h);
  return it;
}
