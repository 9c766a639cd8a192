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
#undef P
#define P 3
#include "bp4_oracle_kernel.inc"
#undef P
#define P 4
#include "bp4_oracle_kernel.inc"
#undef P
#define P 5
#include "bp4_oracle_kernel.inc"
#undef P
#define P 6
#include "bp4_oracle_kernel.inc"
#undef P
#define P 7
#include "bp4_oracle_kernel.inc"
#undef P
#define P 8
#include "bp4_oracle_kernel.inc"
#undef P

typedef void (*cell_fn)(const oracle_tables *, const long *, int, const uint32_t *, const double *,
                        const double *, double *, int);
static cell_fn pick(int p)
{
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
  if (v <= *reduced_tol || v <= tol)
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
