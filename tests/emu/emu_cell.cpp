// CPU emulation of the thread loops of cell_kernel_plain (bp4_kernels.cu) around the SAME
// phase functions (bp4_cell.cuh) the GPU kernel runs.  Test-only: lets the CPU test suite
// check the shared-memory indexing and the contraction algebra of the device code against
// the oracle without a GPU.  Never linked into the product library.
#include <cstdint>
#include <cmath>
#include <algorithm>
#include <cstring>
#include <array>
#include <vector>

#include "../../mf_data_locality_b200/csrc/bp4_tables.h"

template <int P>
static void run(long n_cells, const uint32_t *eidx, const double *verts, const double *src, double *dst)
{
  using G = bp4::Geom<P>;
  bp4::Tab<P> tb;
  bp4::fill_tab<P>(tb);
  std::vector<uint32_t> dtab(G::DOF);
  bp4::build_dof_table<P>(dtab.data());
  std::vector<double> dofs(G::DOFS), work(G::WORK);
  for (long cell = 0; cell < n_cells; ++cell)
    {
      const uint32_t *e = eidx + 27 * cell;
      const double   *v = verts + 24 * cell;
      double          cf[24];
      for (int k = 0; k < 3; ++k)
        {
          cf[0 + k]  = v[0 + k];
          cf[3 + k]  = v[3 + k] - v[0 + k];
          cf[6 + k]  = v[6 + k] - v[0 + k];
          cf[9 + k]  = v[9 + k] - v[6 + k] - (v[3 + k] - v[0 + k]);
          cf[12 + k] = v[12 + k] - v[0 + k];
          cf[15 + k] = v[15 + k] - v[12 + k] - (v[3 + k] - v[0 + k]);
          cf[18 + k] = v[18 + k] - v[12 + k] - (v[6 + k] - v[0 + k]);
          cf[21 + k] = (v[21 + k] - v[18 + k] - (v[15 + k] - v[12 + k]) -
                        (v[9 + k] - v[6 + k] - (v[3 + k] - v[0 + k])));
        }
      for (int m = 0; m < G::DOF; ++m)
        {
          const uint32_t t = dtab[m], base = e[bp4::dtab_ent(t)];
          dofs[bp4::dtab_off<P>(t)] = base != 0xFFFFFFFFu ? src[(size_t)base + bp4::dtab_rel(t)] : 0.;
        }
      for (int it = 0; it < G::ITEMS13; ++it)
        bp4::phase1<P>(tb, dofs.data() + it * G::RD, work.data() + it * G::RW);
      for (int it = 0; it < G::ITEMS2; ++it)
        {
          const int qz = it / G::Q, qx = it % G::Q;
          bp4::phase2<P>(tb, cf, work.data(), qx, qz, tb.xq[qx], tb.xq[qz], tb.wq[qx] * tb.wq[qz]);
        }
      for (int it = 0; it < G::ITEMS13; ++it)
        bp4::phase3<P>(tb, work.data() + it * G::RW, dofs.data() + it * G::RD);
      for (int m = 0; m < G::DOF; ++m)
        {
          const uint32_t t = dtab[m], base = e[bp4::dtab_ent(t)];
          if (base != 0xFFFFFFFFu)
            dst[(size_t)base + bp4::dtab_rel(t)] += dofs[bp4::dtab_off<P>(t)];
        }
    }
}

// same, with phases 1 and 3 as fine-grained sweeps, in place in the work rows (the high-degree
// path of the kernel); each inner loop is what the threads of a block do between two barriers
template <int P>
static void run_fine(long n_cells, const uint32_t *eidx, const double *verts, const double *src, double *dst)
{
  using G = bp4::Geom<P>;
  bp4::Tab<P> tb;
  bp4::fill_tab<P>(tb);
  std::vector<uint32_t> dtab(G::DOF);
  bp4::build_dof_table<P>(dtab.data());
  std::vector<double> work(G::WORK);
  for (long cell = 0; cell < n_cells; ++cell)
    {
      const uint32_t *e = eidx + 27 * cell;
      const double   *v = verts + 24 * cell;
      double          cf[24];
      for (int k = 0; k < 3; ++k)
        {
          cf[0 + k]  = v[0 + k];
          cf[3 + k]  = v[3 + k] - v[0 + k];
          cf[6 + k]  = v[6 + k] - v[0 + k];
          cf[9 + k]  = v[9 + k] - v[6 + k] - (v[3 + k] - v[0 + k]);
          cf[12 + k] = v[12 + k] - v[0 + k];
          cf[15 + k] = v[15 + k] - v[12 + k] - (v[3 + k] - v[0 + k]);
          cf[18 + k] = v[18 + k] - v[12 + k] - (v[6 + k] - v[0 + k]);
          cf[21 + k] = (v[21 + k] - v[18 + k] - (v[15 + k] - v[12 + k]) -
                        (v[9 + k] - v[6 + k] - (v[3 + k] - v[0 + k])));
        }
      for (int m = 0; m < G::DOF; ++m)
        {
          const uint32_t t = dtab[m], base = e[bp4::dtab_ent(t)];
          work[bp4::dtab_off_work<P>(t)] = base != 0xFFFFFFFFu ? src[(size_t)base + bp4::dtab_rel(t)] : 0.;
        }
      // items in reverse order on purpose: a hazard between items of one sweep would show
      for (int it = G::ROWS * G::N - 1; it >= 0; --it)
        bp4::phase1a<P>(tb, work.data() + (it % G::ROWS) * G::RW, it / G::ROWS);
      for (int it = G::ROWS * G::Q - 1; it >= 0; --it)
        bp4::phase1b<P>(tb, work.data() + (it % G::ROWS) * G::RW, it / G::ROWS);
      for (int it = G::ROWS * G::Q - 1; it >= 0; --it)
        bp4::phase1c<P>(tb, work.data() + (it % G::ROWS) * G::RW, it / G::ROWS);
      for (int it = 0; it < G::ITEMS2; ++it)
        {
          const int qz = it / G::Q, qx = it % G::Q;
          bp4::phase2<P>(tb, cf, work.data(), qx, qz, tb.xq[qx], tb.xq[qz], tb.wq[qx] * tb.wq[qz]);
        }
      for (int it = G::ROWS * G::Q - 1; it >= 0; --it)
        bp4::phase3a<P>(tb, work.data() + (it % G::ROWS) * G::RW, it / G::ROWS);
      for (int it = G::ROWS * G::Q - 1; it >= 0; --it)
        bp4::phase3b<P>(tb, work.data() + (it % G::ROWS) * G::RW, it / G::ROWS);
      for (int it = G::ROWS * G::N - 1; it >= 0; --it)
        bp4::phase3c<P>(tb, work.data() + (it % G::ROWS) * G::RW, it / G::ROWS);
      for (int m = 0; m < G::DOF; ++m)
        {
          const uint32_t t = dtab[m], base = e[bp4::dtab_ent(t)];
          if (base != 0xFFFFFFFFu)
            dst[(size_t)base + bp4::dtab_rel(t)] += work[bp4::dtab_off_work<P>(t)];
        }
    }
}

extern "C" int emu_vmult_cells_fine(int p, long n_cells, const uint32_t *eidx, const double *verts,
                                     const double *src, double *dst)
{
  switch (p)
    {
      case 2: run_fine<2>(n_cells, eidx, verts, src, dst); return 0;
      case 3: run_fine<3>(n_cells, eidx, verts, src, dst); return 0;
      case 4: run_fine<4>(n_cells, eidx, verts, src, dst); return 0;
      case 5: run_fine<5>(n_cells, eidx, verts, src, dst); return 0;
      case 6: run_fine<6>(n_cells, eidx, verts, src, dst); return 0;
      case 7: run_fine<7>(n_cells, eidx, verts, src, dst); return 0;
      case 8: run_fine<8>(n_cells, eidx, verts, src, dst); return 0;
    }
  return -1;
}

extern "C" int emu_vmult_cells(int p, long n_cells, const uint32_t *eidx, const double *verts,
                               const double *src, double *dst)
{
  switch (p)
    {
      case 2: run<2>(n_cells, eidx, verts, src, dst); return 0;
      case 3: run<3>(n_cells, eidx, verts, src, dst); return 0;
      case 4: run<4>(n_cells, eidx, verts, src, dst); return 0;
      case 5: run<5>(n_cells, eidx, verts, src, dst); return 0;
      case 6: run<6>(n_cells, eidx, verts, src, dst); return 0;
      case 7: run<7>(n_cells, eidx, verts, src, dst); return 0;
      case 8: run<8>(n_cells, eidx, verts, src, dst); return 0;
    }
  return -1;
}

extern "C" void emu_tables(int p, double *S, double *Dn, double *D, double *xq, double *wq)
{
#define T(PP) case PP: { bp4::Tab<PP> tb; bp4::fill_tab<PP>(tb); memcpy(S, tb.S, sizeof(tb.S)); memcpy(Dn, tb.Dn, sizeof(tb.Dn)); memcpy(D, tb.D, sizeof(tb.D)); memcpy(xq, tb.xq, sizeof(tb.xq)); memcpy(wq, tb.wq, sizeof(tb.wq)); } break;
  switch (p) { T(2) T(3) T(4) T(5) T(6) T(7) T(8) }
#undef T
}

// even-odd 1-D contractions against the dense sums, for every matrix and both directions:
// returns the largest |eo - dense| over random inputs (scaled by the largest |dense| entry)
template <int P>
static double eo_check()
{
  constexpr int N = P + 1, Q = P + 2;
  bp4::Tab<P>   tb;
  bp4::fill_tab<P>(tb);
  double   worst = 0., scale = 1e-300;
  uint64_t st    = 0x9E3779B97F4A7C15ull + P;
  auto     rnd   = [&]() {
    st = st * 6364136223846793005ull + 1442695040888963407ull;
    return double(int64_t(st >> 11)) / double(1ll << 52) - 1.0;
  };
  auto cmp = [&](const double got, const double want) {
    worst = std::max(worst, std::fabs(got - want));
    scale = std::max(scale, std::fabs(want));
  };
  for (int rep = 0; rep < 8; ++rep)
    {
      double xn[N], xq[Q], o_q[Q], o_n[N], o_qq[Q];
      for (double &v : xn)
        v = rnd();
      for (double &v : xq)
        v = rnd();
      bp4::eo_first<N, Q, 1>(tb.Sfp, tb.Sfm, tb.S, xn, o_q);
      for (int q = 0; q < Q; ++q)
        {
          double s = 0;
          for (int i = 0; i < N; ++i)
            s += tb.S[i][q] * xn[i];
          cmp(o_q[q], s);
        }
      bp4::eo_first<N, Q, -1>(tb.Dnfp, tb.Dnfm, tb.Dn, xn, o_q);
      for (int q = 0; q < Q; ++q)
        {
          double s = 0;
          for (int i = 0; i < N; ++i)
            s += tb.Dn[i][q] * xn[i];
          cmp(o_q[q], s);
        }
      bp4::eo_first<Q, Q, -1>(tb.Dfp, tb.Dfm, tb.D, xq, o_qq);
      for (int q = 0; q < Q; ++q)
        {
          double s = 0;
          for (int i = 0; i < Q; ++i)
            s += tb.D[i][q] * xq[i];
          cmp(o_qq[q], s);
        }
      bp4::eo_second<N, Q, 1>(tb.Ssp, tb.Ssm, tb.S, xq, o_n);
      for (int i = 0; i < N; ++i)
        {
          double s = 0;
          for (int q = 0; q < Q; ++q)
            s += tb.S[i][q] * xq[q];
          cmp(o_n[i], s);
        }
      bp4::eo_second<N, Q, -1>(tb.Dnsp, tb.Dnsm, tb.Dn, xq, o_n);
      for (int i = 0; i < N; ++i)
        {
          double s = 0;
          for (int q = 0; q < Q; ++q)
            s += tb.Dn[i][q] * xq[q];
          cmp(o_n[i], s);
        }
      bp4::eo_second<Q, Q, -1>(tb.Dsp, tb.Dsm, tb.D, xq, o_qq);
      for (int i = 0; i < Q; ++i)
        {
          double s = 0;
          for (int q = 0; q < Q; ++q)
            s += tb.D[i][q] * xq[q];
          cmp(o_qq[i], s);
        }
    }
  return worst / scale;
}

extern "C" double emu_eo_check(int p)
{
  switch (p)
    {
      case 2: return eo_check<2>();
      case 3: return eo_check<3>();
      case 4: return eo_check<4>();
      case 5: return eo_check<5>();
      case 6: return eo_check<6>();
      case 7: return eo_check<7>();
      case 8: return eo_check<8>();
    }
  return 1e300;
}
