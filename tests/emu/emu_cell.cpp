// CPU emulation of the thread loops of cell_kernel_plain (bp4_kernels.cu) around the SAME
// phase functions (bp4_cell.cuh) the GPU kernel runs.  Test-only: lets the CPU test suite
// check the shared-memory indexing and the contraction algebra of the device code against
// the oracle without a GPU.  Never linked into the product library.
#include <cstdint>
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
