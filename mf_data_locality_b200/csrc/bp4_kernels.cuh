// bp4_kernels.cuh -- sm_100a kernels of the BP4 hot path (FP64).
//   cell kernel, MODE_PLAIN : dst += A src                      (a1/a3/a6/a7 of SURVEY 8a)
//   cell kernel, MODE_MERGED: do_cg_update4b fused into the gather, do_cg_update3b fused
//                             into the scatter through a last-toucher protocol (a2/a8/a9)
//   streaming kernels for the unfused variant, the plain-CG BLAS-1 and the Jacobi apply.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bp4_cell.cuh"

namespace bp4
{
  constexpr int kThreads = 256;

  // cells per thread block: Q^2 * CPB close to (but not above) kThreads so that phase 2,
  // which carries ~2/3 of the FP64 work, fills the block's eight warps
  template <int P>
  struct Cfg
  {
    static constexpr int Q   = P + 2;
    static constexpr int CPB = (kThreads / (Q * Q)) > 0 ? (kThreads / (Q * Q)) : 1;
  };

  template <int P, int CPB>
  struct alignas(16) CellSmem
  {
    using G = Geom<P>;
    double   work[CPB][G::WORK];
    double   dofs[CPB][G::DOF];
    double   coef[CPB][24];
    double   xq[G::Q];
    double   wq[G::Q];
    uint32_t eidx[CPB][28];
    uint32_t walk[G::N3];
  };

  struct CellArgs
  {
    const uint32_t *entity_index; // [n_cells][27]
    const double   *coef;         // [n_cells][24] tri-linear coefficients
    const uint32_t *walk;         // [N^3] packed entity walk (build_walk)
    uint64_t        n_cells;
    const double   *src;
    double         *dst;
  };

  template <int P>
  struct TabSym; // per-degree __constant__ table, defined in bp4_kernels.cu

} // namespace bp4
