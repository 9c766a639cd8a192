// bp4_kernels.cuh -- sm_100a kernels of the BP4 hot path (FP64).
//   cell kernel, plain : dst += A src                              (a1/a3/a6/a7 of SURVEY 8a)
//   cell kernel, merged: do_cg_update4b fused into the gather, do_cg_update3b fused into
//                        the scatter through a last-toucher protocol        (a2/a8/a9)
//   streaming kernels for the unfused variant, the plain-CG BLAS-1 and the Jacobi apply.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bp4_cell.cuh"

namespace bp4
{
  // 128 threads (one warp per SM sub-partition) and two resident blocks per SM: the block
  // may use up to 255 registers per thread (phase 2 keeps 9*Q doubles live), and while one
  // block waits on its gather or a barrier the other keeps the FP64 pipe busy.
  constexpr int kThreads     = 128;
#ifndef BP4_BLOCKS_PER_SM
#  define BP4_BLOCKS_PER_SM 2
#endif
  constexpr int kBlocksPerSM = BP4_BLOCKS_PER_SM;
  constexpr int kSmemBudget  = (227 * 1024) / kBlocksPerSM - 512; // per block

  template <int P>
  struct Cfg
  {
    using G = Geom<P>;
    // resident blocks per SM of the classic cell kernel: three (<= 168 registers, smaller batches)
    // measured faster at Q2 (+12 %) and Q6 (+13 %), slower or equal elsewhere
    static constexpr int BLOCKS   = (P == 2 || P == 6) ? 3 : kBlocksPerSM;
    static constexpr int budget   = (227 * 1024) / BLOCKS - 512;
    static constexpr int per_cell = (G::WORK + 2 * 24) * 8 + 2 * 28 * 4 + 28 + 64 * 5 + 16;
    static constexpr int fit      = (budget - 4 * G::DOF - 16 * G::Q - 256) / per_cell;
    // phase 2 carries ~2/3 of the FP64 work: prefer Q^2*CPB close to a multiple of the block
    static constexpr int want = (2 * kThreads) / (G::Q * G::Q) > 0 ? (2 * kThreads) / (G::Q * G::Q) : 1;
    static constexpr int CPB  = fit < 1 ? 1 : (fit < want ? fit : want);
  };

  template <int P, int CPB>
  struct alignas(16) CellSmem
  {
    using G = Geom<P>;
    // the gathered DoFs, the three phases and the result all live in the work rows: gather
    // fills the first N*N slots of each row, phases 1 and 3 run in place
    double   work[CPB * G::WORK];
    double   coef[2][CPB][24]; // double-buffered: the next batch's metadata is prefetched
    double   xq[G::Q];
    double   wq[G::Q];
    uint32_t eidx[2][CPB][28];
    uint32_t dtab[G::DOF];
    // merged kernel only
    uint8_t  meta[CPB][28];
    uint32_t chunk_base[CPB * 64];
    uint8_t  chunk_len[CPB * 64];
    uint32_t n_chunks;
  };

  struct CellArgs
  {
    const uint32_t *entity_index; // [n_cells][27]
    const double   *coef;         // [n_cells][24] tri-linear coefficients
    const uint32_t *dtab;         // [3 N^3] gather/scatter table (build_dof_table)
    uint64_t        n_cells;
    const double   *src;
    double         *dst;
  };

  // entity meta byte: bits 0-3 = (number of local cells touching the entity) - 1,
  // bit 7 = this cell is the entity's owner (the one that writes r, p, x in the fused pre)
  constexpr uint8_t kMetaOwner = 0x80;

  struct MergedArgs
  {
    const uint32_t *entity_index;
    const double   *coef;
    const uint32_t *dtab;
    const uint8_t  *meta;     // [n_cells][28]
    uint32_t       *counters; // [n_nodes] arrival counters, indexed by first node of the entity
    uint64_t        n_cells;
    const double   *r_old, *p_old;
    double         *h_old;    // read in the pre, zeroed on shared entities after their post
    double         *r_new, *p_new, *h_new, *x;
    const double   *prec;
    double          alpha, beta, c1, c2; // c1 = alpha + alpha_old/beta_old, c2 = alpha_old/beta_old
    int             update_x;            // alpha_old != 0
    double         *acc;                 // [7]
  };

  // ---- trio variant: 256 threads, <= 128 registers, phase 2 split over three lanes per y-line --
  constexpr int kTrioThreads = 256;
  template <int P>
  struct TrioCfg
  {
    using G = Geom<P>;
    static constexpr int per_cell = (G::WORK + 24) * 8 + 28 * 4 + 16;
    static constexpr int fit      = (kSmemBudget - 4 * G::DOF - 16 * G::Q - 256) / per_cell;
    // lines per round = 10 per warp; prefer a batch that fills whole rounds
    static constexpr int lines_per_round = 10 * (kTrioThreads / 32);
    static constexpr int want = (3 * lines_per_round) / (G::Q * G::Q) > 0 ? (3 * lines_per_round) / (G::Q * G::Q) : 1;
    static constexpr int CPB  = fit < 1 ? 1 : (fit < want ? fit : want);
  };

  // ---- prefetching variant of the plain cell kernel -------------------------------------------
  // gather of batch i+1 is issued with cp.async (LDGSTS, zero-filled for Dirichlet entities)
  // right after phase 1 of batch i has consumed the input staging, so its latency hides behind
  // phases 2 and 3; the scatter is fire-and-forget (RED).
  template <int P>
  struct PfCfg
  {
    using G = Geom<P>;
    static constexpr int per_cell = (G::WORK + G::DOFS + 2 * 24) * 8 + 2 * 28 * 4 + 16;
    static constexpr int fit      = (kSmemBudget - 4 * G::DOF - 16 * G::Q - 512) / per_cell;
    static constexpr int want = (2 * kThreads) / (G::Q * G::Q) > 0 ? (2 * kThreads) / (G::Q * G::Q) : 1;
    static constexpr int CPB  = fit < 1 ? 1 : (fit < want ? fit : want);
  };

  template <int P, int CPB>
  struct alignas(16) PfSmem
  {
    using G = Geom<P>;
    double   work[CPB * G::WORK];
    double   dofs[CPB * G::DOFS];
    double   coef[2][CPB][24];
    double   xq[G::Q];
    double   wq[G::Q];
    uint32_t eidx[2][CPB][28];
    uint32_t dtab[G::DOF];
  };

  // ---- TMA variant of the plain cell kernel ---------------------------------------------------
  // The 27 entity segments of a cell are contiguous in the vector, so they move as BULK copies:
  //   gather : cp.async.bulk.shared.global (UBLKCP) per entity, completion on an mbarrier,
  //            issued one batch ahead (double-buffered stage) so the latency hides behind
  //            phases 2 and 3;
  //   scatter: cp.reduce.async.bulk.global.shared .add.f64 (UBLKRED) per entity: the FP64
  //            scatter-add is done by the TMA unit at L2, not by 375 per-lane REDs per cell.
  // Segments start at multiples of 24 B; a segment whose first DoF index is odd is moved from
  // its 16-byte-aligned predecessor with one or two pad doubles (gathered pads are never read,
  // scattered pads are 0.0 and add nothing).  Vectors carry two doubles of slack for that.
  template <int P>
  struct TmaCfg
  {
    using G = Geom<P>;
    static constexpr int per_cell = (G::WORK + 2 * Stage<P>::SIZE + 2 * 24) * 8 + 2 * 28 * 4 + 2 * 28 * 2 + 16;
    static constexpr int fit      = (kSmemBudget - 2 * G::DOF - 16 * G::Q - 512) / per_cell;
    static constexpr int want = (2 * kThreads) / (G::Q * G::Q) > 0 ? (2 * kThreads) / (G::Q * G::Q) : 1;
    static constexpr int CPB  = fit < 1 ? 1 : (fit < want ? fit : want);
  };

  template <int P, int CPB>
  struct alignas(16) TmaSmem
  {
    using G = Geom<P>;
    double             work[CPB * G::WORK];
    alignas(16) double stage_in[CPB * Stage<P>::SIZE]; // bulk copies need 16-byte aligned smem
    alignas(16) double stage_out[CPB * Stage<P>::SIZE];
    double             coef[2][CPB][24];
    double             xq[G::Q];
    double             wq[G::Q];
    unsigned long long mbar;
    uint32_t           eidx[2][CPB][28];
    uint16_t           off[2][CPB][28]; // slot + (first DoF & 1) of every entity
    uint16_t           slot[28];
    uint16_t           itab[G::N * G::N * G::ROWS];
  };

  struct TmaArgs
  {
    const uint32_t *entity_index;
    const double   *coef;
    const uint16_t *slot; // [27]
    const uint16_t *itab; // [N*N][ROWS]
    uint64_t        n_cells;
    const double   *src;
    double         *dst;
  };

  // ---- warp-specialised variant ------------------------------------------------------------
  // 256 threads: warps 0-3 = compute warpgroup (phases 1-3, FP64 only), warps 4-7 = memory
  // warpgroup (metadata, gather + fused pre, scatter + fused post).  setmaxnreg moves registers
  // from the memory warps to the compute warps; two blocks per SM.
  constexpr int kWsThreads  = 256;
  constexpr int kWsRegsComp = 192;
  constexpr int kWsRegsMem  = 64;

  template <int P>
  struct WsCfg
  {
    using G = Geom<P>;
    static constexpr int per_cell = (G::WORK + G::DOFS + 2 * 24) * 8 + 2 * 28 * 4 + 2 * 28 + 64 * 5 + 16;
    static constexpr int fit      = (kSmemBudget - 4 * G::DOF - 16 * G::Q - 512) / per_cell;
    static constexpr int want = (2 * 128) / (G::Q * G::Q) > 0 ? (2 * 128) / (G::Q * G::Q) : 1;
    static constexpr int CPB  = fit < 1 ? 1 : (fit < want ? fit : want);
  };

  template <int P, int CPB>
  struct alignas(16) WsSmem
  {
    using G = Geom<P>;
    double   work[CPB * G::WORK]; // phase 3 leaves its result in the first N*N slots of each row
    double   dofs[CPB * G::DOFS]; // gathered input
    double   coef[2][CPB][24];
    double   xq[G::Q];
    double   wq[G::Q];
    double   red[7][4];
    uint32_t eidx[2][CPB][28];
    uint32_t dtab[G::DOF];
    uint32_t chunk_base[CPB * 64];
    uint32_t n_chunks;
    uint8_t  meta[2][CPB][28];
    uint8_t  chunk_len[CPB * 64];
  };
} // namespace bp4
