// bp4_kernels.cuh -- sm_100a kernels of the BP4 hot path (FP64).
//   cell kernel, plain : dst += A src                              (a1/a3/a6/a7 of SURVEY 8a)
//   cell kernel, fused : the same loop with do_cg_update4b run on a cell-batch range's private
//                        DoFs before its first cell and do_cg_update3b after its last cell
//                        (poisson_operator.h:339-364)                          (a2/a8/a9)
//   streaming kernels for the DoFs shared between ranges, the plain-CG BLAS-1 and the Jacobi apply.
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
#ifndef BP4_WIDE
#  define BP4_WIDE(P) false
#endif
#ifndef BP4_STAGED
// degrees with a fused cell kernel (in-loop vector updates staged through shared memory)
#  define BP4_STAGED(P) ((P) <= 4)
#endif
  constexpr int kBlocksPerSM = 2;

  // QUAD: all 27 geometry coefficients per cell (81 doubles) instead of the 8 tri-linear ones (24)
  template <int P, bool QUAD = false>
  struct Cfg
  {
    using G = Geom<P>;
    static constexpr int NCOEF = QUAD ? 81 : 24;
    // "wide" configuration: ONE block of 256 threads per SM (255 registers, no spills) whose
    // batch fills phase 2 exactly (Q6: 4 cells = 256 lines, Q7: 3 cells = 243 lines)
    static constexpr bool WIDE    = BP4_WIDE(P);
    static constexpr int  THREADS = WIDE ? 256 : kThreads;
    // resident blocks per SM of the cell kernel: three (<= 168 registers, smaller batches)
    // measured faster at Q2 (+12 %) and Q6 (+13 %), slower or equal elsewhere
    static constexpr int BLOCKS   = WIDE ? 1 : ((P == 2 || P == 6) ? 3 : kBlocksPerSM);
    static constexpr int budget   = (227 * 1024) / BLOCKS - 512;
    static constexpr int per_cell = (G::WORK + 2 * NCOEF) * 8 + 2 * 28 * 4;
    static constexpr int fit      = (budget - 4 * G::DOF - 16 * G::Q - 256) / per_cell;
    // phase 2 carries ~2/3 of the FP64 work: prefer Q^2*CPB close to a multiple of the block
    static constexpr int rounds = WIDE ? 1 : 2;
    static constexpr int want   = (rounds * THREADS) / (G::Q * G::Q) > 0 ? (rounds * THREADS) / (G::Q * G::Q) : 1;
    static constexpr int CPB    = fit < 1 ? 1 : (fit < want ? fit : want);
  };

  template <int P, int CPB, int NCOEF = 24>
  struct alignas(16) CellSmem
  {
    using G = Geom<P>;
    // the gathered DoFs, the three phases and the result all live in the work rows: gather
    // fills the first N*N slots of each row, phases 1 and 3 run in place
    double   work[CPB * G::WORK];
    double   coef[2][CPB][NCOEF]; // double-buffered: the next batch's metadata is prefetched
    double   xq[G::Q];
    double   wq[G::Q];
    double   red[8];           // (spare)
    uint32_t units[8];         // ring of the units this block has claimed
    uint32_t n_claimed;
    uint32_t eidx[2][CPB][28];
    uint32_t dtab[G::DOF];
  };

  // One batch of the fused cell loop: CPB consecutive cells of one unit (= a group of whole
  // cell-batch ranges owned by one thread block) and the two DoF intervals the vector updates
  // are hooked to (MatrixFree::cell_loop's pre/post contract):
  //   [pre_begin, pre_end)   private DoFs of the ranges whose FIRST cell lies in this batch:
  //                          do_cg_update4b must run on them before the batch is gathered
  //   [post_begin, post_end) private DoFs of the ranges whose LAST cell lies in this batch:
  //                          do_cg_update3b runs on them once the batch has been scattered
  // "private" = touched by the cells of exactly one range: the first group of Renumber's
  // cellbatch_range grouping (renumber_dofs_for_mf.h:556-590, :622-671), contiguous per range.
  struct alignas(16) BatchDesc
  {
    uint32_t cell0, n_cells;
    uint32_t pre_begin, pre_end;
    uint32_t post_begin, post_end;
    uint32_t pad[2];
  };

  // Staging rows of the fused cell loop (appended to CellSmem in dynamic shared memory).  The
  // in-loop do_cg_update4b / do_cg_update3b are cut into "jobs" of at most JOB DoFs; asynchronous
  // copies (cp.async + mbarrier) bring a job's r, p, h and diagonal entries into these rows while
  // the block runs a compute phase, the threads consume them from shared memory at the end of
  // that phase: no thread waits for a DRAM round trip or holds loaded values in registers
  // (the register-staged variant made the memory window 3x longer).
  // A job [b, e) may start at an odd index: the copy is widened to 16-byte boundaries, element i
  // sits at row[..][i - (b & ~1)], diagonal entry i/3 at prec[i/3 - ((b/3) & ~1)].
  template <int JOB>
  struct alignas(16) JobSmem
  {
    static constexpr int ROW  = JOB + 2;
    static constexpr int PREC = ((JOB / 3 + 4) + 1) & ~1;
    double             row[3][ROW]; // r, p (= d), h
    double             prec[PREC];
    double             redw[8][8]; // the seven merged sums, one row per warp
    BatchDesc          ring[8];    // descriptors of the batches i-2 .. i+4 of this block
    unsigned long long mbar;
    unsigned long long pad;
  };
  // JOB = DoFs per job that fit next to the cell kernel's own shared memory; a batch's pre / post
  // run is handled by two jobs -> LIMIT.  OK: the degree has a fused kernel at all (the runs of
  // the high degrees are far longer than what is left of the shared memory: Q5 2187 DoFs per range
  // vs 146 per job)
  template <int P>
  struct Stage
  {
    static constexpr long avail = (long)Cfg<P>::budget - (long)sizeof(CellSmem<P, Cfg<P>::CPB, 24>) - 16 - 512 - 256;
    // 3 rows of (JOB + 2) + JOB / 3 + 5 doubles
    static constexpr long cap   = ((avail / 8 - 11) * 3) / 10;
    static constexpr int  JOB   = cap < 64 ? 0 : (int)(cap & ~1L);
    static constexpr int  LIMIT = 2 * JOB;
    static constexpr bool OK    = BP4_STAGED(P) && JOB >= 64;
  };

  struct CellArgs
  {
    const uint32_t *entity_index; // [n_cells][27]
    const double   *coef;         // [n_cells][24] tri-linear or [n_cells][81] quadratic coefficients
    const uint32_t *dtab;         // [3 N^3] gather/scatter table (build_dof_table)
    uint64_t        n_cells;      // plain: cells of this launch, batches are cut on the fly
    const double   *src;          // plain: input vector; fused: the search direction d (= p)
    double         *dst;          // plain: output vector; fused: h
    uint32_t       *sched;        // unit counter of this launch (zero at launch), null: strided
    uint32_t        stagger_ns;   // start offsets of the blocks are spread over [0, stagger_ns)
    uint32_t        claim_depth;  // plain: units a block holds claimed (the current one included)
    // ---- fused (vmult_with_merged_sums) only ----
    const BatchDesc *batch;       // batches of all units
    const uint32_t  *unit_batch;  // [n_units + 1] first batch of every unit of this launch
    uint32_t         n_units;
    double          *r, *p, *x;   // g, d, x of SolverCGFullMerge; p == src
    const double    *prec;        // [n_owned / 3]
    double           alpha, beta, c1, c2; // c1 = alpha + alpha_old/beta_old, c2 = alpha_old/beta_old
    int              first;       // alpha == 0: p = -P r, h = 0 only
    int              update_x;    // alpha_old != 0
    double          *acc;         // [7]
  };
} // namespace bp4
