// bp4_kernels.cu -- kernel definitions and launchers (sm_100a, FP64).  See bp4_kernels.cuh.
#include <algorithm>
#include <cstdio>

#include "bp4_kernels.cuh"
#include "bp4_launch.h"
#include "bp4_tables.h"

// degrees >= BP4_FINE_FROM run phases 1 and 3 as fine-grained sweeps (phase1a..c, phase3a..c)
#ifndef BP4_P2_CALL_FROM
#  define BP4_P2_CALL_FROM 6 // degrees >= this call phase 2 out of line
#endif
#ifndef BP4_PIPE_GATHER
#  define BP4_PIPE_GATHER(P) ((P) <= 4)
#endif
#ifndef BP4_SU
#  define BP4_SU 6 // scatter unroll
#endif
#ifndef BP4_FINE_FROM
#  define BP4_FINE_FROM 6
#endif

namespace bp4
{
  // one table per degree in the constant bank: with fully unrolled contractions every
  // matrix entry becomes a c[bank][offset] operand of a DFMA, no load instruction
  template <int P>
  __constant__ Tab<P> c_tab;

  // ---------------------------------------------------------------------------------------
  // cell kernels
  // ---------------------------------------------------------------------------------------
  constexpr int kGatherUnroll = 4;

  template <int P, int CPB>
  __device__ __forceinline__ void load_tables(CellSmem<P, CPB> &sm, const uint32_t *dtab)
  {
    for (int i = threadIdx.x; i < Geom<P>::DOF; i += kThreads)
      sm.dtab[i] = dtab[i];
    if (threadIdx.x < Geom<P>::Q)
      {
        sm.xq[threadIdx.x] = c_tab<P>.xq[threadIdx.x];
        sm.wq[threadIdx.x] = c_tab<P>.wq[threadIdx.x];
      }
  }

  // phases 1-3 on the nc cells staged in the work rows (in place), with the barriers between them
  template <int P, int CPB>
  __device__ __forceinline__ void apply_staged(CellSmem<P, CPB> &sm, const int nc)
  {
    using G          = Geom<P>;
    constexpr int Q  = G::Q;
    const Tab<P> &tb = c_tab<P>;
    const int     tid = threadIdx.x;
    for (int it = tid; it < nc * G::ITEMS13; it += kThreads)
      phase1<P>(tb, sm.work + it * G::RW, sm.work + it * G::RW);
    __syncthreads();
    for (int it = tid; it < nc * G::ITEMS2; it += kThreads)
      {
        const int cell = it / G::ITEMS2, r = it % G::ITEMS2;
        const int qz = r / Q, qx = r % Q;
        phase2<P>(tb, sm.coef[0][cell], sm.work + cell * G::WORK, qx, qz, sm.xq[qx], sm.xq[qz],
                  sm.wq[qx] * sm.wq[qz]);
      }
    __syncthreads();
    for (int it = tid; it < nc * G::ITEMS13; it += kThreads)
      phase3<P>(tb, sm.work + it * G::RW, sm.work + it * G::RW);
    __syncthreads();
  }

  // optional per-phase clock accounting (compile with -DBP4_PHASE_TIMING, run with
  // BP4_PHASE_TIMING=1): how a warp's time splits over metadata / gather / phases / scatter
#ifdef BP4_PHASE_TIMING
  __device__ unsigned long long g_phase_clk[8];
#  define BP4_TICK_INIT unsigned long long tc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long t0 = clock64();
#  define BP4_TICK(k) { const long long t1 = clock64(); tc[k] += t1 - t0; t0 = t1; }
#  define BP4_TICK_FLUSH if ((threadIdx.x & 31) == 0) for (int k = 0; k < 8; ++k) atomicAdd(&g_phase_clk[k], tc[k]);
#else
#  define BP4_TICK_INIT
#  define BP4_TICK(k)
#  define BP4_TICK_FLUSH
#endif

  // Phase 2 as a real call for the high degrees: its ~100 live doubles get a register allocation
  // of their own instead of competing with whatever the gather/scatter code around it keeps
  // alive (the inlined form spilled 2.6x more after an unrelated change to the gather).
  template <int P>
  __device__ __noinline__ void phase2_call(const uint32_t cf_off, const uint32_t work_off, const int qx,
                                           const int qz, const double x, const double z, const double wxz)
  {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    phase2<P>(c_tab<P>, reinterpret_cast<const double *>(smem_raw + cf_off),
              reinterpret_cast<double *>(smem_raw + work_off), qx, qz, x, z, wxz);
  }

  // Classic cell kernel: every warp does every phase; two blocks per SM overlap one block's
  // memory phases with the other's FP64 phases.  The memory phases are kept short:
  //  * the next batch's metadata (27 indices + 24 coefficients per cell) is loaded into
  //    registers before phase 1 and parked in the other half of a double buffer after phase 3;
  //  * the gather issues all loads of the batch before the first use (one latency, not many);
  //  * the scatter reads its table/index/value operands for several DoFs before the REDs go out.
  template <int P, int CPB>
  __global__ void __launch_bounds__(kThreads, Cfg<P>::BLOCKS) cell_kernel_plain(const CellArgs a)
  {
    using G         = Geom<P>;
    constexpr int Q = G::Q, NN = G::N * G::N;
    // high degrees: phases 1 and 3 as one sweep per 1-D contraction (see phase1a)
    constexpr bool kFine = P >= BP4_FINE_FROM;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CellSmem<P, CPB> &sm  = *reinterpret_cast<CellSmem<P, CPB> *>(smem_raw);
    const int         tid = threadIdx.x;
    const Tab<P>     &tb  = c_tab<P>;
    load_tables<P, CPB>(sm, a.dtab);
    BP4_TICK_INIT

    const uint64_t n_batches = (a.n_cells + CPB - 1) / CPB;
    const int      my_n =
      n_batches > blockIdx.x ? (int)((n_batches - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    auto batch_cells = [&](const int i, uint64_t &cell0) {
      cell0 = ((uint64_t)blockIdx.x + (uint64_t)i * gridDim.x) * CPB;
      return (int)min((uint64_t)CPB, a.n_cells - cell0);
    };
    // metadata items of a batch handled by this thread: at most ME indices and MC coefficients
    constexpr int ME = (CPB * 27 + kThreads - 1) / kThreads, MC = (CPB * 24 + kThreads - 1) / kThreads;
    uint32_t      me[ME];
    double        mc[MC];
    auto          fetch_meta = [&](const int i) {
      uint64_t  cell0;
      const int nc = batch_cells(i, cell0);
#pragma unroll
      for (int u = 0; u < ME; ++u)
        {
          const int k = tid + u * kThreads;
          me[u]       = k < nc * 27 ? __ldg(a.entity_index + cell0 * 27 + k) : 0xFFFFFFFFu;
        }
#pragma unroll
      for (int u = 0; u < MC; ++u)
        {
          const int k = tid + u * kThreads;
          mc[u]       = k < nc * 24 ? __ldg(a.coef + cell0 * 24 + k) : 0.;
        }
    };
    auto park_meta = [&](const int bf) {
#pragma unroll
      for (int u = 0; u < ME; ++u)
        {
          const int k = tid + u * kThreads;
          if (k < CPB * 27)
            sm.eidx[bf][k / 27][k % 27] = me[u];
        }
#pragma unroll
      for (int u = 0; u < MC; ++u)
        {
          const int k = tid + u * kThreads;
          if (k < CPB * 24)
            sm.coef[bf][k / 24][k % 24] = mc[u];
        }
    };
    if (my_n > 0)
      {
        fetch_meta(0);
        park_meta(0);
      }
    __syncthreads();

    // gather (vector_access_reduced.h:175-258): consecutive threads walk an entity's contiguous
    // DoF segment.  Thread tid owns elements tid + r * kThreads of EVERY cell of the batch, so the
    // table entry is decoded once per r and the (cell, r) slots unroll with compile-time cell
    // offsets: ~7 instructions per element, all loads of the batch in flight together.  Cells
    // missing from a ragged last batch carry invalid entity indices (fetch_meta) and gather zeros.
    // With kPipe the gather is software-pipelined: the loads of batch i+1 are issued
    // (gather_issue) before the scatter of batch i and land in registers while the scatter runs;
    // they are stored to the work rows (gather_store) once the scatter has released them.
    // Measured: Q2 +4 %, Q3 +1 %, Q4 +3 %, Q5 -2 %, Q8 -4 % (the extra live registers spill there).
    constexpr bool kPipe = BP4_PIPE_GATHER(P);
    constexpr int R = (G::DOF + kThreads - 1) / kThreads, S = CPB * R;
    // only the last r can run past the end of the cell
    const bool last_on = tid < G::DOF - (R - 1) * kThreads;
    auto       on      = [&](const int r) { return r < R - 1 || last_on; };
    double     gv[S];
    auto       gather_issue = [&](const int bf) {
      uint32_t tt[R], idx[S];
#pragma unroll
      for (int r = 0; r < R; ++r)
        tt[r] = sm.dtab[on(r) ? tid + r * kThreads : 0];
#pragma unroll
      for (int s1 = 0; s1 < S; ++s1)
        {
          const int      cell = s1 / R, r = s1 % R;
          const uint32_t base = sm.eidx[bf][cell][dtab_ent(tt[r])];
          idx[s1] = on(r) && base != 0xFFFFFFFFu ? base + dtab_rel(tt[r]) : 0xFFFFFFFFu;
        }
#pragma unroll
      for (int s1 = 0; s1 < S; ++s1)
        gv[s1] = idx[s1] != 0xFFFFFFFFu ? __ldg(a.src + idx[s1]) : 0.;
    };
    auto gather_store = [&]() {
      uint32_t tt[R];
#pragma unroll
      for (int r = 0; r < R; ++r)
        tt[r] = sm.dtab[on(r) ? tid + r * kThreads : 0];
#pragma unroll
      for (int s1 = 0; s1 < S; ++s1)
        {
          const int cell = s1 / R, r = s1 % R;
          if (on(r))
            sm.work[cell * G::WORK + dtab_off_work<P>(tt[r])] = gv[s1];
        }
    };

    for (int i = 0; i < my_n; ++i)
      {
        uint64_t  cell0;
        const int nc = batch_cells(i, cell0), bf = i & 1;
        BP4_TICK(0)
        if (!kPipe || i == 0)
          {
            gather_issue(bf);
            gather_store();
          }
        BP4_TICK(1)
        __syncthreads();
        BP4_TICK(2)
        if (i + 1 < my_n)
          fetch_meta(i + 1); // lands during the phases
        // phase 1/3 items are handed out from the LAST thread downwards: the ragged final round
        // of phase 2 lands on the first warps, so the two kinds of partial rounds end up on
        // different warps (= different SM sub-partitions) instead of piling up on warp 0
        if (!kFine)
          {
            for (int it = kThreads - 1 - tid; it < nc * G::ITEMS13; it += kThreads)
              phase1<P>(tb, sm.work + it * G::RW, sm.work + it * G::RW);
          }
        else
          {
            // one 1-D line per item, consecutive lanes on consecutive rows (odd row stride:
            // no bank conflicts); see phase1a
            const int n_rows = nc * G::ITEMS13;
            for (int it = tid; it < n_rows * G::N; it += kThreads)
              phase1a<P>(tb, sm.work + (it % n_rows) * G::RW, it / n_rows);
            __syncthreads();
            for (int it = tid; it < n_rows * Q; it += kThreads)
              phase1b<P>(tb, sm.work + (it % n_rows) * G::RW, it / n_rows);
            __syncthreads();
            for (int it = tid; it < n_rows * Q; it += kThreads)
              phase1c<P>(tb, sm.work + (it % n_rows) * G::RW, it / n_rows);
          }
        BP4_TICK(3)
        __syncthreads();
        BP4_TICK(2)
        for (int it = tid; it < nc * G::ITEMS2; it += kThreads)
          {
            const int cell = it / G::ITEMS2, r = it % G::ITEMS2;
            const int qz = r / Q, qx = r % Q;
            if constexpr (P >= BP4_P2_CALL_FROM)
              phase2_call<P>((uint32_t)((const unsigned char *)sm.coef[bf][cell] - smem_raw),
                             (uint32_t)((const unsigned char *)(sm.work + cell * G::WORK) - smem_raw), qx, qz,
                             sm.xq[qx], sm.xq[qz], sm.wq[qx] * sm.wq[qz]);
            else
              phase2<P>(tb, sm.coef[bf][cell], sm.work + cell * G::WORK, qx, qz, sm.xq[qx], sm.xq[qz],
                        sm.wq[qx] * sm.wq[qz]);
          }
        BP4_TICK(4)
        __syncthreads();
        BP4_TICK(2)
        if (!kFine)
          {
            for (int it = kThreads - 1 - tid; it < nc * G::ITEMS13; it += kThreads)
              phase3<P>(tb, sm.work + it * G::RW, sm.work + it * G::RW);
          }
        else
          {
            const int n_rows = nc * G::ITEMS13;
            for (int it = tid; it < n_rows * Q; it += kThreads)
              phase3a<P>(tb, sm.work + (it % n_rows) * G::RW, it / n_rows);
            __syncthreads();
            for (int it = tid; it < n_rows * Q; it += kThreads)
              phase3b<P>(tb, sm.work + (it % n_rows) * G::RW, it / n_rows);
            __syncthreads();
            for (int it = tid; it < n_rows * G::N; it += kThreads)
              phase3c<P>(tb, sm.work + (it % n_rows) * G::RW, it / n_rows);
          }
        if (i + 1 < my_n)
          park_meta(bf ^ 1);
        BP4_TICK(5)
        __syncthreads();
        BP4_TICK(2)
        // scatter-add (vector_access_reduced.h:437-521); the cell-interior entity (13) is
        // touched by this cell only -> plain store
        if (kPipe && i + 1 < my_n)
          gather_issue(bf ^ 1); // in flight during the scatter
        constexpr int SU = BP4_SU;
        uint32_t      tt[R];
#pragma unroll
        for (int r = 0; r < R; ++r)
          tt[r] = sm.dtab[on(r) ? tid + r * kThreads : 0];
#pragma unroll
        for (int s0 = 0; s0 < S; s0 += SU)
          {
            double   v[SU];
            uint32_t adr[SU];
#pragma unroll
            for (int u = 0; u < SU; ++u)
              if (s0 + u < S)
                {
                  const int      cell = (s0 + u) / R, r = (s0 + u) % R;
                  const uint32_t base = sm.eidx[bf][cell][dtab_ent(tt[r])];
                  adr[u] = on(r) && base != 0xFFFFFFFFu ? base + dtab_rel(tt[r]) : 0xFFFFFFFFu;
                  v[u]   = sm.work[cell * G::WORK + dtab_off_work<P>(tt[r])];
                }
#pragma unroll
            for (int u = 0; u < SU; ++u)
              if (s0 + u < S && adr[u] != 0xFFFFFFFFu)
                {
                  if (dtab_ent(tt[(s0 + u) % R]) == 13u)
                    a.dst[adr[u]] = v[u];
                  else
                    atomicAdd(a.dst + adr[u], v[u]);
                }
          }
        BP4_TICK(6)
        __syncthreads();
        BP4_TICK(2)
        if (kPipe && i + 1 < my_n)
          gather_store(); // the barrier after it is the one at the top of the next iteration
      }
    BP4_TICK_FLUSH
  }

  template <int P, int CPB>
  __global__ void __launch_bounds__(kThreads, kBlocksPerSM) cell_kernel_pf(const CellArgs a)
  {
    using G         = Geom<P>;
    constexpr int Q = G::Q;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PfSmem<P, CPB> &sm  = *reinterpret_cast<PfSmem<P, CPB> *>(smem_raw);
    const int       tid = threadIdx.x;
    const Tab<P>   &tb  = c_tab<P>;
    for (int i = tid; i < G::DOF; i += kThreads)
      sm.dtab[i] = a.dtab[i];
    if (tid < Q)
      {
        sm.xq[tid] = tb.xq[tid];
        sm.wq[tid] = tb.wq[tid];
      }
    const uint64_t n_batches = (a.n_cells + CPB - 1) / CPB;
    const int      my_n =
      n_batches > blockIdx.x ? (int)((n_batches - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    auto batch_cells = [&](const int i, uint64_t &cell0) {
      cell0 = ((uint64_t)blockIdx.x + (uint64_t)i * gridDim.x) * CPB;
      return (int)min((uint64_t)CPB, a.n_cells - cell0);
    };
    auto load_meta = [&](const int i) {
      uint64_t  cell0;
      const int nc = batch_cells(i, cell0), bf = i & 1;
      for (int k = tid; k < nc * 27; k += kThreads)
        sm.eidx[bf][k / 27][k % 27] = a.entity_index[cell0 * 27 + k];
      for (int k = tid; k < nc * 24; k += kThreads)
        sm.coef[bf][k / 24][k % 24] = a.coef[cell0 * 24 + k];
    };
    // asynchronous gather of batch i into sm.dofs (vector_access_reduced.h:175-258)
    auto issue_gather = [&](const int i) {
      uint64_t  cell0;
      const int nc = batch_cells(i, cell0), bf = i & 1;
      const int total = nc * G::DOF;
      for (int m = tid; m < total; m += kThreads)
        {
          const int      cell = m / G::DOF;
          const uint32_t t    = sm.dtab[m - cell * G::DOF];
          const uint32_t base = sm.eidx[bf][cell][dtab_ent(t)];
          const bool     ok   = base != 0xFFFFFFFFu;
          const double  *g    = a.src + (ok ? (size_t)base + dtab_rel(t) : 0);
          const uint32_t sa =
            (uint32_t)__cvta_generic_to_shared(sm.dofs + cell * G::DOFS + dtab_off<P>(t));
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa), "l"(g), "r"(ok ? 8 : 0)
                       : "memory");
        }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };

    if (my_n > 0)
      load_meta(0);
    __syncthreads();
    if (my_n > 0)
      issue_gather(0);

    for (int i = 0; i < my_n; ++i)
      {
        uint64_t  cell0;
        const int nc = batch_cells(i, cell0), bf = i & 1;
        if (i + 1 < my_n)
          load_meta(i + 1);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        for (int it = tid; it < nc * G::ITEMS13; it += kThreads)
          phase1<P>(tb, sm.dofs + it * G::RD, sm.work + it * G::RW);
        __syncthreads();
        if (i + 1 < my_n)
          issue_gather(i + 1);
        {
          // full rounds first; the ragged tail rotates over the warps from batch to batch
          const int n2 = nc * G::ITEMS2, full = (n2 / kThreads) * kThreads;
          const int rot = (tid + 32 * (i & 3)) & (kThreads - 1);
          for (int it = tid; it < full + kThreads; it += kThreads)
            {
              const int item = it < full ? it : full + rot;
              if (item < n2)
                {
                  const int cell = item / G::ITEMS2, r = item % G::ITEMS2;
                  const int qz = r / Q, qx = r % Q;
                  phase2<P>(tb, sm.coef[bf][cell], sm.work + cell * G::WORK, qx, qz, sm.xq[qx], sm.xq[qz],
                            sm.wq[qx] * sm.wq[qz]);
                }
            }
        }
        __syncthreads();
        for (int it = tid; it < nc * G::ITEMS13; it += kThreads)
          phase3<P>(tb, sm.work + it * G::RW, sm.work + it * G::RW);
        __syncthreads();
        // scatter-add (vector_access_reduced.h:437-521); interior entity: plain store
        const int total = nc * G::DOF;
        for (int m = tid; m < total; m += kThreads)
          {
            const int      cell = m / G::DOF;
            const uint32_t t    = sm.dtab[m - cell * G::DOF];
            const uint32_t ent  = dtab_ent(t);
            const uint32_t base = sm.eidx[bf][cell][ent];
            if (base != 0xFFFFFFFFu)
              {
                const double v = sm.work[cell * G::WORK + dtab_off_work<P>(t)];
                double      *p = a.dst + (size_t)base + dtab_rel(t);
                if (ent == 13u)
                  *p = v;
                else
                  atomicAdd(p, v);
              }
          }
        __syncthreads();
      }
  }

  template <int P, int CPB>
  __global__ void __launch_bounds__(kTrioThreads, kBlocksPerSM) cell_kernel_trio(const CellArgs a)
  {
    using G         = Geom<P>;
    constexpr int Q = G::Q;
    constexpr int T = kTrioThreads;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CellSmem<P, CPB> &sm  = *reinterpret_cast<CellSmem<P, CPB> *>(smem_raw);
    const int         tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const Tab<P>     &tb  = c_tab<P>;
    for (int i = tid; i < G::DOF; i += T)
      sm.dtab[i] = a.dtab[i];
    if (tid < Q)
      {
        sm.xq[tid] = tb.xq[tid];
        sm.wq[tid] = tb.wq[tid];
      }
    const uint64_t n_batches = (a.n_cells + CPB - 1) / CPB;
    for (uint64_t batch = blockIdx.x; batch < n_batches; batch += gridDim.x)
      {
        const uint64_t cell0 = batch * CPB;
        const int      nc    = (int)min((uint64_t)CPB, a.n_cells - cell0);
        for (int i = tid; i < nc * 27; i += T)
          sm.eidx[0][i / 27][i % 27] = a.entity_index[cell0 * 27 + i];
        for (int i = tid; i < nc * 24; i += T)
          sm.coef[0][i / 24][i % 24] = a.coef[cell0 * 24 + i];
        __syncthreads();
        constexpr int U     = 6;
        const int     total = nc * G::DOF;
        for (int m0 = tid; m0 < total; m0 += T * U)
          {
            double   v[U];
            uint32_t off[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
              {
                const int m = m0 + u * T;
                v[u]        = 0.;
                off[u]      = 0;
                if (m < total)
                  {
                    const int      cell = m / G::DOF;
                    const uint32_t t    = sm.dtab[m - cell * G::DOF];
                    const uint32_t base = sm.eidx[0][cell][dtab_ent(t)];
                    off[u]              = cell * G::WORK + dtab_off_work<P>(t);
                    if (base != 0xFFFFFFFFu)
                      v[u] = __ldg(a.src + (size_t)base + dtab_rel(t));
                  }
              }
#pragma unroll
            for (int u = 0; u < U; ++u)
              if (m0 + u * T < total)
                sm.work[off[u]] = v[u];
          }
        __syncthreads();
        for (int it = tid; it < nc * G::ITEMS13; it += T)
          phase1<P>(tb, sm.work + it * G::RW, sm.work + it * G::RW);
        __syncthreads();
        {
          // 10 trios (lines) per warp, lanes 30 and 31 idle; every warp runs the same number of
          // rounds so that the shuffles are always executed by the full trio mask
          const int n_lines = nc * G::ITEMS2;
          const int rounds  = (n_lines + 10 * (T / 32) - 1) / (10 * (T / 32));
          const int trio = lane / 3, c = lane - 3 * trio;
          for (int r = 0; r < rounds; ++r)
            {
              if (lane < 30)
                {
                  const int  line   = (r * (T / 32) + warp) * 10 + trio;
                  const bool active = line < n_lines;
                  const int  l      = active ? line : 0;
                  const int  cell = l / G::ITEMS2, rr = l % G::ITEMS2;
                  const int  qz = rr / Q, qx = rr % Q;
                  phase2_trio<P>(tb, sm.coef[0][cell], sm.work + cell * G::WORK, qx, qz, c, 3 * trio, 0x3FFFFFFFu,
                                 active, sm.xq[qx], sm.xq[qz], sm.wq[qx] * sm.wq[qz]);
                }
            }
        }
        __syncthreads();
        for (int it = tid; it < nc * G::ITEMS13; it += T)
          phase3<P>(tb, sm.work + it * G::RW, sm.work + it * G::RW);
        __syncthreads();
        for (int m = tid; m < total; m += T)
          {
            const int      cell = m / G::DOF;
            const uint32_t t    = sm.dtab[m - cell * G::DOF];
            const uint32_t ent  = dtab_ent(t);
            const uint32_t base = sm.eidx[0][cell][ent];
            if (base != 0xFFFFFFFFu)
              {
                const double v = sm.work[cell * G::WORK + dtab_off_work<P>(t)];
                double      *p = a.dst + (size_t)base + dtab_rel(t);
                if (ent == 13u)
                  *p = v;
                else
                  atomicAdd(p, v);
              }
          }
        __syncthreads();
      }
  }

  // ---- mbarrier / bulk-copy wrappers (PTX ISA: mbarrier, cp.async.bulk, cp.reduce.async.bulk) ----
  __device__ __forceinline__ void mbar_init(const uint32_t bar, const uint32_t count)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __device__ __forceinline__ void mbar_arrive(const uint32_t bar)
  {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
  }
  __device__ __forceinline__ void mbar_arrive_expect_tx(const uint32_t bar, const uint32_t bytes)
  {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  }
  __device__ __forceinline__ void mbar_wait(const uint32_t bar, const uint32_t parity)
  {
    asm volatile("{\n\t.reg .pred p;\n"
                 "WAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\t"
                 "bra WAIT_%=;\n"
                 "DONE_%=:\n\t}" ::"r"(bar),
                 "r"(parity)
                 : "memory");
  }
  __device__ __forceinline__ void bulk_load(const uint32_t smem_dst, const void *gmem_src, const uint32_t bytes,
                                            const uint32_t bar)
  {
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
                 "l"(gmem_src), "r"(bytes), "r"(bar)
                 : "memory");
  }
  __device__ __forceinline__ void bulk_reduce_add_f64(void *gmem_dst, const uint32_t smem_src, const uint32_t bytes)
  {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_src), "r"(bytes)
                 : "memory");
  }

  template <int P, int CPB>
  __global__ void __launch_bounds__(kThreads, kBlocksPerSM) cell_kernel_tma(const TmaArgs a)
  {
    using G          = Geom<P>;
    using St         = Stage<P>;
    constexpr int Q  = G::Q;
    constexpr int NN = G::N * G::N;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TmaSmem<P, CPB> &sm   = *reinterpret_cast<TmaSmem<P, CPB> *>(smem_raw);
    const int        tid  = threadIdx.x;
    const Tab<P>    &tb   = c_tab<P>;
    const uint32_t   mbar = (uint32_t)__cvta_generic_to_shared(&sm.mbar);
    for (int i = tid; i < NN * G::ROWS; i += kThreads)
      sm.itab[i] = a.itab[i];
    if (tid < 27)
      sm.slot[tid] = a.slot[tid];
    if (tid < Q)
      {
        sm.xq[tid] = tb.xq[tid];
        sm.wq[tid] = tb.wq[tid];
      }
    if (tid == 0)
      mbar_init(mbar, CPB * 27); // every (cell, entity) item arrives once per batch
    const uint64_t n_batches = (a.n_cells + CPB - 1) / CPB;
    const int      my_n =
      n_batches > blockIdx.x ? (int)((n_batches - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    auto batch_cells = [&](const int i, uint64_t &cell0) {
      cell0 = ((uint64_t)blockIdx.x + (uint64_t)i * gridDim.x) * CPB;
      return (int)min((uint64_t)CPB, a.n_cells - cell0);
    };
    auto load_meta = [&](const int i) {
      uint64_t  cell0;
      const int nc = batch_cells(i, cell0), bf = i & 1;
      for (int k = tid; k < nc * 27; k += kThreads)
        sm.eidx[bf][k / 27][k % 27] = a.entity_index[cell0 * 27 + k];
      for (int k = tid; k < nc * 24; k += kThreads)
        sm.coef[bf][k / 24][k % 24] = a.coef[cell0 * 24 + k];
    };
    // gather of batch i: one bulk copy per valid entity, zero-fill for Dirichlet entities
    auto issue_loads = [&](const int i) {
      uint64_t  cell0;
      const int nc = batch_cells(i, cell0), bf = i & 1;
      for (int k = tid; k < CPB * 27; k += kThreads)
        {
          const int cell = k / 27, e = k % 27;
          if (cell >= nc)
            {
              mbar_arrive(mbar);
              continue;
            }
          const uint32_t b    = sm.eidx[bf][cell][e];
          const uint32_t slot = sm.slot[e];
          const int      n    = St::n_dofs(e);
          double        *dst  = sm.stage_in + cell * St::SIZE + slot;
          if (b == 0xFFFFFFFFu)
            {
              sm.off[bf][cell][e] = (uint16_t)slot;
              for (int q = 0; q < n; ++q)
                dst[q] = 0.;
              mbar_arrive(mbar);
            }
          else
            {
              const uint32_t par  = b & 1u;
              sm.off[bf][cell][e] = (uint16_t)(slot + par);
              const uint32_t bytes = (uint32_t)St::r2(n + (int)par) * 8u;
              mbar_arrive_expect_tx(mbar, bytes);
              bulk_load((uint32_t)__cvta_generic_to_shared(dst), a.src + (b - par), bytes, mbar);
            }
        }
    };

    if (my_n > 0)
      load_meta(0);
    __syncthreads();
    if (my_n > 0)
      issue_loads(0);

    for (int i = 0; i < my_n; ++i)
      {
        uint64_t  cell0;
        const int nc = batch_cells(i, cell0), bf = i & 1;
        mbar_wait(mbar, (uint32_t)(i & 1));                               // stage_in(i) has landed
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // reduces of batch i-1 left stage_out
        __syncthreads();
        if (i + 1 < my_n)
          load_meta(i + 1);
        // pads of the output stage must add 0.0
        for (int k = tid; k < nc * 27; k += kThreads)
          {
            const int cell = k / 27, e = k % 27;
            double   *o    = sm.stage_out + cell * St::SIZE + sm.slot[e];
            const int n    = St::n_dofs(e);
            o[0]           = 0.;
            o[n]           = 0.;
            o[n + 1]       = 0.;
          }
        for (int it = tid; it < nc * G::ITEMS13; it += kThreads)
          {
            const int cell = it / G::ROWS, row = it % G::ROWS;
            phase1_io<P>(tb, StageIn{sm.stage_in + cell * St::SIZE, sm.itab, sm.off[bf][cell], row, G::ROWS},
                         sm.work + it * G::RW);
          }
        __syncthreads();
        if (i + 1 < my_n)
          issue_loads(i + 1);
        {
          const int n2 = nc * G::ITEMS2, full = (n2 / kThreads) * kThreads;
          const int rot = (tid + 32 * (i & 3)) & (kThreads - 1);
          for (int it = tid; it < full + kThreads; it += kThreads)
            {
              const int item = it < full ? it : full + rot;
              if (item < n2)
                {
                  const int cell = item / G::ITEMS2, r = item % G::ITEMS2;
                  const int qz = r / Q, qx = r % Q;
                  phase2<P>(tb, sm.coef[bf][cell], sm.work + cell * G::WORK, qx, qz, sm.xq[qx], sm.xq[qz],
                            sm.wq[qx] * sm.wq[qz]);
                }
            }
        }
        __syncthreads();
        for (int it = tid; it < nc * G::ITEMS13; it += kThreads)
          {
            const int cell = it / G::ROWS, row = it % G::ROWS;
            phase3_io<P>(tb, sm.work + it * G::RW,
                         StageOut{sm.stage_out + cell * St::SIZE, sm.itab, sm.off[bf][cell], row, G::ROWS});
          }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic writes -> async proxy reads
        __syncthreads();
        // scatter-add: one bulk reduction per valid entity
        for (int k = tid; k < nc * 27; k += kThreads)
          {
            const int      cell = k / 27, e = k % 27;
            const uint32_t b    = sm.eidx[bf][cell][e];
            if (b == 0xFFFFFFFFu)
              continue;
            const uint32_t par   = b & 1u;
            const uint32_t bytes = (uint32_t)St::r2(St::n_dofs(e) + (int)par) * 8u;
            bulk_reduce_add_f64(a.dst + (b - par),
                                (uint32_t)__cvta_generic_to_shared(sm.stage_out + cell * St::SIZE + sm.slot[e]),
                                bytes);
          }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  __device__ __forceinline__ double warp_sum(double v)
  {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }

  // block-reduce K partial sums and atomically add them to acc[0..K)
  template <int K>
  __device__ __forceinline__ void block_accumulate(double (&s)[K], double *acc)
  {
    __shared__ double red[K][32];
    const int         lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int         nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k)
      {
        s[k] = warp_sum(s[k]);
        if (lane == 0)
          red[k][warp] = s[k];
      }
    __syncthreads();
    if (warp == 0)
      {
#pragma unroll
        for (int k = 0; k < K; ++k)
          {
            double v = lane < nw ? red[k][lane] : 0.;
            v        = warp_sum(v);
            if (lane == 0)
              atomicAdd(acc + k, v);
          }
      }
  }

  // the seven merged sums of do_cg_update3b (solver_cg_optimized.h:37-44) for one entry
  __device__ __forceinline__ void post_terms(double (&s)[7], const double ri, const double di,
                                             const double hi, const double pr)
  {
    const double zi = pr * hi;
    s[0] += di * hi;
    s[1] += hi * hi;
    s[2] += ri * hi;
    s[3] += ri * ri;
    s[4] += ri * zi;
    s[5] += hi * zi;
    s[6] += ri * pr * ri;
  }

  // Fused merged kernel = LaplaceOperator::vmult_with_merged_sums (poisson_operator.h:327-377)
  // in ONE launch.  The reference runs do_cg_update4b on a DoF range "before its first
  // touch" and do_cg_update3b "after its last touch" of a sequential cell loop; thread blocks
  // have no such order, so
  //   pre : every cell recomputes r' = r + alpha h, p' = beta p - P r' for the DoFs it gathers
  //         from read-only old buffers; the entity's OWNER cell writes r', p' (to the ping-pong
  //         partners) and x (in place).
  //   post: cells scatter-add into h'; after a __threadfence each cell bumps one arrival
  //         counter per shared entity; the cell that completes an entity (last toucher) reads
  //         h', r', p' back through L2, accumulates the seven sums and zeroes the entity's
  //         slots in the old h buffer, which is next iteration's h'.  Cell-interior DoFs
  //         (touched once) are plain-stored and summed straight from shared memory.
  template <int P, int CPB>
  __global__ void __launch_bounds__(kThreads, kBlocksPerSM) cell_kernel_merged(const MergedArgs a)
  {
    using G = Geom<P>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CellSmem<P, CPB> &sm  = *reinterpret_cast<CellSmem<P, CPB> *>(smem_raw);
    const int         tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    load_tables<P, CPB>(sm, a.dtab);
    double     s[7]     = {0., 0., 0., 0., 0., 0., 0.};
    const bool first_it = a.alpha == 0.;

    const uint64_t n_batches = (a.n_cells + CPB - 1) / CPB;
    for (uint64_t batch = blockIdx.x; batch < n_batches; batch += gridDim.x)
      {
        const uint64_t cell0 = batch * CPB;
        const int      nc    = (int)min((uint64_t)CPB, a.n_cells - cell0);
        for (int i = tid; i < nc * 27; i += kThreads)
          {
            sm.eidx[0][i / 27][i % 27] = a.entity_index[cell0 * 27 + i];
            sm.meta[i / 27][i % 27] = a.meta[cell0 * 28 + (i / 27) * 28 + i % 27];
          }
        for (int i = tid; i < nc * 24; i += kThreads)
          sm.coef[0][i / 24][i % 24] = a.coef[cell0 * 24 + i];
        if (tid == 0)
          sm.n_chunks = 0;
        __syncthreads();

        // gather + do_cg_update4b (solver_cg_optimized.h:65-161)
        for (int cell = 0; cell < nc; ++cell)
          {
            const uint32_t *eidx = sm.eidx[0][cell];
            const uint8_t  *meta = sm.meta[cell];
            double         *dofs = sm.work + cell * G::WORK;
            for (int r0 = tid; r0 < G::DOF; r0 += kThreads * kGatherUnroll)
              {
                double   rv[kGatherUnroll], pv[kGatherUnroll], hv[kGatherUnroll], dv[kGatherUnroll];
                uint32_t t[kGatherUnroll], adr[kGatherUnroll];
                bool     own[kGatherUnroll];
#pragma unroll
                for (int u = 0; u < kGatherUnroll; ++u)
                  {
                    const int r = r0 + u * kThreads;
                    adr[u]      = 0xFFFFFFFFu;
                    own[u]      = false;
                    rv[u] = pv[u] = hv[u] = dv[u] = 0.;
                    if (r < G::DOF)
                      {
                        t[u]                = sm.dtab[r];
                        const uint32_t ent  = dtab_ent(t[u]);
                        const uint32_t base = eidx[ent];
                        if (base != 0xFFFFFFFFu)
                          {
                            adr[u] = base + dtab_rel(t[u]);
                            own[u] = (meta[ent] & kMetaOwner) != 0;
                            rv[u]  = a.r_old[adr[u]];
                            dv[u]  = a.prec[adr[u] / 3u];
                            if (!first_it)
                              {
                                pv[u] = a.p_old[adr[u]];
                                hv[u] = a.h_old[adr[u]];
                              }
                          }
                      }
                  }
#pragma unroll
                for (int u = 0; u < kGatherUnroll; ++u)
                  if (r0 + u * kThreads < G::DOF)
                    {
                      double pn = 0.;
                      if (adr[u] != 0xFFFFFFFFu)
                        {
                          const double pr = dv[u];
                          double       ri = rv[u];
                          if (own[u] && a.update_x)
                            a.x[adr[u]] += a.c1 * pv[u] + a.c2 * pr * ri;
                          if (first_it)
                            pn = -pr * ri;
                          else
                            {
                              ri += a.alpha * hv[u];
                              pn = a.beta * pv[u] - pr * ri;
                            }
                          if (own[u])
                            {
                              a.r_new[adr[u]] = ri;
                              a.p_new[adr[u]] = pn;
                            }
                        }
                      dofs[dtab_off_work<P>(t[u])] = pn;
                    }
              }
          }
        __syncthreads();

        apply_staged<P, CPB>(sm, nc);

        // scatter: shared entities through L2 atomics, the interior entity by plain store
        // with its do_cg_update3b terms taken on the spot
        for (int cell = 0; cell < nc; ++cell)
          {
            const uint32_t *eidx = sm.eidx[0][cell];
            const double   *dofs = sm.work + cell * G::WORK;
            for (int r0 = tid; r0 < G::DOF; r0 += kThreads * kGatherUnroll)
              {
                double rv[kGatherUnroll], pv[kGatherUnroll], dv[kGatherUnroll], hv[kGatherUnroll];
                bool   inner[kGatherUnroll];
#pragma unroll
                for (int u = 0; u < kGatherUnroll; ++u)
                  {
                    const int r = r0 + u * kThreads;
                    inner[u]    = false;
                    if (r < G::DOF)
                      {
                        const uint32_t t    = sm.dtab[r];
                        const uint32_t ent  = dtab_ent(t);
                        const uint32_t base = eidx[ent];
                        if (base != 0xFFFFFFFFu)
                          {
                            const double   v   = dofs[dtab_off_work<P>(t)];
                            const uint32_t adr = base + dtab_rel(t);
                            if (ent == 13u)
                              {
                                a.h_new[adr] = v;
                                inner[u]     = true;
                                hv[u]        = v;
                                rv[u]        = __ldcg(a.r_new + adr);
                                pv[u]        = __ldcg(a.p_new + adr);
                                dv[u]        = a.prec[adr / 3u];
                              }
                            else
                              atomicAdd(a.h_new + adr, v);
                          }
                      }
                  }
#pragma unroll
                for (int u = 0; u < kGatherUnroll; ++u)
                  if (inner[u])
                    post_terms(s, rv[u], pv[u], hv[u], dv[u]);
              }
          }
        __threadfence();
        __syncthreads();

        // arrival counters: one per shared entity, wrapping at the number of touching cells
        for (int i = tid; i < nc * 27; i += kThreads)
          {
            const int      cell = i / 27, ent = i % 27;
            const uint32_t base = sm.eidx[0][cell][ent];
            if (ent == 13 || base == 0xFFFFFFFFu)
              continue;
            const uint32_t nt   = (sm.meta[cell][ent] & 15u); // touching cells - 1
            bool           last = true;
            if (nt > 0)
              {
                last = atomicInc(a.counters + base / 3u, nt) == nt;
                if (last)
                  __threadfence();
              }
            if (last)
              {
                const int ex = ent % 3, ey = (ent / 3) % 3, ez = ent / 9;
                const int nd = 3 * (ex == 1 ? P - 1 : 1) * (ey == 1 ? P - 1 : 1) * (ez == 1 ? P - 1 : 1);
                const int nch  = (nd + 31) >> 5;
                const uint32_t slot = atomicAdd(&sm.n_chunks, (uint32_t)nch);
                for (int k = 0; k < nch; ++k)
                  {
                    sm.chunk_base[slot + k] = base + 32u * k;
                    sm.chunk_len[slot + k]  = (uint8_t)min(32, nd - 32 * k);
                  }
              }
          }
        __syncthreads();

        // do_cg_update3b (solver_cg_optimized.h:12-61) on the entities completed by this block
        const int n_chunks = (int)sm.n_chunks;
        for (int ch0 = warp; ch0 < n_chunks; ch0 += 4 * (kThreads / 32))
          {
            double rv[4], pv[4], hv[4], dv[4];
            bool   ok[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
              {
                const int ch = ch0 + u * (kThreads / 32);
                ok[u]        = ch < n_chunks && lane < (int)sm.chunk_len[ch < n_chunks ? ch : 0];
                if (ok[u])
                  {
                    const uint32_t adr = sm.chunk_base[ch] + lane;
                    hv[u]              = __ldcg(a.h_new + adr);
                    rv[u]              = __ldcg(a.r_new + adr);
                    pv[u]              = __ldcg(a.p_new + adr);
                    dv[u]              = a.prec[adr / 3u];
                    a.h_old[adr]       = 0.;
                  }
              }
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (ok[u])
                post_terms(s, rv[u], pv[u], hv[u], dv[u]);
          }
        __syncthreads();
      }
    block_accumulate<7>(s, a.acc);
  }

  // ---------------------------------------------------------------------------------------
  // warp-specialised cell kernel (plain and merged)
  // ---------------------------------------------------------------------------------------
  enum : int
  {
    kBarInFull  = 1, // memory -> compute : dofs of batch i gathered
    kBarInFree  = 2, // compute -> memory : phase 1 done, dofs may be overwritten
    kBarOutFull = 3, // compute -> memory : phase 3 results of batch i are in the work rows
    kBarOutFree = 4, // memory -> compute : results read, work rows may be overwritten
    kBarCompute = 5, // compute warpgroup only
    kBarMemory  = 6  // memory warpgroup only
  };
  __device__ __forceinline__ void bar_sync(const int id, const int n)
  {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
  }
  __device__ __forceinline__ void bar_arrive(const int id, const int n)
  {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
  }

  struct WsArgs
  {
    // mesh
    const uint32_t *entity_index;
    const double   *coef;
    const uint32_t *dtab;
    uint64_t        n_cells;
    // plain: dst += A src
    const double *src;
    double       *dst;
    // merged (see MergedArgs)
    const uint8_t *meta;
    uint32_t      *counters;
    const double  *r_old, *p_old;
    double        *h_old, *r_new, *p_new, *h_new, *x;
    const double  *prec;
    double         alpha, beta, c1, c2;
    int            update_x;
    double        *acc;
  };

  template <int P, int CPB, bool MERGED>
  __global__ void __launch_bounds__(kWsThreads, kBlocksPerSM) cell_kernel_ws(const WsArgs a)
  {
    using G         = Geom<P>;
    constexpr int Q = G::Q;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WsSmem<P, CPB> &sm  = *reinterpret_cast<WsSmem<P, CPB> *>(smem_raw);
    const int       tid = threadIdx.x;
    for (int i = tid; i < G::DOF; i += kWsThreads)
      sm.dtab[i] = a.dtab[i];
    if (tid < Q)
      {
        sm.xq[tid] = c_tab<P>.xq[tid];
        sm.wq[tid] = c_tab<P>.wq[tid];
      }
    __syncthreads();

    const uint64_t n_batches = (a.n_cells + CPB - 1) / CPB;
    const int      my_n =
      n_batches > blockIdx.x ? (int)((n_batches - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    auto batch_cells = [&](const int i, uint64_t &cell0) {
      cell0 = ((uint64_t)blockIdx.x + (uint64_t)i * gridDim.x) * CPB;
      return (int)min((uint64_t)CPB, a.n_cells - cell0);
    };

    if (tid >= 128)
      {
        // =========================== memory warpgroup ===================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kWsRegsMem));
        const int  t = tid - 128, lane = t & 31, warp = t >> 5;
        double     s[7]     = {0., 0., 0., 0., 0., 0., 0.};
        const bool first_it = a.alpha == 0.;
        constexpr int U     = MERGED ? 2 : 8;
        for (int i = 0; i <= my_n; ++i)
          {
            if (i < my_n)
              {
                if (i > 0)
                  bar_sync(kBarInFree, kWsThreads);
                uint64_t  cell0;
                const int nc = batch_cells(i, cell0);
                const int bf = i & 1;
                for (int k = t; k < nc * 27; k += 128)
                  {
                    sm.eidx[bf][k / 27][k % 27] = a.entity_index[cell0 * 27 + k];
                    if (MERGED)
                      sm.meta[bf][k / 27][k % 27] = a.meta[cell0 * 28 + (k / 27) * 28 + k % 27];
                  }
                for (int k = t; k < nc * 24; k += 128)
                  sm.coef[bf][k / 24][k % 24] = a.coef[cell0 * 24 + k];
                bar_sync(kBarMemory, 128);
                // gather (+ do_cg_update4b, solver_cg_optimized.h:65-161)
                const int total = nc * G::DOF;
                for (int m0 = t; m0 < total; m0 += 128 * U)
                  {
                    if (!MERGED)
                      {
                        // asynchronous copies (LDGSTS): no registers held, every load of the
                        // batch in flight at once; src-size 0 zero-fills Dirichlet entities
#pragma unroll
                        for (int u = 0; u < U; ++u)
                          {
                            const int m = m0 + u * 128;
                            if (m < total)
                              {
                                const int      cell = m / G::DOF;
                                const uint32_t tb   = sm.dtab[m - cell * G::DOF];
                                const uint32_t base = sm.eidx[bf][cell][dtab_ent(tb)];
                                const bool     ok   = base != 0xFFFFFFFFu;
                                const double  *g    = a.src + (ok ? (size_t)base + dtab_rel(tb) : 0);
                                const uint32_t sa   = (uint32_t)__cvta_generic_to_shared(
                                  sm.dofs + cell * G::DOFS + dtab_off<P>(tb));
                                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa), "l"(g),
                                             "r"(ok ? 8 : 0)
                                             : "memory");
                              }
                          }
                      }
                    else
                      {
                        double   rv[U], pv[U], hv[U], dv[U];
                        uint32_t off[U], adr[U];
                        bool     own[U];
#pragma unroll
                        for (int u = 0; u < U; ++u)
                          {
                            const int m = m0 + u * 128;
                            adr[u]      = 0xFFFFFFFFu;
                            own[u]      = false;
                            off[u]      = 0;
                            rv[u] = pv[u] = hv[u] = dv[u] = 0.;
                            if (m < total)
                              {
                                const int      cell = m / G::DOF;
                                const uint32_t tb   = sm.dtab[m - cell * G::DOF];
                                const uint32_t ent  = dtab_ent(tb);
                                const uint32_t base = sm.eidx[bf][cell][ent];
                                off[u]              = cell * G::DOFS + dtab_off<P>(tb);
                                if (base != 0xFFFFFFFFu)
                                  {
                                    adr[u] = base + dtab_rel(tb);
                                    own[u] = (sm.meta[bf][cell][ent] & kMetaOwner) != 0;
                                    rv[u]  = a.r_old[adr[u]];
                                    dv[u]  = a.prec[adr[u] / 3u];
                                    if (!first_it)
                                      {
                                        pv[u] = a.p_old[adr[u]];
                                        hv[u] = a.h_old[adr[u]];
                                      }
                                  }
                              }
                          }
#pragma unroll
                        for (int u = 0; u < U; ++u)
                          if (m0 + u * 128 < total)
                            {
                              double pn = 0.;
                              if (adr[u] != 0xFFFFFFFFu)
                                {
                                  const double pr = dv[u];
                                  double       ri = rv[u];
                                  if (own[u] && a.update_x)
                                    a.x[adr[u]] += a.c1 * pv[u] + a.c2 * pr * ri;
                                  if (first_it)
                                    pn = -pr * ri;
                                  else
                                    {
                                      ri += a.alpha * hv[u];
                                      pn = a.beta * pv[u] - pr * ri;
                                    }
                                  if (own[u])
                                    {
                                      a.r_new[adr[u]] = ri;
                                      a.p_new[adr[u]] = pn;
                                    }
                                }
                              sm.dofs[off[u]] = pn;
                            }
                      }
                  }
                if (!MERGED)
                  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
                __syncwarp();
                bar_arrive(kBarInFull, kWsThreads);
              }
            if (i > 0)
              {
                bar_sync(kBarOutFull, kWsThreads);
                uint64_t  cell0;
                const int nc    = batch_cells(i - 1, cell0);
                const int bf    = (i - 1) & 1;
                const int total = nc * G::DOF;
                // scatter-add (vector_access_reduced.h:437-521); cell-interior entity: plain store
                for (int m0 = t; m0 < total; m0 += 128 * 4)
                  {
                    double rv[4], pv[4], dv[4], hv[4];
                    bool   inner[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                      {
                        const int m = m0 + u * 128;
                        inner[u]    = false;
                        if (m < total)
                          {
                            const int      cell = m / G::DOF;
                            const uint32_t tb   = sm.dtab[m - cell * G::DOF];
                            const uint32_t ent  = dtab_ent(tb);
                            const uint32_t base = sm.eidx[bf][cell][ent];
                            if (base != 0xFFFFFFFFu)
                              {
                                const double   v   = sm.work[cell * G::WORK + dtab_off_work<P>(tb)];
                                const uint32_t adr = base + dtab_rel(tb);
                                double        *dst = (MERGED ? a.h_new : a.dst) + adr;
                                if (ent == 13u)
                                  {
                                    *dst = v;
                                    if (MERGED)
                                      {
                                        inner[u] = true;
                                        hv[u]    = v;
                                        rv[u]    = __ldcg(a.r_new + adr);
                                        pv[u]    = __ldcg(a.p_new + adr);
                                        dv[u]    = a.prec[adr / 3u];
                                      }
                                  }
                                else
                                  atomicAdd(dst, v);
                              }
                          }
                      }
                    if (MERGED)
                      {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                          if (inner[u])
                            post_terms(s, rv[u], pv[u], hv[u], dv[u]);
                      }
                  }
                __syncwarp();
                if (i < my_n)
                  bar_arrive(kBarOutFree, kWsThreads);
                if (MERGED)
                  {
                    // last-toucher protocol, see cell_kernel_merged
                    __threadfence();
                    if (t == 0)
                      sm.n_chunks = 0;
                    bar_sync(kBarMemory, 128);
                    for (int k = t; k < nc * 27; k += 128)
                      {
                        const int      cell = k / 27, ent = k % 27;
                        const uint32_t base = sm.eidx[bf][cell][ent];
                        if (ent == 13 || base == 0xFFFFFFFFu)
                          continue;
                        const uint32_t nt   = (sm.meta[bf][cell][ent] & 15u);
                        bool           last = true;
                        if (nt > 0)
                          {
                            last = atomicInc(a.counters + base / 3u, nt) == nt;
                            if (last)
                              __threadfence();
                          }
                        if (last)
                          {
                            const int ex = ent % 3, ey = (ent / 3) % 3, ez = ent / 9;
                            const int nd =
                              3 * (ex == 1 ? P - 1 : 1) * (ey == 1 ? P - 1 : 1) * (ez == 1 ? P - 1 : 1);
                            const int      nch  = (nd + 31) >> 5;
                            const uint32_t slot = atomicAdd(&sm.n_chunks, (uint32_t)nch);
                            for (int q = 0; q < nch; ++q)
                              {
                                sm.chunk_base[slot + q] = base + 32u * q;
                                sm.chunk_len[slot + q]  = (uint8_t)min(32, nd - 32 * q);
                              }
                          }
                      }
                    bar_sync(kBarMemory, 128);
                    const int n_chunks = (int)sm.n_chunks;
                    for (int ch0 = warp; ch0 < n_chunks; ch0 += 4 * 4)
                      {
                        double rv[4], pv[4], hv[4], dv[4];
                        bool   ok[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                          {
                            const int ch = ch0 + u * 4;
                            ok[u]        = ch < n_chunks && lane < (int)sm.chunk_len[ch < n_chunks ? ch : 0];
                            if (ok[u])
                              {
                                const uint32_t adr = sm.chunk_base[ch] + lane;
                                hv[u]              = __ldcg(a.h_new + adr);
                                rv[u]              = __ldcg(a.r_new + adr);
                                pv[u]              = __ldcg(a.p_new + adr);
                                dv[u]              = a.prec[adr / 3u];
                                a.h_old[adr]       = 0.;
                              }
                          }
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                          if (ok[u])
                            post_terms(s, rv[u], pv[u], hv[u], dv[u]);
                      }
                    bar_sync(kBarMemory, 128); // chunk list is rewritten by the next batch
                  }
              }
          }
        if (MERGED)
          {
#pragma unroll
            for (int k = 0; k < 7; ++k)
              {
                s[k] = warp_sum(s[k]);
                if (lane == 0)
                  sm.red[k][warp] = s[k];
              }
            bar_sync(kBarMemory, 128);
            if (t < 7)
              atomicAdd(a.acc + t, sm.red[t][0] + sm.red[t][1] + sm.red[t][2] + sm.red[t][3]);
          }
      }
    else
      {
        // =========================== compute warpgroup ==================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kWsRegsComp));
        const Tab<P> &tb = c_tab<P>;
        for (int i = 0; i < my_n; ++i)
          {
            uint64_t  cell0;
            const int nc = batch_cells(i, cell0);
            const int bf = i & 1;
            bar_sync(kBarInFull, kWsThreads);
            if (i > 0)
              bar_sync(kBarOutFree, kWsThreads);
            // item -> thread maps are rotated by one warp per phase and batch so that the
            // partially filled last round does not always land on the same warps (each warp
            // owns one SM sub-partition's FP64 pipe)
            const int r1 = (tid + 32 * (i & 3)) & 127, r2 = (tid + 32 * ((i + 1) & 3)) & 127,
                      r3 = (tid + 32 * ((i + 2) & 3)) & 127;
            for (int it = r1; it < nc * G::ITEMS13; it += 128)
              phase1<P>(tb, sm.dofs + it * G::RD, sm.work + it * G::RW);
            __syncwarp();
            if (i + 1 < my_n)
              bar_arrive(kBarInFree, kWsThreads);
            bar_sync(kBarCompute, 128);
            {
              // full rounds first; the ragged tail goes to the rotated thread map
              const int n2 = nc * G::ITEMS2, full = (n2 / 128) * 128;
              for (int it = tid; it < full + 128; it += 128)
                {
                  const int item = it < full ? it : full + r2;
                  if (item < n2)
                    {
                      const int cell = item / G::ITEMS2, r = item % G::ITEMS2;
                      const int qz = r / Q, qx = r % Q;
                      phase2<P>(tb, sm.coef[bf][cell], sm.work + cell * G::WORK, qx, qz, sm.xq[qx],
                                sm.xq[qz], sm.wq[qx] * sm.wq[qz]);
                    }
                }
            }
            bar_sync(kBarCompute, 128);
            for (int it = r3; it < nc * G::ITEMS13; it += 128)
              phase3<P>(tb, sm.work + it * G::RW, sm.work + it * G::RW);
            __syncwarp();
            bar_arrive(kBarOutFull, kWsThreads);
          }
      }
  }

  // entity meta data of the fused kernel, built on the device from the entity table alone:
  // touch[first node of entity] = number of local cells holding it, owner = lowest such cell
  __global__ void __launch_bounds__(256) meta_count_kernel(const uint64_t n_cells,
                                                           const uint32_t *__restrict__ entity_index,
                                                           uint32_t *touch, uint32_t *owner)
  {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cells * 27)
      return;
    const uint32_t base = entity_index[i];
    if (base == 0xFFFFFFFFu)
      return;
    atomicAdd(touch + base / 3u, 1u);
    atomicMin(owner + base / 3u, (uint32_t)(i / 27));
  }

  __global__ void __launch_bounds__(256) meta_fill_kernel(const uint64_t n_cells,
                                                          const uint32_t *__restrict__ entity_index,
                                                          const uint32_t *__restrict__ touch,
                                                          const uint32_t *__restrict__ owner,
                                                          uint8_t *meta)
  {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cells * 27)
      return;
    const uint64_t cell = i / 27;
    const int      ent  = (int)(i % 27);
    const uint32_t base = entity_index[i];
    uint8_t        m    = 0;
    if (base != 0xFFFFFFFFu)
      {
        m = (uint8_t)((touch[base / 3u] - 1u) & 15u);
        if (owner[base / 3u] == (uint32_t)cell)
          m |= kMetaOwner;
      }
    meta[cell * 28 + ent] = m;
  }

  // ---------------------------------------------------------------------------------------
  // streaming kernels
  // ---------------------------------------------------------------------------------------
  // do_cg_update4b<3,double,true>, solver_cg_optimized.h:65-161, over [0,n)
  __global__ void __launch_bounds__(256) pre_kernel(const uint64_t n, double *__restrict__ h,
                                                    double *__restrict__ x, double *__restrict__ r,
                                                    double *__restrict__ p,
                                                    const double *__restrict__ prec,
                                                    const double alpha, const double beta,
                                                    const double alpha_old, const double beta_old)
  {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const double   c1 = alpha_old != 0. ? alpha + alpha_old / beta_old : 0.;
    const double   c2 = alpha_old != 0. ? alpha_old / beta_old : 0.;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
      {
        const double pr = prec[i / 3];
        if (alpha == 0.)
          p[i] = -pr * r[i];
        else
          {
            double       ri = r[i];
            const double pi = p[i];
            if (alpha_old != 0.)
              x[i] += c1 * pi + c2 * pr * ri;
            ri += alpha * h[i];
            r[i] = ri;
            p[i] = beta * pi - pr * ri;
          }
        h[i] = 0.;
      }
  }

  // do_cg_update3b<3,double>, solver_cg_optimized.h:12-61, over [0,n) -> acc[7]
  __global__ void __launch_bounds__(256) post_kernel(const uint64_t n, const double *__restrict__ r,
                                                     const double *__restrict__ d,
                                                     const double *__restrict__ h,
                                                     const double *__restrict__ prec, double *acc)
  {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    double         s[7]   = {0., 0., 0., 0., 0., 0., 0.};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
      {
        const double pr = prec[i / 3], ri = r[i], di = d[i], hi = h[i];
        const double zi = pr * hi;
        s[0] += di * hi;
        s[1] += hi * hi;
        s[2] += ri * hi;
        s[3] += ri * ri;
        s[4] += ri * zi;
        s[5] += hi * zi;
        s[6] += ri * pr * ri;
      }
    block_accumulate<7>(s, acc);
  }

  __global__ void __launch_bounds__(256) fixup_kernel(const uint64_t n, const uint32_t *__restrict__ con,
                                                      double *__restrict__ dst,
                                                      const double *__restrict__ src)
  {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
      dst[con[i]] = src[con[i]];
  }

  // dst = s*dst + a*src   (s == 0: dst = a*src without reading dst)
  __global__ void __launch_bounds__(256) sadd_kernel(const uint64_t n, double *__restrict__ dst,
                                                     const double s, const double a,
                                                     const double *__restrict__ src)
  {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
      dst[i] = (s == 0. ? 0. : s * dst[i]) + a * src[i];
  }

  __global__ void __launch_bounds__(256) dot_kernel(const uint64_t n, const double *__restrict__ a,
                                                    const double *__restrict__ b, double *acc)
  {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    double         s[1]   = {0.};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
      s[0] += a[i] * b[i];
    block_accumulate<1>(s, acc);
  }

  // g += a*h ; acc += g.w   (w may alias g)
  __global__ void __launch_bounds__(256) add_and_dot_kernel(const uint64_t n, double *g, const double a,
                                                            const double *__restrict__ h,
                                                            const double *w, double *acc)
  {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    double         s[1]   = {0.};
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
      {
        const double gi = g[i] + a * h[i];
        const double wi = (w == g) ? gi : w[i];
        g[i]            = gi;
        s[0] += gi * wi;
      }
    block_accumulate<1>(s, acc);
  }

  __global__ void __launch_bounds__(256) nonzero_kernel(const uint64_t n, const double *__restrict__ v,
                                                        int *flag)
  {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    int            nz     = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
      nz |= (v[i] != 0.);
    if (__any_sync(0xffffffffu, nz) && (threadIdx.x & 31) == 0)
      atomicOr(flag, 1);
  }

  // ghost exchange helpers: buf[k] = v[idx[k]] and v[idx[k]] += buf[k]
  __global__ void __launch_bounds__(256) pack_kernel(const uint64_t n, const uint32_t *__restrict__ idx,
                                                     const double *__restrict__ v, double *__restrict__ buf)
  {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
      buf[i] = v[idx[i]];
  }
  __global__ void __launch_bounds__(256) unpack_add_kernel(const uint64_t n, const uint32_t *__restrict__ idx,
                                                           const double *__restrict__ buf, double *v)
  {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
      v[idx[i]] += buf[i]; // export lists of different peers may repeat an index: launched per peer
  }
  // out[i] = in[3 i]
  __global__ void __launch_bounds__(256) stride3_kernel(const uint64_t n, const double *__restrict__ in,
                                                        double *__restrict__ out)
  {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
      out[i] = in[3 * i];
  }

  // dst[3i+c] = diag[i]*src[3i+c]   (diagonal_matrix_blocked.h:21-26)
  __global__ void __launch_bounds__(256) jacobi_kernel(const uint64_t n, double *__restrict__ dst,
                                                       const double *__restrict__ src,
                                                       const double *__restrict__ diag)
  {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
      dst[i] = diag[i / 3] * src[i];
  }

  // x += c1*d + c2*prec*g   (solver_cg_optimized.h:260-288)
  __global__ void __launch_bounds__(256) xfinal_kernel(const uint64_t n, double *__restrict__ x,
                                                       const double *__restrict__ d,
                                                       const double *__restrict__ g,
                                                       const double *__restrict__ prec, const double c1,
                                                       const double c2)
  {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
      x[i] += c1 * d[i] + c2 * prec[i / 3] * g[i];
  }

  // inverse diagonal of the scalar GLL(p+1) Laplacian (poisson_operator.h:392-426): one
  // thread per (cell, node); collocation makes the unit-vector gradient non-zero only on
  // the three grid lines through the node.  gll holds x[N], w[N], Dg[N][N] (Dg[i][q]=l_i'(x_q)).
  __device__ __forceinline__ void metric_at(const double *cf, double x, double y, double z, double w,
                                            double &g00, double &g01, double &g02, double &g11,
                                            double &g12, double &g22)
  {
    double r0[3], r1[3], r2[3];
#pragma unroll
    for (int d = 0; d < 3; ++d)
      {
        const double v1 = cf[3 + d], v3 = cf[6 + d], v4 = cf[9 + d], v9 = cf[12 + d],
                     v10 = cf[15 + d], v12 = cf[18 + d], v13 = cf[21 + d];
        r0[d] = (v1 + z * v10) + y * (v4 + z * v13);
        r1[d] = (v3 + z * v12) + x * (v4 + z * v13);
        r2[d] = (v9 + y * v12) + x * (v10 + y * v13);
      }
    double k0[3], k1[3], k2[3];
    k0[0] = r1[1] * r2[2] - r1[2] * r2[1];
    k0[1] = r1[2] * r2[0] - r1[0] * r2[2];
    k0[2] = r1[0] * r2[1] - r1[1] * r2[0];
    k1[0] = r2[1] * r0[2] - r2[2] * r0[1];
    k1[1] = r2[2] * r0[0] - r2[0] * r0[2];
    k1[2] = r2[0] * r0[1] - r2[1] * r0[0];
    k2[0] = r0[1] * r1[2] - r0[2] * r1[1];
    k2[1] = r0[2] * r1[0] - r0[0] * r1[2];
    k2[2] = r0[0] * r1[1] - r0[1] * r1[0];
    const double det = r0[0] * k0[0] + r0[1] * k0[1] + r0[2] * k0[2];
    const double sc  = w / det;
    g00 = sc * (k0[0] * k0[0] + k0[1] * k0[1] + k0[2] * k0[2]);
    g01 = sc * (k0[0] * k1[0] + k0[1] * k1[1] + k0[2] * k1[2]);
    g02 = sc * (k0[0] * k2[0] + k0[1] * k2[1] + k0[2] * k2[2]);
    g11 = sc * (k1[0] * k1[0] + k1[1] * k1[1] + k1[2] * k1[2]);
    g12 = sc * (k1[0] * k2[0] + k1[1] * k2[1] + k1[2] * k2[2]);
    g22 = sc * (k2[0] * k2[0] + k2[1] * k2[1] + k2[2] * k2[2]);
  }

  __global__ void __launch_bounds__(128) diag_kernel(const int p, const uint64_t n_cells,
                                                     const uint32_t *__restrict__ entity_index,
                                                     const double *__restrict__ coef,
                                                     const double *__restrict__ gll, double *diag,
                                                     const int stride)
  {
    const int      N  = p + 1, N3 = N * N * N;
    const uint64_t id = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_cells * N3)
      return;
    const uint64_t cell = id / N3;
    const int      l = (int)(id % N3), i = l % N, j = (l / N) % N, k = l / (N * N);
    const int      ex = i == 0 ? 0 : (i == p ? 2 : 1), ey = j == 0 ? 0 : (j == p ? 2 : 1),
              ez = k == 0 ? 0 : (k == p ? 2 : 1);
    const uint32_t base = entity_index[cell * 27 + ex + 3 * ey + 9 * ez];
    if (base == 0xFFFFFFFFu)
      return;
    const int oi = ex == 1 ? i - 1 : 0, oj = ey == 1 ? j - 1 : 0, ok = ez == 1 ? k - 1 : 0;
    const int sx = ex == 1 ? p - 1 : 1, sy = ey == 1 ? p - 1 : 1;
    const int pos = oi + sx * (oj + sy * ok);
    const double *x = gll, *w = gll + N, *Dg = gll + 2 * N;
    const double *cf = coef + cell * 24;
    double        s = 0., g00, g01, g02, g11, g12, g22;
    for (int q = 0; q < N; ++q)
      {
        const double dx = Dg[i * N + q], dy = Dg[j * N + q], dz = Dg[k * N + q];
        metric_at(cf, x[q], x[j], x[k], w[q] * w[j] * w[k], g00, g01, g02, g11, g12, g22);
        s += dx * dx * g00;
        metric_at(cf, x[i], x[q], x[k], w[i] * w[q] * w[k], g00, g01, g02, g11, g12, g22);
        s += dy * dy * g11;
        metric_at(cf, x[i], x[j], x[q], w[i] * w[j] * w[q], g00, g01, g02, g11, g12, g22);
        s += dz * dz * g22;
      }
    metric_at(cf, x[i], x[j], x[k], w[i] * w[j] * w[k], g00, g01, g02, g11, g12, g22);
    const double di = Dg[i * N + i], dj = Dg[j * N + j], dk = Dg[k * N + k];
    s += 2. * (di * dj * g01 + di * dk * g02 + dj * dk * g12);
    atomicAdd(diag + ((size_t)base / 3 + pos) * stride, s);
  }

  __global__ void __launch_bounds__(256) invert_diag_kernel(const uint64_t n, double *d)
  {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
      d[i] = d[i] == 0. ? 1. : 1. / d[i];
  }

  // ---------------------------------------------------------------------------------------
  // launchers
  // ---------------------------------------------------------------------------------------
  static inline int stream_grid(uint64_t n, int sms)
  {
    const uint64_t blocks = (n + 255) / 256;
    return (int)std::min<uint64_t>(std::max<uint64_t>(blocks, 1), (uint64_t)sms * 8);
  }

  template <int P>
  static cudaError_t init_degree(std::vector<uint32_t> &walk)
  {
    Tab<P> tb;
    fill_tab<P>(tb);
    walk.resize(Geom<P>::DOF);
    build_dof_table<P>(walk.data());
    cudaError_t e = cudaMemcpyToSymbol(c_tab<P>, &tb, sizeof(tb));
    if (e != cudaSuccess)
      return e;
    constexpr int CPB = Cfg<P>::CPB;
    e = cudaFuncSetAttribute(cell_kernel_plain<P, CPB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(CellSmem<P, CPB>));
    if (e != cudaSuccess)
      return e;
    e = cudaFuncSetAttribute(cell_kernel_merged<P, CPB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(CellSmem<P, CPB>));
    if (e != cudaSuccess)
      return e;
    e = cudaFuncSetAttribute(cell_kernel_trio<P, TrioCfg<P>::CPB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(CellSmem<P, TrioCfg<P>::CPB>));
    if (e != cudaSuccess)
      return e;
    e = cudaFuncSetAttribute(cell_kernel_pf<P, PfCfg<P>::CPB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(PfSmem<P, PfCfg<P>::CPB>));
    if (e != cudaSuccess)
      return e;
    e = cudaFuncSetAttribute(cell_kernel_tma<P, TmaCfg<P>::CPB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(TmaSmem<P, TmaCfg<P>::CPB>));
    if (e != cudaSuccess)
      return e;
    constexpr int WCPB = WsCfg<P>::CPB;
    e = cudaFuncSetAttribute(cell_kernel_ws<P, WCPB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sizeof(WsSmem<P, WCPB>));
    if (e != cudaSuccess)
      return e;
    return cudaFuncSetAttribute(cell_kernel_ws<P, WCPB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(WsSmem<P, WCPB>));
  }

  template <int P>
  static cudaError_t run_cell_ws(const WsArgs &a, bool merged, int sms, cudaStream_t st)
  {
    constexpr int  CPB       = WsCfg<P>::CPB;
    const uint64_t n_batches = (a.n_cells + CPB - 1) / CPB;
    const int      grid      = (int)std::min<uint64_t>(n_batches, (uint64_t)sms * kBlocksPerSM);
    if (grid == 0)
      return cudaSuccess;
    if (merged)
      cell_kernel_ws<P, CPB, true><<<grid, kWsThreads, sizeof(WsSmem<P, CPB>), st>>>(a);
    else
      cell_kernel_ws<P, CPB, false><<<grid, kWsThreads, sizeof(WsSmem<P, CPB>), st>>>(a);
    return cudaGetLastError();
  }

  template <int P>
  static cudaError_t run_cell_plain(const CellArgs &a, int sms, cudaStream_t st)
  {
    constexpr int  CPB       = Cfg<P>::CPB;
    const uint64_t n_batches = (a.n_cells + CPB - 1) / CPB;
    const int      grid      = (int)std::min<uint64_t>(n_batches, (uint64_t)sms * Cfg<P>::BLOCKS);
    if (grid == 0)
      return cudaSuccess;
    cell_kernel_plain<P, CPB><<<grid, kThreads, sizeof(CellSmem<P, CPB>), st>>>(a);
#ifdef BP4_PHASE_TIMING
    if (getenv("BP4_PHASE_TIMING"))
      {
        cudaStreamSynchronize(st);
        unsigned long long h[8], z[8] = {0};
        cudaMemcpyFromSymbol(h, g_phase_clk, sizeof(h));
        cudaMemcpyToSymbol(g_phase_clk, z, sizeof(z));
        const double tot = double(h[0] + h[1] + h[2] + h[3] + h[4] + h[5] + h[6]);
        fprintf(stderr, "phase clk share: meta %.1f%% gather %.1f%% barrier %.1f%% P1 %.1f%% P2 %.1f%% P3 %.1f%% scatter %.1f%% | per-warp total %.0f clk\n",
                100 * h[0] / tot, 100 * h[1] / tot, 100 * h[2] / tot, 100 * h[3] / tot, 100 * h[4] / tot,
                100 * h[5] / tot, 100 * h[6] / tot, tot / (grid * (kThreads / 32)));
      }
#endif
    return cudaGetLastError();
  }

  template <int P>
  static cudaError_t run_cell_trio(const CellArgs &a, int sms, cudaStream_t st)
  {
    constexpr int  CPB       = TrioCfg<P>::CPB;
    const uint64_t n_batches = (a.n_cells + CPB - 1) / CPB;
    const int      grid      = (int)std::min<uint64_t>(n_batches, (uint64_t)sms * kBlocksPerSM);
    if (grid == 0)
      return cudaSuccess;
    cell_kernel_trio<P, CPB><<<grid, kTrioThreads, sizeof(CellSmem<P, CPB>), st>>>(a);
    return cudaGetLastError();
  }

  template <int P>
  static cudaError_t run_cell_pf(const CellArgs &a, int sms, cudaStream_t st)
  {
    constexpr int  CPB       = PfCfg<P>::CPB;
    const uint64_t n_batches = (a.n_cells + CPB - 1) / CPB;
    const int      grid      = (int)std::min<uint64_t>(n_batches, (uint64_t)sms * kBlocksPerSM);
    if (grid == 0)
      return cudaSuccess;
    cell_kernel_pf<P, CPB><<<grid, kThreads, sizeof(PfSmem<P, CPB>), st>>>(a);
    return cudaGetLastError();
  }

  template <int P>
  static cudaError_t run_cell_tma(const TmaArgs &a, int sms, cudaStream_t st)
  {
    constexpr int  CPB       = TmaCfg<P>::CPB;
    const uint64_t n_batches = (a.n_cells + CPB - 1) / CPB;
    const int      grid      = (int)std::min<uint64_t>(n_batches, (uint64_t)sms * kBlocksPerSM);
    if (grid == 0)
      return cudaSuccess;
    cell_kernel_tma<P, CPB><<<grid, kThreads, sizeof(TmaSmem<P, CPB>), st>>>(a);
    return cudaGetLastError();
  }

  template <int P>
  static void stage_tables(std::vector<uint16_t> &out)
  {
    out.assign(28 + Geom<P>::N * Geom<P>::N * Geom<P>::ROWS, 0);
    build_stage_tables<P>(out.data(), out.data() + 28);
  }

  template <int P>
  static cudaError_t run_cell_merged(const MergedArgs &a, int sms, cudaStream_t st)
  {
    constexpr int  CPB       = Cfg<P>::CPB;
    const uint64_t n_batches = (a.n_cells + CPB - 1) / CPB;
    const int      grid      = (int)std::min<uint64_t>(n_batches, (uint64_t)sms * kBlocksPerSM);
    if (grid == 0)
      return cudaSuccess;
    cell_kernel_merged<P, CPB><<<grid, kThreads, sizeof(CellSmem<P, CPB>), st>>>(a);
    return cudaGetLastError();
  }

#define BP4_DISPATCH(p, CALL)                    \
  switch (p)                                     \
    {                                            \
      case 2: { constexpr int P = 2; CALL; } break; \
      case 3: { constexpr int P = 3; CALL; } break; \
      case 4: { constexpr int P = 4; CALL; } break; \
      case 5: { constexpr int P = 5; CALL; } break; \
      case 6: { constexpr int P = 6; CALL; } break; \
      case 7: { constexpr int P = 7; CALL; } break; \
      case 8: { constexpr int P = 8; CALL; } break; \
      default: return cudaErrorInvalidValue;     \
    }

  cudaError_t launch_init_degree(int degree, std::vector<uint32_t> &walk)
  {
    BP4_DISPATCH(degree, return init_degree<P>(walk));
    return cudaSuccess;
  }

  cudaError_t launch_cell_plain(int degree, const CellArgs &a, int sms, cudaStream_t st)
  {
    BP4_DISPATCH(degree, return run_cell_plain<P>(a, sms, st));
    return cudaSuccess;
  }

  cudaError_t launch_cell_tma(int degree, const TmaArgs &a, int sms, cudaStream_t st)
  {
    BP4_DISPATCH(degree, return run_cell_tma<P>(a, sms, st));
    return cudaSuccess;
  }

  // [0,28): slot table, [28, ...): inverse table of the TMA variant
  cudaError_t launch_stage_tables(int degree, std::vector<uint16_t> &out)
  {
    BP4_DISPATCH(degree, stage_tables<P>(out));
    return cudaSuccess;
  }

  cudaError_t launch_cell_trio(int degree, const CellArgs &a, int sms, cudaStream_t st)
  {
    BP4_DISPATCH(degree, return run_cell_trio<P>(a, sms, st));
    return cudaSuccess;
  }

  cudaError_t launch_cell_pf(int degree, const CellArgs &a, int sms, cudaStream_t st)
  {
    BP4_DISPATCH(degree, return run_cell_pf<P>(a, sms, st));
    return cudaSuccess;
  }

  cudaError_t launch_cell_ws(int degree, const MergedArgs *m, const CellArgs *p, int sms, cudaStream_t st)
  {
    WsArgs a{};
    if (m)
      {
        a.entity_index = m->entity_index, a.coef = m->coef, a.dtab = m->dtab, a.n_cells = m->n_cells;
        a.meta = m->meta, a.counters = m->counters, a.r_old = m->r_old, a.p_old = m->p_old;
        a.h_old = m->h_old, a.r_new = m->r_new, a.p_new = m->p_new, a.h_new = m->h_new, a.x = m->x;
        a.prec = m->prec, a.alpha = m->alpha, a.beta = m->beta, a.c1 = m->c1, a.c2 = m->c2;
        a.update_x = m->update_x, a.acc = m->acc;
      }
    else
      {
        a.entity_index = p->entity_index, a.coef = p->coef, a.dtab = p->dtab, a.n_cells = p->n_cells;
        a.src = p->src, a.dst = p->dst;
      }
    BP4_DISPATCH(degree, return run_cell_ws<P>(a, m != nullptr, sms, st));
    return cudaSuccess;
  }

  cudaError_t launch_cell_merged(int degree, const MergedArgs &a, int sms, cudaStream_t st)
  {
    BP4_DISPATCH(degree, return run_cell_merged<P>(a, sms, st));
    return cudaSuccess;
  }

  // touch/owner: scratch arrays of n_nodes entries; touch is left ZEROED for use as the
  // arrival counters of the fused kernel
  cudaError_t launch_build_meta(uint64_t n_cells, uint64_t n_nodes, const uint32_t *entity_index,
                                uint32_t *touch, uint32_t *owner, uint8_t *meta, cudaStream_t st)
  {
    cudaError_t e = cudaMemsetAsync(touch, 0, sizeof(uint32_t) * n_nodes, st);
    if (e != cudaSuccess)
      return e;
    e = cudaMemsetAsync(owner, 0xFF, sizeof(uint32_t) * n_nodes, st);
    if (e != cudaSuccess)
      return e;
    const uint64_t n = n_cells * 27;
    if (n)
      {
        meta_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n_cells, entity_index, touch, owner);
        meta_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n_cells, entity_index, touch, owner,
                                                                      meta);
      }
    e = cudaGetLastError();
    if (e != cudaSuccess)
      return e;
    return cudaMemsetAsync(touch, 0, sizeof(uint32_t) * n_nodes, st);
  }

  int cells_per_block(int degree)
  {
    switch (degree)
      {
        case 2: return Cfg<2>::CPB;
        case 3: return Cfg<3>::CPB;
        case 4: return Cfg<4>::CPB;
        case 5: return Cfg<5>::CPB;
        case 6: return Cfg<6>::CPB;
        case 7: return Cfg<7>::CPB;
        case 8: return Cfg<8>::CPB;
      }
    return 0;
  }

  cudaError_t launch_pre(uint64_t n, double *h, double *x, double *r, double *p, const double *prec,
                         double alpha, double beta, double alpha_old, double beta_old, int sms,
                         cudaStream_t st)
  {
    if (n == 0)
      return cudaSuccess;
    pre_kernel<<<stream_grid(n, sms), 256, 0, st>>>(n, h, x, r, p, prec, alpha, beta, alpha_old,
                                                     beta_old);
    return cudaGetLastError();
  }

  cudaError_t launch_post(uint64_t n, const double *r, const double *d, const double *h,
                          const double *prec, double *acc, int sms, cudaStream_t st)
  {
    if (n == 0)
      return cudaSuccess;
    post_kernel<<<stream_grid(n, sms), 256, 0, st>>>(n, r, d, h, prec, acc);
    return cudaGetLastError();
  }

  cudaError_t launch_fixup(uint64_t n, const uint32_t *con, double *dst, const double *src,
                           cudaStream_t st)
  {
    if (n == 0)
      return cudaSuccess;
    fixup_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, con, dst, src);
    return cudaGetLastError();
  }

  cudaError_t launch_sadd(uint64_t n, double *dst, double s, double a, const double *src, int sms,
                          cudaStream_t st)
  {
    if (n == 0)
      return cudaSuccess;
    sadd_kernel<<<stream_grid(n, sms), 256, 0, st>>>(n, dst, s, a, src);
    return cudaGetLastError();
  }

  cudaError_t launch_dot(uint64_t n, const double *a, const double *b, double *acc, int sms,
                         cudaStream_t st)
  {
    if (n == 0)
      return cudaSuccess;
    dot_kernel<<<stream_grid(n, sms), 256, 0, st>>>(n, a, b, acc);
    return cudaGetLastError();
  }

  cudaError_t launch_add_and_dot(uint64_t n, double *g, double a, const double *h, const double *w,
                                 double *acc, int sms, cudaStream_t st)
  {
    if (n == 0)
      return cudaSuccess;
    add_and_dot_kernel<<<stream_grid(n, sms), 256, 0, st>>>(n, g, a, h, w, acc);
    return cudaGetLastError();
  }

  cudaError_t launch_nonzero(uint64_t n, const double *v, int *flag, int sms, cudaStream_t st)
  {
    if (n == 0)
      return cudaSuccess;
    nonzero_kernel<<<stream_grid(n, sms), 256, 0, st>>>(n, v, flag);
    return cudaGetLastError();
  }

  cudaError_t launch_jacobi(uint64_t n, double *dst, const double *src, const double *diag, int sms,
                            cudaStream_t st)
  {
    if (n == 0)
      return cudaSuccess;
    jacobi_kernel<<<stream_grid(n, sms), 256, 0, st>>>(n, dst, src, diag);
    return cudaGetLastError();
  }

  cudaError_t launch_xfinal(uint64_t n, double *x, const double *d, const double *g, const double *prec,
                            double c1, double c2, int sms, cudaStream_t st)
  {
    if (n == 0)
      return cudaSuccess;
    xfinal_kernel<<<stream_grid(n, sms), 256, 0, st>>>(n, x, d, g, prec, c1, c2);
    return cudaGetLastError();
  }

  // assemble the scalar GLL diagonal: entry of node i goes to diag[i * stride]
  cudaError_t launch_diag_assemble(int degree, uint64_t n_cells, const uint32_t *entity_index,
                                   const double *coef, const double *gll, double *diag, int stride,
                                   cudaStream_t st)
  {
    const int      N3 = (degree + 1) * (degree + 1) * (degree + 1);
    const uint64_t n  = n_cells * N3;
    if (n > 0)
      diag_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(degree, n_cells, entity_index, coef, gll,
                                                                diag, stride);
    return cudaGetLastError();
  }

  cudaError_t launch_diag_invert(uint64_t n_nodes, double *diag, cudaStream_t st)
  {
    if (n_nodes > 0)
      invert_diag_kernel<<<(unsigned)((n_nodes + 255) / 256), 256, 0, st>>>(n_nodes, diag);
    return cudaGetLastError();
  }

  cudaError_t launch_pack(uint64_t n, const uint32_t *idx, const double *v, double *buf, cudaStream_t st)
  {
    if (n > 0)
      pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, idx, v, buf);
    return cudaGetLastError();
  }

  cudaError_t launch_unpack_add(uint64_t n, const uint32_t *idx, const double *buf, double *v, cudaStream_t st)
  {
    if (n > 0)
      unpack_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, idx, buf, v);
    return cudaGetLastError();
  }

  cudaError_t launch_stride3(uint64_t n, const double *in, double *out, cudaStream_t st)
  {
    if (n > 0)
      stride3_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, in, out);
    return cudaGetLastError();
  }

  // host table for the diagonal kernel: x[N], w[N], Dg[N][N]
  void gll_table(int degree, std::vector<double> &out)
  {
    const int           N = degree + 1;
    std::vector<double> x, w;
    gauss_lobatto_01(N, x, w);
    out.assign(2 * N + N * N, 0.);
    for (int i = 0; i < N; ++i)
      {
        out[i]     = x[i];
        out[N + i] = w[i];
        for (int q = 0; q < N; ++q)
          out[2 * N + i * N + q] = lagrange_deriv(x, i, x[q]);
      }
  }
} // namespace bp4
