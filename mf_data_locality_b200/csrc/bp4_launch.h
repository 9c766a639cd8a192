// bp4_launch.h -- launcher prototypes shared by bp4_kernels.cu and bp4_capi.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "bp4_kernels.cuh"

namespace bp4
{
  cudaError_t launch_init_degree(int degree, std::vector<uint32_t> &walk);
  int         cells_per_block(int degree);
  cudaError_t launch_cell_plain(int degree, const CellArgs &a, int sms, cudaStream_t st);
  cudaError_t launch_cell_trio(int degree, const CellArgs &a, int sms, cudaStream_t st);
  cudaError_t launch_cell_pf(int degree, const CellArgs &a, int sms, cudaStream_t st);
  cudaError_t launch_cell_tma(int degree, const TmaArgs &a, int sms, cudaStream_t st);
  cudaError_t launch_stage_tables(int degree, std::vector<uint16_t> &out);
  // warp-specialised kernel: pass exactly one of m (merged) / p (plain)
  cudaError_t launch_cell_ws(int degree, const MergedArgs *m, const CellArgs *p, int sms, cudaStream_t st);
  cudaError_t launch_cell_merged(int degree, const MergedArgs &a, int sms, cudaStream_t st);
  cudaError_t launch_build_meta(uint64_t n_cells, uint64_t n_nodes, const uint32_t *entity_index,
                                uint32_t *touch, uint32_t *owner, uint8_t *meta, cudaStream_t st);
  cudaError_t launch_pre(uint64_t n, double *h, double *x, double *r, double *p, const double *prec,
                         double alpha, double beta, double alpha_old, double beta_old, int sms,
                         cudaStream_t st);
  cudaError_t launch_post(uint64_t n, const double *r, const double *d, const double *h,
                          const double *prec, double *acc, int sms, cudaStream_t st);
  cudaError_t launch_fixup(uint64_t n, const uint32_t *con, double *dst, const double *src,
                           cudaStream_t st);
  cudaError_t launch_sadd(uint64_t n, double *dst, double s, double a, const double *src, int sms,
                          cudaStream_t st);
  cudaError_t launch_dot(uint64_t n, const double *a, const double *b, double *acc, int sms,
                         cudaStream_t st);
  cudaError_t launch_add_and_dot(uint64_t n, double *g, double a, const double *h, const double *w,
                                 double *acc, int sms, cudaStream_t st);
  cudaError_t launch_nonzero(uint64_t n, const double *v, int *flag, int sms, cudaStream_t st);
  cudaError_t launch_jacobi(uint64_t n, double *dst, const double *src, const double *diag, int sms,
                            cudaStream_t st);
  cudaError_t launch_xfinal(uint64_t n, double *x, const double *d, const double *g, const double *prec,
                            double c1, double c2, int sms, cudaStream_t st);
  cudaError_t launch_diag_assemble(int degree, uint64_t n_cells, const uint32_t *entity_index,
                                   const double *coef, const double *gll, double *diag, int stride,
                                   cudaStream_t st);
  cudaError_t launch_diag_invert(uint64_t n_nodes, double *diag, cudaStream_t st);
  cudaError_t launch_pack(uint64_t n, const uint32_t *idx, const double *v, double *buf, cudaStream_t st);
  cudaError_t launch_unpack_add(uint64_t n, const uint32_t *idx, const double *buf, double *v, cudaStream_t st);
  cudaError_t launch_stride3(uint64_t n, const double *in, double *out, cudaStream_t st);
  void        gll_table(int degree, std::vector<double> &out);
} // namespace bp4
