// Jacobi preconditioner of the BP4 benchmark: ONE inverse-diagonal entry per lattice node,
// applied to all `dim` vector components of that node.  Host-side mirror of the reference class
// of the same name (diagonal_matrix_blocked.h:6-36): same template parameters, same members
// (vmult, get_vector, diagonal); the scaling itself runs on the device (bp4_jacobi_vmult).
#pragma once
#include <string>

#include "device_vector.h"

template <int dim, typename Number>
class DiagonalMatrixBlocked
{
public:
  using VectorType = dealii::LinearAlgebra::distributed::Vector<Number>;

  // n_nodes inverse-diagonal values, filled by LaplaceOperator::compute_inverse_diagonal
  VectorType diagonal;

  const VectorType &get_vector() const { return diagonal; }

  // dst[dim * i + c] = diagonal[i] * src[dim * i + c]
  void vmult(VectorType &dst, const VectorType &src) const
  {
    const auto n_vec = dst.size(), n_diag = diagonal.size();
    if (n_vec != dim * n_diag) // the reference's AssertThrow, same wording
      throw std::runtime_error("Dimension mismatch " + std::to_string(n_vec) + " vs " + std::to_string(dim) +
                               " x " + std::to_string(n_diag));
    dealii::bp4_check(bp4_jacobi_vmult(dst.context(), dst.handle(), src.handle(), diagonal.handle()));
  }
};
