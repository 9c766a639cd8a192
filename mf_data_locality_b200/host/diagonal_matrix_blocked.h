// diagonal_matrix_blocked.h -- Jacobi preconditioner with one diagonal entry per node applied
// to all `dim` vector components; mirror of the reference's DiagonalMatrixBlocked
// (diagonal_matrix_blocked.h:6-36), executed by bp4_jacobi_vmult on the device.
#pragma once
#include "device_vector.h"

template <int dim, typename Number>
class DiagonalMatrixBlocked
{
public:
  using VectorType = dealii::LinearAlgebra::distributed::Vector<Number>;

  void vmult(VectorType &dst, const VectorType &src) const
  {
    if (dst.size() != dim * diagonal.size())
      throw std::runtime_error("Dimension mismatch " + std::to_string(dst.size()) + " vs " +
                               std::to_string(dim) + " x " + std::to_string(diagonal.size()));
    dealii::bp4_check(bp4_jacobi_vmult(dst.context(), dst.handle(), src.handle(), diagonal.handle()));
  }

  const VectorType &get_vector() const { return diagonal; }

  VectorType diagonal;
};
