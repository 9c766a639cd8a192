// benchmark_precond_merged: BP4 with the fully merged CG (SolverCGFullMerge: vector updates,
// Jacobi scaling and the seven dot products fused into the operator loop).  This file is the
// `run_cg_solver` plugin the driver in host/benchmark.h calls -- the counterpart of the
// reference's benchmark_precond_merged/bench.cc:4-25.
#include "../host/benchmark.h"

using DeviceVector = LinearAlgebra::distributed::Vector<double>;

template <typename Operator, typename Preconditioner>
unsigned int run_cg_solver(const Operator &A, DeviceVector &x, const DeviceVector &b, const Preconditioner &P)
{
  return solve_and_count<SolverCGFullMerge<DeviceVector>>(A, x, b, P); // ReductionControl(100, 1e-15, 1e-8)
}

#ifndef BP4_NO_MAIN
int main(int argc, char **argv)
{
  run(argc, argv); // bench <degree> [s] [compact_output]
}
#endif
