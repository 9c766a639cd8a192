// benchmark_precond_merged: BP4 with the fully merged CG (SolverCGFullMerge: vector updates,
// Jacobi scaling and the seven dot products fused into the operator loop) -- plugin for the
// driver in host/benchmark.h, the counterpart of the reference's
// benchmark_precond_merged/bench.cc:4-25.
#include "../host/benchmark.h"

template <typename Operator, typename Preconditioner>
unsigned int run_cg_solver(const Operator &A, LinearAlgebra::distributed::Vector<double> &x,
                           const LinearAlgebra::distributed::Vector<double> &b, const Preconditioner &P)
{
  const SolverSettings &st = solver_settings(); // ReductionControl(100, 1e-15, 1e-8) by default
  ReductionControl      control(st.max_steps, st.abs_tol, st.rel_tol);
  try
    {
      SolverCGFullMerge<LinearAlgebra::distributed::Vector<double>>(control).solve(A, x, b, P);
    }
  catch (const SolverControl::NoConvergence &)
    {}
  return control.last_step();
}

#ifndef BP4_NO_MAIN
int main(int argc, char **argv)
{
  run(argc, argv);
  return 0;
}
#endif
