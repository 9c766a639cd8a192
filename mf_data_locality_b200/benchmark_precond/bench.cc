// benchmark_precond: BP4 with the stock preconditioned CG (deal.II SolverCG) and the blocked
// Jacobi preconditioner -- plugin for the driver in host/benchmark.h, the counterpart of the
// reference's benchmark_precond/bench.cc:4-25.
#include "../host/benchmark.h"

template <typename Operator, typename Preconditioner>
unsigned int run_cg_solver(const Operator &A, LinearAlgebra::distributed::Vector<double> &x,
                           const LinearAlgebra::distributed::Vector<double> &b, const Preconditioner &P)
{
  // at most 100 iterations, absolute 1e-15, relative 1e-8; running out of iterations is not
  // an error for the benchmark, it just reports 100
  const SolverSettings &st = solver_settings(); // ReductionControl(100, 1e-15, 1e-8) by default
  ReductionControl      control(st.max_steps, st.abs_tol, st.rel_tol);
  try
    {
      SolverCG<LinearAlgebra::distributed::Vector<double>>(control).solve(A, x, b, P);
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
