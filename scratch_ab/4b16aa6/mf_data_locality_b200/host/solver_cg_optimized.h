// solver_cg_optimized.h -- host-side mirror of the reference's SolverCGFullMerge
// (solver_cg_optimized.h:165-303).  The two fused vector kernels of that file,
// do_cg_update4b (:65-161) and do_cg_update3b (:12-61), live on the device inside
// bp4_vmult_merged (csrc/bp4_kernels.cu); what remains here is the scalar recurrence that
// rebuilds alpha, beta and the residual norm from the seven sums and updates x only every
// second iteration.
#pragma once
#include <cmath>

#include "device_vector.h"
#include "solver_control.h"

template <typename VectorType>
class SolverCGFullMerge : public dealii::SolverBase<VectorType>
{
public:
  using size_type = dealii::types::global_dof_index;

  explicit SolverCGFullMerge(dealii::SolverControl &cn) : dealii::SolverBase<VectorType>(cn) {}
  virtual ~SolverCGFullMerge() = default;

  template <typename MatrixType, typename PreconditionerType>
  void solve(const MatrixType &A, VectorType &x, const VectorType &b, const PreconditionerType &preconditioner)
  {
    using number = typename VectorType::value_type;
    dealii::SolverControl::State conv = dealii::SolverControl::iterate;
    VectorType                   g, d, h; // VectorMemory pointers in the reference (:201-203)
    g.reinit(x, true);
    d.reinit(x, true);
    h.reinit(x, true);

    // residual; a zero start vector skips the operator (:221-227)
    if (!x.all_zero())
      {
        A.vmult(g, x);
        g.add(-1., b);
      }
    else
      g.equ(-1., b);
    double res_norm = g.l2_norm();
    conv            = this->iteration_status(0, res_norm, x);
    if (conv != dealii::SolverControl::iterate)
      return;

    number alpha = 0., beta = 0., alpha_old = 0., beta_old = 0.;
    int    it = 0;
    while (conv == dealii::SolverControl::iterate)
      {
        ++it;
        // x is brought up to date only on odd iterations, two search directions at once
        const auto s = A.vmult_with_merged_sums(x, g, d, h, preconditioner, alpha, beta,
                                                it % 2 == 1 ? alpha_old : number(0.), beta_old);
        alpha_old = alpha;
        beta_old  = beta;
        if (s[0] == 0.)
          throw std::runtime_error("SolverCGFullMerge: division by zero");
        alpha    = s[6] / s[0];                                              // r.Pr / d.Ad
        res_norm = std::sqrt(s[3] + 2 * alpha * s[2] + alpha * alpha * s[1]); // |r + alpha h|
        conv     = this->iteration_status(it, res_norm, x);
        if (conv != dealii::SolverControl::iterate)
          {
            // the pending update(s) of x (:256-288)
            if (it % 2 == 1)
              x.add(alpha, d);
            else
              dealii::bp4_check(bp4_x_finalize_even(x.context(), x.handle(), d.handle(), g.handle(),
                                                    preconditioner.get_vector().handle(),
                                                    alpha + alpha_old / beta_old, alpha_old / beta_old));
            break;
          }
        // r_{k+1}^T P r_{k+1} without another sweep (:292-295)
        beta = alpha * (s[4] + alpha * s[5]) / s[6];
      }
    if (conv != dealii::SolverControl::success)
      throw dealii::SolverControl::NoConvergence(it, res_norm);
  }
};
