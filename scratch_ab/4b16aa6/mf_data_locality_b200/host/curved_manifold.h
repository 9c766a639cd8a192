// curved_manifold.h -- the chart that deforms the benchmark mesh: mirror of the part of the
// reference's MyManifold that defines vertex positions (curved_manifold.h:25-35).
// pull_back (Newton) and warmup_code() are not needed: vertices are created directly from
// lattice points and there are no CPU cores to spin up.
#pragma once
#include <cmath>

#include "dealii_standin.h"

template <int dim>
class MyManifold
{
public:
  MyManifold() : factor(0.1) {}

  // x -> x + factor * prod_d sin(pi x_d), added to every coordinate
  dealii::Point3 push_forward(const dealii::Point3 &p) const
  {
    double sinval = factor;
    for (unsigned int d = 0; d < dim; ++d)
      sinval *= std::sin(dealii::numbers::PI * p[d]);
    dealii::Point3 out;
    for (unsigned int d = 0; d < dim; ++d)
      out[d] = p[d] + sinval;
    return out;
  }

private:
  const double factor;
};
