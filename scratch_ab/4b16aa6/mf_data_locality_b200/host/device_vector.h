// device_vector.h -- stand-in for LinearAlgebra::distributed::Vector<double> whose storage
// lives in B200 HBM behind the C ABI (bp4_vec).  Offers the members the reference's solvers
// call (solver_cg_optimized.h:215-228, :257; deal.II SolverCG: operator*, add_and_dot, sadd).
#pragma once
#include <bp4.h>

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace dealii
{
  inline void bp4_check(int code)
  {
    if (code != 0)
      throw std::runtime_error(std::string("bp4: ") + bp4_last_error());
  }

  namespace LinearAlgebra
  {
    namespace distributed
    {
      template <typename Number>
      class Vector
      {
        static_assert(sizeof(Number) == sizeof(double), "FP64 only");

      public:
        using value_type = Number;
        Vector() = default;
        Vector(const Vector &) = delete;
        Vector &operator=(const Vector &) = delete;
        ~Vector() { release(); }

        // n_local entries are owned, the vector additionally carries n_ghost ghost slots
        void reinit(bp4_ctx *ctx_, std::uint64_t n_local_, std::uint64_t n_ghost_ = 0,
                    std::uint64_t n_global_ = 0, const bool omit_zeroing_entries = false)
        {
          release();
          ctx      = ctx_;
          n_local  = n_local_;
          n_ghost  = n_ghost_;
          n_global = n_global_ ? n_global_ : n_local_;
          if (omit_zeroing_entries)
            bp4_check(bp4_vec_alloc_uninitialized(ctx, n_local + n_ghost, &h));
          else
            bp4_check(bp4_vec_alloc(ctx, n_local + n_ghost, &h));
        }
        // omit_zeroing_entries = true leaves the entries undefined, as in deal.II (the solvers
        // take their temporaries this way, solver_cg_optimized.h:215-217)
        void reinit(const Vector &other, const bool omit_zeroing_entries = false)
        {
          reinit(other.ctx, other.n_local, other.n_ghost, other.n_global, omit_zeroing_entries);
        }
        Vector &operator=(const Number s)
        {
          if (s != Number())
            throw std::runtime_error("only assignment of zero is supported");
          bp4_check(bp4_vec_set_zero(ctx, h));
          return *this;
        }
        std::uint64_t size() const { return n_global; }
        std::uint64_t local_size() const { return n_local; }

        void   equ(const Number a, const Vector &v) { bp4_check(bp4_equ(ctx, h, a, v.h)); }
        void   add(const Number a, const Vector &v) { bp4_check(bp4_add(ctx, h, a, v.h)); }
        void   sadd(const Number s, const Number a, const Vector &v) { bp4_check(bp4_sadd(ctx, h, s, a, v.h)); }
        Number operator*(const Vector &v) const
        {
          double r;
          bp4_check(bp4_dot(ctx, h, v.h, &r));
          return r;
        }
        Number add_and_dot(const Number a, const Vector &v, const Vector &w)
        {
          double r;
          bp4_check(bp4_add_and_dot(ctx, h, a, v.h, w.h, &r));
          return r;
        }
        Number l2_norm() const
        {
          double r;
          bp4_check(bp4_l2_norm(ctx, h, &r));
          return r;
        }
        bool all_zero() const
        {
          int r;
          bp4_check(bp4_all_zero(ctx, h, &r));
          return r != 0;
        }
        void upload(const Number *host, std::uint64_t n) { bp4_check(bp4_vec_upload(ctx, h, host, n)); }
        void download(Number *host, std::uint64_t n) const { bp4_check(bp4_vec_download(ctx, h, host, n)); }

        bp4_vec *handle() const { return h; }
        bp4_ctx *context() const { return ctx; }

      private:
        void release()
        {
          if (h)
            bp4_vec_free(ctx, h);
          h = nullptr;
        }
        bp4_ctx      *ctx = nullptr;
        bp4_vec      *h   = nullptr;
        std::uint64_t n_local = 0, n_ghost = 0, n_global = 0;
      };
    } // namespace distributed
  }   // namespace LinearAlgebra
} // namespace dealii
