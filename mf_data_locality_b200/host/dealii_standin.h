// dealii_standin.h -- the handful of deal.II roles the BP4 hot path of mf_data_locality
// consumes, re-created for the one mesh family the benchmark uses (a refined, deformed
// box; benchmark.h:66-102).  These are NOT deal.II re-implementations: each class offers
// just the members the reference's own headers call, so that the host-side mirror
// (poisson_operator.h, renumber_dofs_for_mf.h, solver_cg_optimized.h, benchmark.h in this
// directory) reads like the reference and a maintainer can swap real deal.II back in.
//
// deal.II behaviours that are not visible in the reference tree (cell traversal order,
// SIMD batch width, cell-batch ranges, p4est partition, ownership of interface DoFs) are
// explicit parameters (MatrixFree::AdditionalData, Triangulation), see SURVEY App. B.
//
// Every rank process holds the whole (structured, cheap) lattice description and derives
// ownership geometrically, so no host-side message passing is needed during setup.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <functional>
#include <memory>
#include <numeric>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#define AssertThrow(cond, msg)         \
  do                                   \
    {                                  \
      if (!(cond))                     \
        throw std::runtime_error(msg); \
    }                                  \
  while (0)

namespace dealii
{
  // std::vector that leaves trivially constructible elements uninitialised on resize(): the big
  // per-node arrays are written completely by parallel passes right after they are sized, so the
  // pages are first touched by the worker threads instead of being zeroed by one
  template <typename T>
  struct default_init_allocator : std::allocator<T>
  {
    template <typename U>
    struct rebind
    {
      using other = default_init_allocator<U>;
    };
    template <typename U, typename... Args>
    void construct(U *ptr, Args &&...args)
    {
      if constexpr (sizeof...(Args) == 0)
        ::new (static_cast<void *>(ptr)) U;
      else
        ::new (static_cast<void *>(ptr)) U(std::forward<Args>(args)...);
    }
  };
  template <typename T>
  using raw_vector = std::vector<T, default_init_allocator<T>>;

  // threads one parallel_chunks call may use; callers that already run several workers lower it
  inline unsigned int &parallel_cap()
  {
    thread_local unsigned int cap = 16;
    return cap;
  }

  // f(begin, end) on disjoint chunks of [0, n), one host thread per chunk (set-up loops over cells)
  template <typename F>
  inline void parallel_chunks(const std::uint64_t n, F &&f, const std::uint64_t grain = 4096)
  {
    const unsigned int nt = (unsigned int)std::max<std::uint64_t>(
      1, std::min<std::uint64_t>({(std::uint64_t)std::thread::hardware_concurrency(), parallel_cap(), n / grain + 1, n ? n : 1}));
    std::vector<std::string> errors(nt);
    std::vector<std::thread> workers;
    auto                     run = [&](const unsigned int t) {
      try
        {
          f(n * t / nt, n * (t + 1) / nt);
        }
      catch (const std::exception &e)
        {
          errors[t] = e.what();
        }
    };
    for (unsigned int t = 1; t < nt; ++t)
      workers.emplace_back(run, t);
    run(0);
    for (auto &w : workers)
      w.join();
    for (const auto &e : errors)
      if (!e.empty())
        throw std::runtime_error(e);
  }

  namespace types
  {
    using global_dof_index = std::uint64_t;
  }
  namespace numbers
  {
    constexpr unsigned int invalid_unsigned_int = 0xFFFFFFFFu;
    constexpr double       PI                   = 3.14159265358979323846;
  } // namespace numbers

  inline std::string ExcMessage(const std::string &s) { return s; }

  using Point3 = std::array<double, 3>;

  // contiguous index range [first, last) -- all the reference needs of IndexSet
  struct IndexSet
  {
    types::global_dof_index first = 0, last = 0;
    types::global_dof_index n_elements() const { return last - first; }
    bool is_element(types::global_dof_index i) const { return i >= first && i < last; }
    types::global_dof_index index_within_set(types::global_dof_index i) const { return i - first; }
    types::global_dof_index nth_index_in_set(types::global_dof_index n) const { return first + n; }
  };

  // ---------------------------------------------------------------------------------------
  // Triangulation: refine_global()-ed subdivided_hyper_rectangle transformed by a chart
  // (benchmark.h:78-88).  Active cells are kept in deal.II's traversal order: coarse cells
  // x-fastest, children by child index cx + 2 cy + 4 cz (Morton order inside a coarse cell).
  // Ranks own equal contiguous chunks of that order (p4est space-filling curve).
  // ---------------------------------------------------------------------------------------
  class Triangulation
  {
  public:
    Triangulation(unsigned int n_ranks = 1, unsigned int this_rank = 0)
      : n_ranks(n_ranks), this_rank(this_rank)
    {}

    void build(const std::array<unsigned int, 3> &subdivisions, const unsigned int n_refine,
               std::function<Point3(const Point3 &)> push_forward)
    {
      sub         = subdivisions;
      refinements = n_refine;
      chart       = std::move(push_forward);
      for (int d = 0; d < 3; ++d)
        n_cells_dir[d] = subdivisions[d] << n_refine;
      const std::uint64_t per_coarse = std::uint64_t(1) << (3 * n_refine);
      cells.resize(per_coarse * sub[0] * sub[1] * sub[2]);
      std::uint64_t idx = 0;
      for (unsigned int kz = 0; kz < sub[2]; ++kz)
        for (unsigned int ky = 0; ky < sub[1]; ++ky)
          for (unsigned int kx = 0; kx < sub[0]; ++kx)
            for (std::uint64_t m = 0; m < per_coarse; ++m, ++idx)
              {
                std::uint32_t l[3] = {0, 0, 0};
                for (unsigned int b = 0; b < n_refine; ++b)
                  for (int d = 0; d < 3; ++d)
                    l[d] |= std::uint32_t((m >> (3 * b + d)) & 1u) << b;
                cells[idx] = {{l[0] + (kx << n_refine), l[1] + (ky << n_refine), l[2] + (kz << n_refine)}};
              }
    }

    std::uint64_t n_global_active_cells() const { return cells.size(); }
    std::uint64_t cells_per_rank() const { return cells.size() / n_ranks; }
    unsigned int  subdomain_id(std::uint64_t cell) const
    {
      const std::uint64_t chunk = cells_per_rank();
      return (unsigned int)std::min<std::uint64_t>(cell / (chunk ? chunk : 1), n_ranks - 1);
    }
    // bits of a 21-bit number moved to every third position (Morton interleave, one dimension)
    static std::uint64_t spread3(std::uint64_t x)
    {
      x &= 0x1fffff;
      x = (x | x << 32) & 0x1f00000000ffffULL;
      x = (x | x << 16) & 0x1f0000ff0000ffULL;
      x = (x | x << 8) & 0x100f00f00f00f00fULL;
      x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
      x = (x | x << 2) & 0x1249249249249249ULL;
      return x;
    }
    // contribution of coordinate c of direction d to the active-cell index: the Morton bits of
    // its position inside the coarse cell plus the coarse cell's share of the linear index
    std::uint64_t index_part(const int d, const std::uint32_t c) const
    {
      const std::uint32_t side = 1u << refinements;
      const std::uint64_t k    = c >> refinements;
      const std::uint64_t coarse = d == 0 ? k : (d == 1 ? k * sub[0] : k * sub[0] * sub[1]);
      return (coarse << (3 * refinements)) + (spread3(c & (side - 1)) << d);
    }
    // active-cell index of the cell at lattice position c (inverse of the traversal order)
    std::uint64_t cell_index(const std::array<std::uint32_t, 3> &c) const
    {
      return index_part(0, c[0]) + index_part(1, c[1]) + index_part(2, c[2]);
    }
    // vertex v = x + 2y + 4z of a cell (poisson_operator.h:153-160)
    Point3 vertex(std::uint64_t cell, unsigned int v) const
    {
      const double h = 1.0 / double(1u << refinements);
      Point3       p;
      for (int d = 0; d < 3; ++d)
        p[d] = double(cells[cell][d] + ((v >> d) & 1u)) * h;
      return chart(p);
    }

    // geometry node (i, j, k) / 2, i, j, k in {0, 1, 2}, of a cell under a quadratic mapping:
    // the chart evaluated at the lattice half-steps (what MappingQ(2) places on the manifold)
    Point3 point27(std::uint64_t cell, unsigned int i, unsigned int j, unsigned int k) const
    {
      const double       h   = 1.0 / double(1u << refinements);
      const unsigned int o[3] = {i, j, k};
      Point3             p;
      for (int d = 0; d < 3; ++d)
        p[d] = (double(cells[cell][d]) + 0.5 * double(o[d])) * h;
      return chart(p);
    }

    std::vector<std::array<std::uint32_t, 3>> cells; // lattice position per active cell
    std::array<unsigned int, 3>               sub{{1, 1, 1}};
    std::array<unsigned int, 3>               n_cells_dir{{1, 1, 1}};
    unsigned int                              refinements = 0;
    unsigned int                              n_ranks, this_rank;
    std::function<Point3(const Point3 &)>     chart;
  };

  // FESystem(FE_Q(degree), n_components): only the sizes are needed
  struct FESystem
  {
    unsigned int degree = 1, n_comp = 3;
    unsigned int n_components() const { return n_comp; }
    unsigned int n_base_elements() const { return 1; }
    unsigned int dofs_per_cell() const { return n_comp * (degree + 1) * (degree + 1) * (degree + 1); }
  };

  // ---------------------------------------------------------------------------------------
  // DoFHandler: Q_p^3 DoFs on the node lattice; a DoF is (lattice node, component) and its
  // global number is 3 * node_number + component.  node_number[] is what
  // Renumber::renumber rewrites (for ALL ranks, see renumber_dofs_for_mf.h here).
  // ---------------------------------------------------------------------------------------
  class DoFHandler
  {
  public:
    explicit DoFHandler(const Triangulation &tria) : tria(&tria) {}

    // initial numbering: rank by rank (owner = lowest rank among the cells touching a node,
    // SURVEY App. B4), lattice order inside a rank, so that every rank owns a contiguous range
    void distribute_dofs(const FESystem &fe_)
    {
      fe = fe_;
      const unsigned int p = fe.degree;
      for (int d = 0; d < 3; ++d)
        nn[d] = std::uint64_t(tria->n_cells_dir[d]) * p + 1;
      n_nodes = nn[0] * nn[1] * nn[2];
      AssertThrow(n_nodes < 0xFFFFFFFFull, "node lattice exceeds 32-bit node numbers");
      owner.resize(n_nodes); // every entry of the three arrays is written by the passes below
      shared.resize(n_nodes);
      AssertThrow(tria->n_ranks <= 255, "at most 255 ranks");
      // owner = min rank over the (up to 8) cells touching a node; shared = more than one rank.
      // Row by row of the lattice, in parallel; per direction the one or two candidate cell
      // coordinates of a lattice coordinate and their shares of the active-cell index are tabulated.
      std::array<std::vector<std::array<std::uint64_t, 2>>, 3> part;
      std::array<std::vector<unsigned char>, 3>                two;
      for (int d = 0; d < 3; ++d)
        {
          part[d].resize(nn[d]);
          two[d].resize(nn[d]);
          for (std::uint64_t t = 0; t < nn[d]; ++t)
            {
              const std::uint32_t q  = (std::uint32_t)(t / p);
              const std::uint32_t hi = std::min<std::uint32_t>(q, tria->n_cells_dir[d] - 1);
              const std::uint32_t lo = (t % p == 0 && q > 0) ? q - 1 : hi;
              part[d][t]             = {{tria->index_part(d, lo), tria->index_part(d, hi)}};
              two[d][t]              = lo != hi;
            }
        }
      const std::uint64_t chunk = std::max<std::uint64_t>(tria->cells_per_rank(), 1);
      const unsigned int  last  = tria->n_ranks - 1;
      parallel_chunks(nn[1] * nn[2], [&](const std::uint64_t r0, const std::uint64_t r1) {
        for (std::uint64_t row = r0; row < r1; ++row)
          {
            const std::uint64_t y = row % nn[1], z = row / nn[1];
            const unsigned int  ny = two[1][y] ? 2 : 1, nz = two[2][z] ? 2 : 1;
            std::uint64_t       yz[4];
            unsigned int        nyz = 0;
            for (unsigned int b = 0; b < nz; ++b)
              for (unsigned int a = 0; a < ny; ++a)
                yz[nyz++] = part[1][y][a] + part[2][z][b];
            unsigned char *o = owner.data() + row * nn[0], *sh = shared.data() + row * nn[0];
            for (std::uint64_t x = 0; x < nn[0]; ++x)
              {
                const unsigned int nx = two[0][x] ? 2 : 1;
                unsigned int       lo = 255, hi = 0;
                for (unsigned int k = 0; k < nyz; ++k)
                  for (unsigned int a = 0; a < nx; ++a)
                    {
                      const unsigned int r = (unsigned int)std::min<std::uint64_t>((part[0][x][a] + yz[k]) / chunk, last);
                      lo = std::min(lo, r);
                      hi = std::max(hi, r);
                    }
                o[x]  = (unsigned char)lo;
                sh[x] = lo != hi;
              }
          }
      }, 16);
      // numbers in lattice order inside every rank: count per slice and rank, prefix sums over
      // (rank, slice), every slice numbers its own nodes
      const unsigned int                      n_slices = 64, nr = tria->n_ranks;
      std::vector<std::vector<std::uint64_t>> cnt(n_slices, std::vector<std::uint64_t>(nr, 0));
      parallel_chunks(n_slices, [&](const std::uint64_t a, const std::uint64_t b) {
        for (std::uint64_t sl = a; sl < b; ++sl)
          for (std::uint64_t n = n_nodes * sl / n_slices; n < n_nodes * (sl + 1) / n_slices; ++n)
            ++cnt[sl][owner[n]];
      }, 1);
      rank_offset.assign(nr + 1, 0);
      std::uint64_t run = 0;
      for (unsigned int r = 0; r < nr; ++r)
        {
          rank_offset[r] = run;
          for (unsigned int sl = 0; sl < n_slices; ++sl)
            {
              const std::uint64_t c = cnt[sl][r];
              cnt[sl][r]            = run;
              run += c;
            }
        }
      rank_offset[nr] = run;
      node_number.resize(n_nodes);
      parallel_chunks(n_slices, [&](const std::uint64_t a, const std::uint64_t b) {
        for (std::uint64_t sl = a; sl < b; ++sl)
          {
            std::vector<std::uint64_t> next = cnt[sl];
            for (std::uint64_t n = n_nodes * sl / n_slices; n < n_nodes * (sl + 1) / n_slices; ++n)
              node_number[n] = (std::uint32_t)next[owner[n]]++;
          }
      }, 1);
    }

    const FESystem      &get_fe() const { return fe; }
    const Triangulation &get_triangulation() const { return *tria; }
    types::global_dof_index n_dofs() const { return 3 * n_nodes; }
    IndexSet                locally_owned_dofs(unsigned int rank) const
    {
      return IndexSet{3 * rank_offset[rank], 3 * rank_offset[rank + 1]};
    }
    IndexSet locally_owned_dofs() const { return locally_owned_dofs(tria->this_rank); }

    // active-cell indices of the cells a lattice node belongs to (1, 2, 4 or 8 of them)
    unsigned int incident_cells(const std::uint64_t node, std::uint64_t (&cells)[8]) const
    {
      const unsigned int  p      = fe.degree;
      const std::uint64_t idx[3] = {node % nn[0], (node / nn[0]) % nn[1], node / (nn[0] * nn[1])};
      std::uint32_t       lo[3], hi[3];
      for (int d = 0; d < 3; ++d)
        {
          const std::uint32_t q = (std::uint32_t)(idx[d] / p);
          hi[d]                 = std::min<std::uint32_t>(q, tria->n_cells_dir[d] - 1);
          lo[d]                 = (idx[d] % p == 0 && q > 0) ? q - 1 : hi[d];
        }
      // the index is a sum of one term per direction: two candidates per direction at most
      std::uint64_t part[3][2];
      for (int d = 0; d < 3; ++d)
        {
          part[d][0] = tria->index_part(d, lo[d]);
          part[d][1] = hi[d] != lo[d] ? tria->index_part(d, hi[d]) : part[d][0];
        }
      unsigned int n = 0;
      for (std::uint32_t z = lo[2]; z <= hi[2]; ++z)
        for (std::uint32_t y = lo[1]; y <= hi[1]; ++y)
          for (std::uint32_t x = lo[0]; x <= hi[0]; ++x)
            cells[n++] = part[0][x - lo[0]] + part[1][y - lo[1]] + part[2][z - lo[2]];
      return n;
    }

    // f(node_begin, node_end) on the lattice rows of the bounding box of `rank`'s cells, in parallel:
    // every node the rank owns or touches lies in there, so the per-node passes of the set-up
    // cost what the rank's share of the lattice costs, not what the whole lattice costs
    template <typename F>
    void parallel_rank_nodes(const unsigned int rank, F &&f) const
    {
      const std::uint64_t chunk = tria->cells_per_rank();
      const std::uint64_t c0 = chunk * rank, c1 = rank + 1 == tria->n_ranks ? tria->n_global_active_cells() : c0 + chunk;
      std::uint32_t       lo[3] = {~0u, ~0u, ~0u}, hi[3] = {0, 0, 0};
      for (std::uint64_t c = c0; c < c1; ++c)
        for (int d = 0; d < 3; ++d)
          {
            lo[d] = std::min(lo[d], tria->cells[c][d]);
            hi[d] = std::max(hi[d], tria->cells[c][d]);
          }
      if (c1 <= c0)
        return;
      const unsigned int  p  = fe.degree;
      const std::uint64_t x0 = std::uint64_t(lo[0]) * p, x1 = std::uint64_t(hi[0] + 1) * p + 1;
      const std::uint64_t y0 = std::uint64_t(lo[1]) * p, ny = std::uint64_t(hi[1] + 1) * p + 1 - y0;
      const std::uint64_t z0 = std::uint64_t(lo[2]) * p, nz = std::uint64_t(hi[2] + 1) * p + 1 - z0;
      parallel_chunks(ny * nz, [&](const std::uint64_t r0, const std::uint64_t r1) {
        for (std::uint64_t row = r0; row < r1; ++row)
          {
            const std::uint64_t base = ((z0 + row / ny) * nn[1] + (y0 + row % ny)) * nn[0];
            f(base + x0, base + x1);
          }
      }, 16);
    }

    template <typename F>
    void for_each_cell_node(std::uint64_t cell, F &&f) const
    {
      const unsigned int  p = fe.degree;
      const auto         &c = tria->cells[cell];
      const std::uint64_t I0 = std::uint64_t(c[0]) * p, J0 = std::uint64_t(c[1]) * p,
                          K0 = std::uint64_t(c[2]) * p;
      for (unsigned int k = 0; k <= p; ++k)
        for (unsigned int j = 0; j <= p; ++j)
          for (unsigned int i = 0; i <= p; ++i)
            f(((K0 + k) * nn[1] + (J0 + j)) * nn[0] + (I0 + i), (int)i, (int)j, (int)k);
    }
    std::uint64_t cell_node(std::uint64_t cell, unsigned int i, unsigned int j, unsigned int k) const
    {
      const unsigned int p = fe.degree;
      const auto        &c = tria->cells[cell];
      return ((std::uint64_t(c[2]) * p + k) * nn[1] + (std::uint64_t(c[1]) * p + j)) * nn[0] +
             (std::uint64_t(c[0]) * p + i);
    }
    bool node_on_boundary(std::uint64_t node) const
    {
      const std::uint64_t I = node % nn[0], J = (node / nn[0]) % nn[1], K = node / (nn[0] * nn[1]);
      return I == 0 || I == nn[0] - 1 || J == 0 || J == nn[1] - 1 || K == 0 || K == nn[2] - 1;
    }
    types::global_dof_index dof_number(std::uint64_t node, unsigned int c) const
    {
      return 3 * types::global_dof_index(node_number[node]) + c;
    }

    std::array<std::uint64_t, 3> nn{{1, 1, 1}};
    std::uint64_t                n_nodes = 0;
    raw_vector<unsigned char>    owner;       // [n_nodes] owning rank
    raw_vector<unsigned char>    shared;      // [n_nodes] touched by cells of several ranks
    raw_vector<std::uint32_t>    node_number; // [n_nodes] current global node number
    std::vector<std::uint64_t>   rank_offset; // [n_ranks+1] first node number of each rank

  private:
    const Triangulation *tria;
    FESystem             fe;
  };

  // all boundary DoFs constrained to zero (VectorTools::interpolate_boundary_values with a
  // ZeroFunction, benchmark.h:96-102).  Geometric, hence independent of the numbering.
  class AffineConstraints
  {
  public:
    void reinit(const DoFHandler &dh) { dof_handler = &dh; }
    void clear() {}
    void close() {}
    bool node_is_constrained(std::uint64_t node) const { return dof_handler->node_on_boundary(node); }
    const DoFHandler *dof_handler = nullptr;
  };

  namespace VectorTools
  {
    inline void interpolate_boundary_values(const DoFHandler &dof_handler, AffineConstraints &constraints)
    {
      constraints.reinit(dof_handler);
    }
  } // namespace VectorTools

  namespace Utilities
  {
    namespace MPI
    {
      // layout of LinearAlgebra::distributed::Vector: owned range, then ghosts sorted by
      // global index, plus the point-to-point exchange plan
      struct Partitioner
      {
        IndexSet                   owned;       // global DoF range
        std::vector<std::uint32_t> ghost_nodes; // global node numbers of the ghosts, sorted
        unsigned int locally_owned_size() const { return (unsigned int)owned.n_elements(); }
        unsigned int n_ghost_indices() const { return 3u * (unsigned int)ghost_nodes.size(); }
        unsigned int global_to_local(types::global_dof_index g) const
        {
          if (owned.is_element(g))
            return (unsigned int)(g - owned.first);
          const std::uint32_t node = (std::uint32_t)(g / 3);
          const auto          it   = std::lower_bound(ghost_nodes.begin(), ghost_nodes.end(), node);
          AssertThrow(it != ghost_nodes.end() && *it == node, "global index is neither owned nor ghost");
          return locally_owned_size() + 3u * (unsigned int)(it - ghost_nodes.begin()) + (unsigned int)(g % 3);
        }
        types::global_dof_index local_to_global(unsigned int l) const
        {
          if (l < locally_owned_size())
            return owned.first + l;
          const unsigned int g = l - locally_owned_size();
          return 3 * types::global_dof_index(ghost_nodes[g / 3]) + g % 3;
        }
        std::vector<int>           peers;         // ranks I exchange with, ascending
        std::vector<std::uint64_t> import_offset; // [peers+1] DoF offsets into the ghost range
        std::vector<std::uint64_t> export_offset; // [peers+1]
        std::vector<std::uint32_t> export_index;  // owned local DoFs each peer ghosts
      };
    } // namespace MPI
  }   // namespace Utilities
} // namespace dealii
