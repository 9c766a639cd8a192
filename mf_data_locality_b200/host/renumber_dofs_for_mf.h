// renumber_dofs_for_mf.h -- DoF renumbering for data locality of the matrix-free loop.
// Host-side mirror of the reference's Renumber<dim,Number> (renumber_dofs_for_mf.h:14-730):
// same constructor triple (assembly, renumber, grouping strategy), same renumber() entry
// point and get_renumber_string(), same result: DoFs numbered in the order the cell loop
// first (or last) touches them, then grouped into
//   [touched by exactly one cell-batch (range)] [touched by several / none] [shared between ranks]
// (renumber_dofs_for_mf.h:492-535, :556-590).
//
// Implementation notes (this is not a translation):
//  * The three vector components of a node are always touched together and in order
//    (renumber_dofs_for_mf.h:340-356 loops c innermost), so the algorithm runs on lattice
//    NODES with flat arrays; DoF number = 3 * node number + component.  This is what lets a
//    50-800 M DoF numbering finish in seconds on the host.
//  * No ghost-number exchange (DoFHandler::renumber_dofs, renumber_dofs_for_mf.h:144): the mesh is
//    structured, so a process derives the numbers of its ghosts itself.  For the default
//    strategy (cell assembly, first touch) it numbers its own rank completely and, of every
//    other rank, only the nodes shared between ranks -- the owner's last group, ordered by the
//    owner's cell loop (shared_numbers_of); for the other strategies it numbers all ranks.
//  * assembly strategy 1 (cellbatch_assembly, :363-459) numbers batch by batch, FE_Q slot by
//    slot, lane by lane.  The reference walks the (p+1)^3 slots of a scalar element there; on
//    lattice nodes that is exactly this walk.  It interleaves the nodes of one entity over the
//    cells of a batch, so for p >= 3 LaplaceOperator::initialize rejects the result ("Expected
//    contiguous numbering", poisson_operator.h:198) -- in the reference as here; for p = 2
//    (one node per entity) it is a valid input of the operator.
#pragma once
#include <array>
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <thread>

#include "matrix_free_standin.h"

template <int dim, typename Number>
class Renumber
{
public:
  Renumber(const unsigned int a, const unsigned int r, const unsigned int g)
    : assembly_strat(a), renumber_strat(r), grouping_strat(g)
  {}

  void renumber(dealii::DoFHandler &dof_handler, const dealii::AffineConstraints &constraints,
                const dealii::MatrixFree::AdditionalData &mf_data) const
  {
    static_assert(dim == 3, "the BP4 path is three-dimensional");
    if (renumber_strat == 0) // "base": keep the numbering (renumber_dofs_for_mf.h:111-113)
      return;
    AssertThrow(assembly_strat <= 1 && renumber_strat <= 2 && grouping_strat <= 2,
                "unknown renumbering strategy");
    const unsigned int n_ranks = dof_handler.get_triangulation().n_ranks;
    // The ranks' numberings are independent of each other (a rank reads and rewrites the numbers
    // of its own nodes only), so they are applied in place and may run concurrently.
    auto renumber_rank = [&](const unsigned int rank) {
        const bool show = std::getenv("BP4_SETUP_TIMING") != nullptr && rank == 0;
        auto       t0   = std::chrono::steady_clock::now();
        auto       lap  = [&](const char *what) {
          const auto t1 = std::chrono::steady_clock::now();
          if (show)
            std::cerr << "    renumber: " << what << " " << std::chrono::duration<double>(t1 - t0).count() << " s\n";
          t0 = t1;
        };
        dealii::MatrixFree matrix_free;
        matrix_free.reinit(dof_handler, constraints, dof_handler.get_fe().degree + 1, mf_data, (int)rank, true);
        lap("matrix_free.reinit");
        const std::uint64_t first = dof_handler.rank_offset[rank],
                            n_own = dof_handler.rank_offset[rank + 1] - first;
        // key[i] = position of owned node i in the matrix-free traversal
        std::vector<std::uint64_t> key = cell_assembly(matrix_free);
        AssertThrow(key.size() == n_own, "Expected " + std::to_string(n_own) + " nodes");
        lap("cell_assembly");
        const auto new_numbers = grouping(matrix_free, key);
        lap("grouping");
        AssertThrow(new_numbers.size() == n_own, "Dimension mismatch " + std::to_string(new_numbers.size()) +
                                                   " vs " + std::to_string(n_own));
        // new_numbers[i] = old owned index that moves to position i (:139-144)
        dealii::raw_vector<std::uint32_t> new_of_old(n_own); // written completely below
        dealii::parallel_chunks(n_own, [&](const std::uint64_t a, const std::uint64_t b) {
          for (std::uint64_t i = a; i < b; ++i)
            new_of_old[new_numbers[i]] = (std::uint32_t)(first + i);
        });
        dof_handler.parallel_rank_nodes(rank, [&](const std::uint64_t a, const std::uint64_t b) {
          for (std::uint64_t n = a; n < b; ++n)
            if (dof_handler.owner[n] == rank)
              dof_handler.node_number[n] = new_of_old[dof_handler.node_number[n] - first];
        });
        lap("apply");
    };
    // Default strategy on several ranks: this process needs the complete numbering of ITS rank
    // only.  Of the other ranks it only ever references nodes shared between ranks (its ghosts),
    // and those are the last group of their owner, ordered by first touch: their numbers follow
    // from the owner's cell loop order alone (shared_numbers_of), no pass over the owner's nodes.
    const unsigned int this_rank = dof_handler.get_triangulation().this_rank;
    // (BP4_RENUMBER_ALL_RANKS=1: number every rank completely, as for the other strategies)
    if (n_ranks > 1 && assembly_strat == 0 && renumber_strat == 1 && std::getenv("BP4_RENUMBER_ALL_RANKS") == nullptr)
      {
        renumber_rank(this_rank);
        std::vector<std::string> errs(n_ranks);
        dealii::parallel_chunks(n_ranks, [&](const std::uint64_t a, const std::uint64_t b) {
          const unsigned int cap_before = dealii::parallel_cap();
          dealii::parallel_cap()        = std::max(1u, 16u / n_ranks);
          for (std::uint64_t q = a; q < b; ++q)
            if (q != this_rank)
              try
                {
                  shared_numbers_of((unsigned int)q, dof_handler, constraints, mf_data);
                }
              catch (const std::exception &e)
                {
                  errs[q] = e.what();
                }
          dealii::parallel_cap() = cap_before;
        }, 1);
        for (const auto &e : errs)
          AssertThrow(e.empty(), e);
        // every other node of the other ranks keeps its old number: never referenced here
        return;
      }
    std::vector<std::string> errors(n_ranks);
    std::vector<std::thread> workers;
    for (unsigned int rank = 1; rank < n_ranks; ++rank)
      workers.emplace_back([&, rank] {
        try
          {
            dealii::parallel_cap() = std::max(1u, 16u / n_ranks); // the rank workers share the cores
            renumber_rank(rank);
          }
        catch (const std::exception &e)
          {
            errors[rank] = e.what();
          }
      });
    const unsigned int cap_before = dealii::parallel_cap();
    try
      {
        dealii::parallel_cap() = std::max(1u, 16u / n_ranks);
        renumber_rank(0);
      }
    catch (const std::exception &e)
      {
        errors[0] = e.what();
      }
    dealii::parallel_cap() = cap_before;
    for (auto &w : workers)
      w.join();
    for (const auto &e : errors)
      AssertThrow(e.empty(), e);
  }

  std::string get_renumber_string() const
  {
    static const char *a[] = {"cell", "cellbatch"}, *r[] = {"base", "first", "last"},
                      *g[] = {"base", "cellbatch", "cellbatch_range"};
    std::stringstream ss;
    ss << a[assembly_strat] << "-" << r[renumber_strat] << "-" << g[grouping_strat];
    return ss.str();
  }

private:
  // lexicographic walk of the 27 cell objects a = ex + 3 ey + 9 ez, nodes lexicographic inside
  // each object: the order of renumber_dofs_for_mf.h:333-357 (object table :289-316; the
  // i1-outer loop for a = 10, 16 is the lexicographic walk of the y-faces)
  template <typename F>
  static void walk_cell_objects(const dealii::DoFHandler &dh, const std::uint64_t cell, F &&f)
  {
    const unsigned int p = dh.get_fe().degree;
    const unsigned int lo[3] = {0, 1, p}, hi[3] = {1, p, p + 1};
    for (unsigned int a = 0; a < 27; ++a)
      {
        const unsigned int ex = a % 3, ey = (a / 3) % 3, ez = a / 9;
        for (unsigned int k = lo[ez]; k < hi[ez]; ++k)
          for (unsigned int j = lo[ey]; j < hi[ey]; ++j)
            for (unsigned int i = lo[ex]; i < hi[ex]; ++i)
              f(dh.cell_node(cell, i, j, k), a);
      }
  }

  // cell-local nodes (i, j, k) in deal.II's FE_Q<3>(p) numbering: 8 vertices, 12 lines, 6 quads
  // (local coordinates (y,z), (z,x), (x,y), first one fastest), interior -- the order `cf` runs
  // through in cellbatch_assembly (renumber_dofs_for_mf.h:394-456)
  static std::vector<std::array<unsigned int, 3>> fe_q_walk(const unsigned int p)
  {
    std::vector<std::array<unsigned int, 3>> w;
    for (unsigned int v = 0; v < 8; ++v)
      w.push_back({{(v & 1) * p, ((v >> 1) & 1) * p, ((v >> 2) & 1) * p}});
    for (unsigned int z : {0u, p})
      {
        for (unsigned int x : {0u, p})
          for (unsigned int t = 1; t < p; ++t)
            w.push_back({{x, t, z}});
        for (unsigned int y : {0u, p})
          for (unsigned int t = 1; t < p; ++t)
            w.push_back({{t, y, z}});
      }
    for (unsigned int y : {0u, p})
      for (unsigned int x : {0u, p})
        for (unsigned int t = 1; t < p; ++t)
          w.push_back({{x, y, t}});
    for (unsigned int x : {0u, p})
      for (unsigned int b = 1; b < p; ++b)
        for (unsigned int a = 1; a < p; ++a)
          w.push_back({{x, a, b}});
    for (unsigned int y : {0u, p})
      for (unsigned int a = 1; a < p; ++a)
        for (unsigned int b = 1; b < p; ++b)
          w.push_back({{a, y, b}});
    for (unsigned int z : {0u, p})
      for (unsigned int b = 1; b < p; ++b)
        for (unsigned int a = 1; a < p; ++a)
          w.push_back({{a, b, z}});
    for (unsigned int k = 1; k < p; ++k)
      for (unsigned int j = 1; j < p; ++j)
        for (unsigned int i = 1; i < p; ++i)
          w.push_back({{i, j, k}});
    return w;
  }

  // New numbers of the nodes rank q owns AND shares with other ranks, for (cell_assembly,
  // first_touch, any grouping): they form q's last group (grouping(): group 2), ordered by first
  // touch = (loop position of the first of q's cells around the node, position of the node in
  // that cell's object walk).  Only q's cells that hold such a node are looked at.
  void shared_numbers_of(const unsigned int q, dealii::DoFHandler &dh, const dealii::AffineConstraints &con,
                         const dealii::MatrixFree::AdditionalData &mf_data) const
  {
    dealii::MatrixFree mf;
    mf.reinit(dh, con, dh.get_fe().degree + 1, mf_data, (int)q, true);
    const std::uint64_t first = dh.rank_offset[q], n_own = dh.rank_offset[q + 1] - first;
    const unsigned int  p = dh.get_fe().degree, rep[3] = {0, 1, p};
    std::vector<std::uint64_t> order; // the shared nodes of q in the order of their first touch
    for (std::uint32_t pos = 0; pos < mf.n_physical_cells(); ++pos)
      {
        const std::uint64_t cell = mf.cell_order[pos];
        // all nodes of an entity sit on the same cells: one representative per entity tells
        // whether the cell holds a node that q owns and shares
        bool any = false;
        for (unsigned int a = 0; a < 27 && !any; ++a)
          if (p > 1 || (a % 3 != 1 && (a / 3) % 3 != 1 && a / 9 != 1))
            {
              const std::uint64_t node = dh.cell_node(cell, rep[a % 3], rep[(a / 3) % 3], rep[a / 9]);
              any = dh.shared[node] && dh.owner[node] == q;
            }
        if (!any)
          continue;
        walk_cell_objects(dh, cell, [&](const std::uint64_t node, unsigned int) {
          if (dh.owner[node] != q || !dh.shared[node])
            return;
          std::uint32_t      ps[8];
          const unsigned int n  = mf.incident_positions(node, ps);
          std::uint32_t      lo = 0xFFFFFFFFu;
          for (unsigned int j = 0; j < n; ++j)
            lo = std::min(lo, ps[j]);
          if (lo == pos) // this cell is the first one of q's loop that touches the node
            order.push_back(node);
        });
      }
    AssertThrow(order.size() <= n_own, "shared nodes of a neighbour rank: count mismatch");
    const std::uint64_t base = first + n_own - order.size(); // the last group of q
    for (std::size_t k = 0; k < order.size(); ++k)
      dh.node_number[order[k]] = (std::uint32_t)(base + k);
  }

  // cell_assembly / cellbatch_assembly with first_touch_renumber / last_touch_renumber
  // (:247-361, :363-459, :461-490)
  std::vector<std::uint64_t> cell_assembly(const dealii::MatrixFree &mf) const
  {
    const dealii::DoFHandler &dh    = mf.get_dof_handler();
    const unsigned int        rank  = mf.get_rank();
    const std::uint64_t       first = dh.rank_offset[rank], n_own = dh.rank_offset[rank + 1] - first;
    constexpr std::uint64_t   unset = ~std::uint64_t(0);
    std::vector<std::uint64_t> key(n_own, unset);
    std::uint64_t              counter = 0;
    if (assembly_strat == 0 && renumber_strat == 1)
      {
        // First touch, cell by cell, without the sequential sweep: a node is numbered by the
        // FIRST cell of the loop that holds it (the smallest loop position among the up to 8
        // cells around it), and inside that cell in the order of the object walk.  So: cells
        // count the owned nodes they are first for (parallel), a prefix sum over the loop
        // order gives every cell its first key, and a second parallel pass hands the keys out.
        const std::uint32_t        n_cells = mf.n_physical_cells();
        std::vector<std::uint64_t> first_key(n_cells + 1, 0);
        std::vector<std::uint32_t> first_pos(n_own, 0xFFFFFFFFu); // loop position of the first cell
        dh.parallel_rank_nodes(rank, [&](const std::uint64_t a, const std::uint64_t b) {
          for (std::uint64_t node = a; node < b; ++node)
            {
              if (dh.owner[node] != rank)
                continue;
              std::uint32_t      ps[8];
              const unsigned int n  = mf.incident_positions(node, ps);
              std::uint32_t      lo = 0xFFFFFFFFu;
              for (unsigned int k = 0; k < n; ++k)
                lo = std::min(lo, ps[k]);
              first_pos[dh.node_number[node] - first] = lo;
            }
        });
        auto for_first_nodes = [&](const std::uint32_t pos, auto &&f) {
          walk_cell_objects(dh, mf.cell_order[pos], [&](const std::uint64_t node, unsigned int) {
            if (dh.owner[node] == rank && first_pos[dh.node_number[node] - first] == pos)
              f(node);
          });
        };
        dealii::parallel_chunks(n_cells, [&](const std::uint64_t a, const std::uint64_t b) {
          for (std::uint64_t pos = a; pos < b; ++pos)
            {
              std::uint64_t cnt = 0;
              for_first_nodes((std::uint32_t)pos, [&](std::uint64_t) { ++cnt; });
              first_key[pos + 1] = cnt;
            }
        });
        for (std::uint32_t pos = 0; pos < n_cells; ++pos)
          first_key[pos + 1] += first_key[pos];
        dealii::parallel_chunks(n_cells, [&](const std::uint64_t a, const std::uint64_t b) {
          for (std::uint64_t pos = a; pos < b; ++pos)
            {
              std::uint64_t k = first_key[pos];
              for_first_nodes((std::uint32_t)pos, [&](const std::uint64_t node) {
                key[dh.node_number[node] - first] = k++;
              });
            }
        });
        dealii::parallel_chunks(n_own, [&](const std::uint64_t a, const std::uint64_t b) {
          for (std::uint64_t i = a; i < b; ++i)
            AssertThrow(key[i] != unset, "owned node never touched by a local cell");
        });
        return key;
      }
    auto touch = [&](const std::uint64_t node) {
      if (dh.owner[node] != rank)
        return;
      std::uint64_t &k = key[dh.node_number[node] - first];
      if (renumber_strat == 1)
        {
          if (k == unset)
            k = counter++;
        }
      else // last touch: the by-value set copy at :481 makes every touch renumber
        k = counter++;
    };
    if (assembly_strat == 0)
      for (unsigned int b = 0; b < mf.n_cell_batches(); ++b)
        for (unsigned int l = 0; l < mf.n_active_entries_per_cell_batch(b); ++l)
          walk_cell_objects(dh, mf.get_cell(b, l), [&](const std::uint64_t node, unsigned int) { touch(node); });
    else
      {
        const auto slots = fe_q_walk(dh.get_fe().degree);
        for (unsigned int b = 0; b < mf.n_cell_batches(); ++b)
          for (const auto &ijk : slots)
            for (unsigned int l = 0; l < mf.n_active_entries_per_cell_batch(b); ++l)
              touch(dh.cell_node(mf.get_cell(b, l), ijk[0], ijk[1], ijk[2]));
      }
    for (const std::uint64_t k : key)
      AssertThrow(k != unset, "owned node never touched by a local cell");
    return key;
  }

  // number of cell batches / cell-batch ranges touching every owned node; constrained nodes
  // are absent from MatrixFree's index lists, hence 0 (:592-671)
  std::vector<unsigned char> touch_count(const dealii::MatrixFree &mf, const bool by_range) const
  {
    const dealii::DoFHandler &dh    = mf.get_dof_handler();
    const unsigned int        rank  = mf.get_rank();
    const std::uint64_t       first = dh.rank_offset[rank], n_own = dh.rank_offset[rank + 1] - first;
    // per node: the distinct groups (cell batches or cell-batch ranges) among the local cells
    // around it -- the same count the sweep over the loop produces, node by node in parallel
    std::vector<unsigned char>        count(n_own, 0);
    const std::vector<std::uint32_t> &group_of_pos = by_range ? mf.range_of_pos : mf.batch_of_pos;
    dh.parallel_rank_nodes(rank, [&](const std::uint64_t a, const std::uint64_t b) {
      for (std::uint64_t node = a; node < b; ++node)
        {
          if (dh.owner[node] != rank || mf.get_constraints().node_is_constrained(node))
            continue;
          std::uint32_t      ps[8], seen[8];
          const unsigned int n = mf.incident_positions(node, ps);
          unsigned int       c = 0;
          for (unsigned int k = 0; k < n; ++k)
            {
              const std::uint32_t g   = group_of_pos[ps[k]];
              bool                dup = false;
              for (unsigned int q = 0; q < c; ++q)
                dup |= seen[q] == g;
              if (!dup)
                seen[c++] = g;
            }
          count[dh.node_number[node] - first] = (unsigned char)c;
        }
    });
    return count;
  }

  // grouping (:492-535) with base_grouping (:537-554) / touch_count_grouping (:556-590);
  // "multi-domain" nodes = owned nodes that also sit on a ghost cell (domain_dof_mapping, :673-730)
  dealii::raw_vector<std::uint32_t> grouping(const dealii::MatrixFree &mf, const std::vector<std::uint64_t> &key) const
  {
    const dealii::DoFHandler &dh    = mf.get_dof_handler();
    const unsigned int        rank  = mf.get_rank();
    const std::uint64_t       first = dh.rank_offset[rank], n_own = dh.rank_offset[rank + 1] - first;
    std::vector<unsigned char> group(n_own, 0); // 0: single range, 1: several/none, 2: multi-rank
    if (grouping_strat != 0)
      {
        const std::vector<unsigned char> tc = touch_count(mf, grouping_strat == 2);
        dealii::parallel_chunks(n_own, [&](const std::uint64_t a, const std::uint64_t b) {
          for (std::uint64_t i = a; i < b; ++i)
            group[i] = tc[i] == 1 ? 0 : 1;
        });
      }
    dh.parallel_rank_nodes(rank, [&](const std::uint64_t a, const std::uint64_t b) {
      for (std::uint64_t n = a; n < b; ++n)
        if (dh.owner[n] == rank && dh.shared[n])
          group[dh.node_number[n] - first] = 2;
    });
    // order by key inside each group.  First-touch keys are a permutation of 0 .. n_own-1: one
    // scatter instead of a sort; last-touch keys have gaps (every touch draws a new number)
    dealii::raw_vector<std::uint32_t> by_key(n_own); // both written completely below
    if (renumber_strat == 1)
      dealii::parallel_chunks(n_own, [&](const std::uint64_t a, const std::uint64_t b) {
        for (std::uint64_t i = a; i < b; ++i)
          by_key[key[i]] = (std::uint32_t)i;
      });
    else
      {
        std::iota(by_key.begin(), by_key.end(), 0u);
        std::sort(by_key.begin(), by_key.end(), [&](std::uint32_t a, std::uint32_t b) { return key[a] < key[b]; });
      }
    // stable partition by group: count per slice and group, prefix sums, every slice writes its own
    const unsigned int                        n_slices = 64;
    std::vector<std::array<std::uint64_t, 3>> cnt(n_slices, std::array<std::uint64_t, 3>{{0, 0, 0}});
    dealii::parallel_chunks(n_slices, [&](const std::uint64_t a, const std::uint64_t b) {
      for (std::uint64_t sl = a; sl < b; ++sl)
        for (std::uint64_t q = n_own * sl / n_slices; q < n_own * (sl + 1) / n_slices; ++q)
          ++cnt[sl][group[by_key[q]]];
    }, 1);
    std::uint64_t run = 0;
    for (unsigned int g = 0; g < 3; ++g)
      for (unsigned int sl = 0; sl < n_slices; ++sl)
        {
          const std::uint64_t c = cnt[sl][g];
          cnt[sl][g]            = run;
          run += c;
        }
    dealii::raw_vector<std::uint32_t> out(n_own);
    dealii::parallel_chunks(n_slices, [&](const std::uint64_t a, const std::uint64_t b) {
      for (std::uint64_t sl = a; sl < b; ++sl)
        {
          std::array<std::uint64_t, 3> at = cnt[sl];
          for (std::uint64_t q = n_own * sl / n_slices; q < n_own * (sl + 1) / n_slices; ++q)
            out[at[group[by_key[q]]]++] = by_key[q];
        }
    }, 1);
    return out;
  }

  const unsigned int assembly_strat, renumber_strat, grouping_strat;
};
