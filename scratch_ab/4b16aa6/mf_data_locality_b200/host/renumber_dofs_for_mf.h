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
//  * Every process computes the numbering of ALL ranks (the mesh is structured and cheap to
//    re-derive), which replaces the ghost-number exchange inside
//    DoFHandler::renumber_dofs (renumber_dofs_for_mf.h:144).
//  * assembly strategy 1 (cellbatch_assembly, :363-459) interleaves components of different
//    nodes, which the compressed operator rejects ("Expected contiguous numbering",
//    poisson_operator.h:198); it is listed as "next"  and throws here.
#pragma once
#include <sstream>

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
    AssertThrow(assembly_strat == 0, "cellbatch assembly is not supported by the compressed operator");
    AssertThrow(renumber_strat <= 2 && grouping_strat <= 2, "unknown renumbering strategy");
    const unsigned int         n_ranks = dof_handler.get_triangulation().n_ranks;
    std::vector<std::uint32_t> new_node_number(dof_handler.n_nodes);
    for (unsigned int rank = 0; rank < n_ranks; ++rank)
      {
        dealii::MatrixFree matrix_free;
        matrix_free.reinit(dof_handler, constraints, dof_handler.get_fe().degree + 1, mf_data, (int)rank);
        const std::uint64_t first = dof_handler.rank_offset[rank],
                            n_own = dof_handler.rank_offset[rank + 1] - first;
        // key[i] = position of owned node i in the matrix-free traversal
        std::vector<std::uint64_t> key = cell_assembly(matrix_free);
        AssertThrow(key.size() == n_own, "Expected " + std::to_string(n_own) + " nodes");
        const std::vector<std::uint32_t> new_numbers = grouping(matrix_free, key);
        AssertThrow(new_numbers.size() == n_own, "Dimension mismatch " + std::to_string(new_numbers.size()) +
                                                   " vs " + std::to_string(n_own));
        // new_numbers[i] = old owned index that moves to position i (:139-144)
        std::vector<std::uint32_t> new_of_old(n_own);
        for (std::uint64_t i = 0; i < n_own; ++i)
          new_of_old[new_numbers[i]] = (std::uint32_t)(first + i);
        for (std::uint64_t n = 0; n < dof_handler.n_nodes; ++n)
          if (dof_handler.owner[n] == rank)
            new_node_number[n] = new_of_old[dof_handler.node_number[n] - first];
      }
    dof_handler.node_number.swap(new_node_number);
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

  // cell_assembly with first_touch_renumber / last_touch_renumber (:247-361, :461-490)
  std::vector<std::uint64_t> cell_assembly(const dealii::MatrixFree &mf) const
  {
    const dealii::DoFHandler &dh    = mf.get_dof_handler();
    const unsigned int        rank  = mf.get_rank();
    const std::uint64_t       first = dh.rank_offset[rank], n_own = dh.rank_offset[rank + 1] - first;
    constexpr std::uint64_t   unset = ~std::uint64_t(0);
    std::vector<std::uint64_t> key(n_own, unset);
    std::uint64_t              counter = 0;
    for (unsigned int b = 0; b < mf.n_cell_batches(); ++b)
      for (unsigned int l = 0; l < mf.n_active_entries_per_cell_batch(b); ++l)
        walk_cell_objects(dh, mf.get_cell(b, l), [&](const std::uint64_t node, unsigned int) {
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
        });
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
    std::vector<unsigned char> count(n_own, 0);
    std::vector<std::uint32_t> last_group(n_own, 0xFFFFFFFFu);
    const auto                &ti      = mf.get_task_info();
    std::uint32_t              group   = 0;
    auto                       visit_batch = [&](const unsigned int b) {
      for (unsigned int l = 0; l < mf.n_active_entries_per_cell_batch(b); ++l)
        dh.for_each_cell_node(mf.get_cell(b, l), [&](const std::uint64_t node, int, int, int) {
          if (dh.owner[node] != rank || mf.get_constraints().node_is_constrained(node))
            return;
          const std::uint64_t i = dh.node_number[node] - first;
          if (last_group[i] != group)
            {
              last_group[i] = group;
              ++count[i];
            }
        });
    };
    if (by_range)
      for (unsigned int part = 0; part + 2 < ti.partition_row_index.size(); ++part)
        for (unsigned int r = ti.partition_row_index[part]; r < ti.partition_row_index[part + 1]; ++r, ++group)
          for (unsigned int b = ti.cell_partition_data[r]; b < ti.cell_partition_data[r + 1]; ++b)
            visit_batch(b);
    else
      for (unsigned int b = 0; b < mf.n_cell_batches(); ++b, ++group)
        visit_batch(b);
    return count;
  }

  // grouping (:492-535) with base_grouping (:537-554) / touch_count_grouping (:556-590);
  // "multi-domain" nodes = owned nodes that also sit on a ghost cell (domain_dof_mapping, :673-730)
  std::vector<std::uint32_t> grouping(const dealii::MatrixFree &mf, const std::vector<std::uint64_t> &key) const
  {
    const dealii::DoFHandler &dh    = mf.get_dof_handler();
    const unsigned int        rank  = mf.get_rank();
    const std::uint64_t       first = dh.rank_offset[rank], n_own = dh.rank_offset[rank + 1] - first;
    std::vector<unsigned char> group(n_own, 0); // 0: single range, 1: several/none, 2: multi-rank
    if (grouping_strat != 0)
      {
        const std::vector<unsigned char> tc = touch_count(mf, grouping_strat == 2);
        for (std::uint64_t i = 0; i < n_own; ++i)
          group[i] = tc[i] == 1 ? 0 : 1;
      }
    for (std::uint64_t n = 0; n < dh.n_nodes; ++n)
      if (dh.owner[n] == rank && dh.shared[n])
        group[dh.node_number[n] - first] = 2;
    // stable order by key inside each group
    std::vector<std::uint32_t> by_key(n_own);
    std::iota(by_key.begin(), by_key.end(), 0u);
    std::sort(by_key.begin(), by_key.end(), [&](std::uint32_t a, std::uint32_t b) { return key[a] < key[b]; });
    std::vector<std::uint32_t> out;
    out.reserve(n_own);
    for (unsigned char g = 0; g < 3; ++g)
      for (const std::uint32_t i : by_key)
        if (group[i] == g)
          out.push_back(i);
    return out;
  }

  const unsigned int assembly_strat, renumber_strat, grouping_strat;
};
