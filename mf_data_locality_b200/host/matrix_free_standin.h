// matrix_free_standin.h -- the MatrixFree<dim,double> roles used by the reference's
// Renumber (renumber_dofs_for_mf.h:115-124, :320-331, :601-669) and LaplaceOperator
// (poisson_operator.h:108-135, :188-196, :311): cell batches in loop order, the task_info
// range tables, the vector partitioner, constrained DoFs.  Behaviour recalled from deal.II
// 9.3 with the non-observable choices exposed in AdditionalData.
#pragma once
#include <mutex>
#include "dealii_standin.h"

namespace dealii
{
  struct MatrixFreeAdditionalData
  {
    unsigned int n_lanes            = 8;    // VectorizedArray<double>::size(), AVX-512
    unsigned int batches_per_range  = 1;    // cell batches per cell_partition_data range
    bool         initialize_mapping = true; // unused: geometry is evaluated on the fly
    // degree of the Mapping handed to MatrixFree::reinit in the reference (MappingQGeneric(1),
    // benchmark.h:89).  2 = genuinely quadratic cells: LaplaceOperator::initialize then fills all
    // 27 coefficient vectors (the TODO of benchmark.h:75-77, poisson_operator.h:169 "for now")
    unsigned int mapping_degree = 1;
  };

  class MatrixFree
  {
  public:
    using AdditionalData = MatrixFreeAdditionalData;
    struct TaskInfo
    {
      std::vector<unsigned int> partition_row_index; // ranges per partition (+2 trailing entries)
      std::vector<unsigned int> cell_partition_data; // first batch of every range (+ sentinel)
    };
    struct DoFInfo
    {
      std::shared_ptr<const Utilities::MPI::Partitioner> vector_partitioner;
    };

    // `rank` defaults to the triangulation's own rank; Renumber also builds the object for the
    // other ranks (every process derives the whole numbering, no host message passing)
    // order_only: just the cell loop order of `rank_` (cell_order, cell_pos, batches, ranges) --
    // what Renumber needs of a NEIGHBOUR rank to number the nodes it shares with it
    void reinit(const DoFHandler &dh, const AffineConstraints &con, const unsigned int n_q_points_1d_,
                const AdditionalData &ad = AdditionalData(), const int rank_ = -1, const bool order_only = false)
    {
      dof_handler   = &dh;
      constraints   = &con;
      n_q_points_1d = n_q_points_1d_;
      data          = ad;
      const Triangulation &tria = dh.get_triangulation();
      rank = rank_ < 0 ? tria.this_rank : (unsigned int)rank_;
      const std::uint64_t chunk = tria.cells_per_rank();
      const std::uint64_t c0 = chunk * rank,
                          c1 = rank + 1 == tria.n_ranks ? tria.n_global_active_cells() : c0 + chunk;
      // cells that touch a node owned elsewhere go to the middle partition (they read ghost
      // values / write ghost contributions); the others are split around it
      std::vector<std::uint64_t> inner, comm;
      {
        std::vector<unsigned char> needs_comm(c1 - c0, 0);
        if (tria.n_ranks > 1)
          parallel_chunks(c1 - c0, [&](const std::uint64_t a, const std::uint64_t b) {
            for (std::uint64_t c = c0 + a; c < c0 + b; ++c)
              {
                // all nodes of an entity (vertex, line, face, interior) sit on the same cells and
                // hence have the same owner: one representative per entity is enough
                bool               nc = false;
                const unsigned int pd = dh.get_fe().degree, rep[3] = {0, 1, pd};
                for (unsigned int a = 0; a < 27; ++a)
                  if (pd > 1 || (a % 3 != 1 && (a / 3) % 3 != 1 && a / 9 != 1))
                    nc |= dh.owner[dh.cell_node(c, rep[a % 3], rep[(a / 3) % 3], rep[a / 9])] != rank;
                needs_comm[c - c0] = nc;
              }
          });
        for (std::uint64_t c = c0; c < c1; ++c)
          (needs_comm[c - c0] ? comm : inner).push_back(c);
      }
      const unsigned int L = ad.n_lanes;
      cell_order.clear();
      batch_start.assign(1, 0);
      task_info.partition_row_index.assign(1, 0);
      task_info.cell_partition_data.assign(1, 0);
      auto add_partition = [&](const std::uint64_t *first, const std::uint64_t *last) {
        const unsigned int b0 = (unsigned int)batch_start.size() - 1;
        for (const std::uint64_t *c = first; c < last; c += L)
          {
            const std::uint64_t *e = std::min(c + L, last);
            cell_order.insert(cell_order.end(), c, e);
            batch_start.push_back((unsigned int)cell_order.size());
          }
        const unsigned int b1 = (unsigned int)batch_start.size() - 1;
        for (unsigned int b = b0; b < b1; b += ad.batches_per_range)
          task_info.cell_partition_data.push_back(std::min(b + ad.batches_per_range, b1));
        task_info.partition_row_index.push_back((unsigned int)task_info.cell_partition_data.size() - 1);
      };
      if (tria.n_ranks > 1)
        {
          const std::size_t n_before = ((inner.size() / L) / 2) * L;
          add_partition(inner.data(), inner.data() + n_before);
          add_partition(comm.data(), comm.data() + comm.size());
          add_partition(inner.data() + n_before, inner.data() + inner.size());
          n_cells_before_comm = n_before;
          n_cells_comm        = comm.size();
        }
      else
        add_partition(inner.data(), inner.data() + inner.size());
      // deal.II appends the ghost-face partitions; Renumber loops over size() - 2
      task_info.partition_row_index.push_back(task_info.partition_row_index.back());
      // inverse maps of the loop order: position, batch and range of every local cell
      first_cell = c0;
      cell_pos.assign(c1 - c0, 0);
      batch_of_pos.assign(cell_order.size(), 0);
      range_of_pos.assign(cell_order.size(), 0);
      for (std::uint32_t i = 0; i < cell_order.size(); ++i)
        cell_pos[cell_order[i] - c0] = i;
      for (unsigned int b = 0; b + 1 < batch_start.size(); ++b)
        for (unsigned int i = batch_start[b]; i < batch_start[b + 1]; ++i)
          batch_of_pos[i] = b;
      for (unsigned int r = 0; r + 1 < task_info.cell_partition_data.size(); ++r)
        for (unsigned int b = task_info.cell_partition_data[r]; b < task_info.cell_partition_data[r + 1]; ++b)
          for (unsigned int i = batch_start[b]; i < batch_start[b + 1]; ++i)
            range_of_pos[i] = r;

      if (order_only)
        return;

      // vector partitioner: owned range + ghost nodes sorted by (current) global number
      auto part   = std::make_shared<Utilities::MPI::Partitioner>();
      part->owned = dh.locally_owned_dofs(rank);
      if (tria.n_ranks > 1)
        {
          for (const std::uint64_t c : comm)
            dh.for_each_cell_node(c, [&](std::uint64_t node, int, int, int) {
              if (dh.owner[node] != rank)
                part->ghost_nodes.push_back(dh.node_number[node]);
            });
          std::sort(part->ghost_nodes.begin(), part->ghost_nodes.end());
          part->ghost_nodes.erase(std::unique(part->ghost_nodes.begin(), part->ghost_nodes.end()),
                                  part->ghost_nodes.end());
        }
      // point-to-point plan.  Ghosts are sorted by global number and every rank owns a contiguous
      // global range, so the ghosts owned by one peer are one contiguous block; a peer ghosts
      // exactly those of my nodes that one of its cells touches.  Both sides derive the same
      // sets in the same (global number) order, so send and receive counts match by construction.
      if (tria.n_ranks > 1)
        {
          std::vector<std::vector<std::uint32_t>> exports(tria.n_ranks); // owned local node indices
          std::vector<std::uint64_t>              imports(tria.n_ranks, 0);
          for (const std::uint32_t g : part->ghost_nodes)
            {
              const unsigned int o = (unsigned int)(std::upper_bound(dh.rank_offset.begin(), dh.rank_offset.end(),
                                                                     (std::uint64_t)g) - dh.rank_offset.begin()) - 1;
              ++imports[o];
            }
          const std::uint64_t first = dh.rank_offset[rank];
          // my shared owned nodes and the other ranks among the (up to 8) cells touching them
          // (row by row of the rank's box; the lists are sorted below, so the order of the
          // insertions does not matter)
          std::mutex merge;
          dh.parallel_rank_nodes(rank, [&](const std::uint64_t n0, const std::uint64_t n1) {
            for (std::uint64_t node = n0; node < n1; ++node)
              {
                if (dh.owner[node] != rank || !dh.shared[node])
                  continue;
                const std::uint32_t ln = dh.node_number[node] - (std::uint32_t)first;
                std::uint64_t       cells[8];
                const unsigned int  ncell = dh.incident_cells(node, cells);
                unsigned int        ranks[8], nr = 0;
                for (unsigned int q = 0; q < ncell; ++q)
                  {
                    const unsigned int r   = tria.subdomain_id(cells[q]);
                    bool               dup = r == rank;
                    for (unsigned int k = 0; k < nr; ++k)
                      dup |= ranks[k] == r;
                    if (!dup)
                      ranks[nr++] = r;
                  }
                std::lock_guard<std::mutex> lock(merge);
                for (unsigned int k = 0; k < nr; ++k)
                  exports[ranks[k]].push_back(ln);
              }
          });
          part->import_offset.assign(1, 0);
          part->export_offset.assign(1, 0);
          for (unsigned int r = 0; r < tria.n_ranks; ++r)
            if (imports[r] || !exports[r].empty())
              {
                std::sort(exports[r].begin(), exports[r].end());
                part->peers.push_back((int)r);
                part->import_offset.push_back(part->import_offset.back() + 3 * imports[r]);
                for (const std::uint32_t ln : exports[r])
                  for (unsigned int c = 0; c < 3; ++c)
                    part->export_index.push_back(3 * ln + c);
                part->export_offset.push_back(part->export_index.size());
              }
        }
      dof_info.vector_partitioner = part;

      // owned constrained DoFs, local indices ascending (get_constrained_dofs)
      constrained_dofs.clear();
      {
        std::vector<std::uint32_t> nodes;
        std::mutex                 merge;
        dh.parallel_rank_nodes(rank, [&](const std::uint64_t n0, const std::uint64_t n1) {
          std::uint32_t row[64];
          unsigned int  nr = 0;
          auto          flush = [&] {
            std::lock_guard<std::mutex> lock(merge);
            nodes.insert(nodes.end(), row, row + nr);
            nr = 0;
          };
          for (std::uint64_t node = n0; node < n1; ++node)
            if (dh.owner[node] == rank && con.node_is_constrained(node))
              {
                row[nr++] = dh.node_number[node];
                if (nr == 64)
                  flush();
              }
          if (nr)
            flush();
        });
        std::sort(nodes.begin(), nodes.end());
        const std::uint64_t first = dh.rank_offset[rank];
        for (const std::uint32_t n : nodes)
          for (unsigned int c = 0; c < 3; ++c)
            constrained_dofs.push_back((unsigned int)(3 * (n - first) + c));
      }
    }

    unsigned int n_cell_batches() const { return (unsigned int)batch_start.size() - 1; }
    unsigned int n_active_entries_per_cell_batch(unsigned int b) const
    {
      return batch_start[b + 1] - batch_start[b];
    }
    unsigned int  n_physical_cells() const { return (unsigned int)cell_order.size(); }
    // active-cell index of lane l of batch b (get_cell_iterator(b, l))
    std::uint64_t get_cell(unsigned int b, unsigned int l) const { return cell_order[batch_start[b] + l]; }
    const TaskInfo                  &get_task_info() const { return task_info; }
    const DoFInfo                   &get_dof_info() const { return dof_info; }
    const DoFHandler                &get_dof_handler() const { return *dof_handler; }
    const AffineConstraints         &get_constraints() const { return *constraints; }
    const std::vector<unsigned int> &get_constrained_dofs() const { return constrained_dofs; }
    unsigned int                     get_rank() const { return rank; }
    const AdditionalData            &get_additional_data() const { return data; }

    // the local cells touching a lattice node, as positions in the loop order
    unsigned int incident_positions(const std::uint64_t node, std::uint32_t (&pos)[8]) const
    {
      std::uint64_t      cells[8];
      const unsigned int nc = dof_handler->incident_cells(node, cells);
      unsigned int       n  = 0;
      for (unsigned int k = 0; k < nc; ++k)
        if (cells[k] >= first_cell && cells[k] - first_cell < cell_pos.size())
          pos[n++] = cell_pos[cells[k] - first_cell];
      return n;
    }

    std::uint64_t              first_cell = 0;
    std::vector<std::uint32_t> cell_pos;     // loop position of local cell (active index - first_cell)
    std::vector<std::uint32_t> batch_of_pos; // cell batch of every loop position
    std::vector<std::uint32_t> range_of_pos; // cell-batch range of every loop position
    std::vector<std::uint64_t> cell_order;  // active-cell index per physical cell, loop order
    std::vector<unsigned int>  batch_start; // first physical cell of every batch (+ sentinel)
    unsigned int               n_q_points_1d = 0;
    std::uint64_t              n_cells_before_comm = 0, n_cells_comm = 0; // physical cells per partition

  private:
    const DoFHandler         *dof_handler = nullptr;
    const AffineConstraints  *constraints = nullptr;
    AdditionalData            data;
    TaskInfo                  task_info;
    DoFInfo                   dof_info;
    std::vector<unsigned int> constrained_dofs;
    unsigned int              rank = 0;
  };
} // namespace dealii
