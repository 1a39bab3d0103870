// Host-side partition helpers shared by hdd_mesh_create / hdd_comm_init and the device-free hdd_partition_plan.
#pragma once
#include <cstdint>
#include <map>
#include <vector>

#include "common.hpp"

namespace hdd {

// non-owned cells sharing a vertex with an owned cell, split into global ids below / above the owned range (sorted)
void compute_halo(int nl, int64_t n_cells, int64_t n_verts, const int32_t* cell_verts, int64_t cell_begin,
                  int64_t cell_end, std::vector<int32_t>& halo_lo, std::vector<int32_t>& halo_hi);
// The same halo found from the owned side only (hdd_mesh_create on N > 1 ranks): no sweep over the cells of the other
// ranks.  A vertex shared by an owned and a non-owned cell is an end point of a partition face (a face of an owned cell
// whose neighbour is not owned: the cells around a vertex form a fan, and the fan changes owner across such a face), so the
// halo is reached by walking face neighbours from the far side of the partition faces for as long as the cells touch such
// a vertex.  Work and memory are proportional to the partition boundary.  Also returns the owned cells touching a
// partition vertex (sorted global ids) - the only owned cells a neighbour can keep in its halo.
// nf faces per cell with the Dune face -> vertex numbering of `nl` (3: simplex, 4: cube); cell_neigh: -1 = domain boundary.
void compute_halo_local(int nl, const int32_t* cell_verts, const int32_t* cell_neigh, int64_t n_cells, int64_t cell_begin,
                        int64_t cell_end, std::vector<int32_t>& halo_lo, std::vector<int32_t>& halo_hi,
                        std::vector<int32_t>& boundary_owned);
int owner_of(const std::vector<int64_t>& rank_cell_offsets, int64_t global_cell);
// per peer rank: the boundary owned cells (any consistent numbering, sorted) sharing a vertex with a halo cell owned by
// that peer; bowned_verts / halo_verts: nl vertex ids per listed cell
void compute_send_cells_local(int nl, const std::vector<int32_t>& bowned_cells, const int32_t* bowned_verts,
                              const std::vector<int32_t>& halo_cells, const int32_t* halo_verts,
                              const std::vector<int>& halo_owner, std::map<int, std::vector<int32_t>>& send_cells);
// per peer rank: owned cells [own_begin, own_end) (sorted) whose DoFs that peer holds in its halo.  Cell and vertex
// ids may be global or rank-local, as long as cell_verts, the owned range and halo_cells use the same numbering.
void compute_send_cells(int nl, int64_t n_verts, const int32_t* cell_verts, int64_t own_begin, int64_t own_end,
                        const std::vector<int32_t>& halo_cells, const std::vector<int>& halo_owner,
                        std::map<int, std::vector<int32_t>>& send_cells);

}  // namespace hdd
