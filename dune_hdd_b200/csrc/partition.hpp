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
int owner_of(const std::vector<int64_t>& rank_cell_offsets, int64_t global_cell);
// per peer rank: owned cells [own_begin, own_end) (sorted) whose DoFs that peer holds in its halo.  Cell and vertex
// ids may be global or rank-local, as long as cell_verts, the owned range and halo_cells use the same numbering.
void compute_send_cells(int nl, int64_t n_verts, const int32_t* cell_verts, int64_t own_begin, int64_t own_end,
                        const std::vector<int32_t>& halo_cells, const std::vector<int>& halo_owner,
                        std::map<int, std::vector<int32_t>>& send_cells);

}  // namespace hdd
