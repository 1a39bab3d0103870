// Host-side partition logic of the multi-GPU path (no device needed, so it is testable on a CPU-only box):
// each rank owns a contiguous range of subdomain-major cells; its halo is every other cell sharing a vertex with an
// owned cell (the face neighbours the SpMV needs plus the vertex neighbours the Oswald interpolation needs).
// Because ranks own contiguous ranges and halos are sorted by global id, a rank's halo is grouped by owner and the
// send list for a peer - owned cells sharing a vertex with a halo cell owned by that peer, sorted by global id - is
// exactly that peer's receive range.  No index exchange is ever needed.
#include "partition.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <iterator>

namespace hdd {

void compute_halo(int nl, int64_t n_cells, int64_t n_verts, const int32_t* cell_verts, int64_t cell_begin,
                  int64_t cell_end, std::vector<int32_t>& halo_lo, std::vector<int32_t>& halo_hi) {
  halo_lo.clear();
  halo_hi.clear();
  if (cell_begin == 0 && cell_end == n_cells) return;
  // threaded: both passes are full sweeps over the caller's arrays.  Marking stores the same value from every thread.
  std::vector<uint8_t> vmark(size_t(n_verts), 0);
  uint8_t* mark = vmark.data();
  parallel_for(cell_end - cell_begin, [&](int64_t a, int64_t b) {
    for (int64_t c = cell_begin + a; c < cell_begin + b; ++c)
      for (int i = 0; i < nl; ++i) mark[size_t(cell_verts[c * nl + i])] = 1;
  });
  const int nt = worker_count(n_cells);
  std::vector<std::vector<int32_t>> lo, hi;
  lo.resize(size_t(nt));
  hi.resize(size_t(nt));
  parallel_for_indexed(n_cells, nt, [&](int t, int64_t a, int64_t b) {  // contiguous chunks in order => sorted output
    for (int64_t c = a; c < b; ++c) {
      if (c >= cell_begin && c < cell_end) continue;
      bool touch = false;
      for (int i = 0; i < nl; ++i) touch |= mark[size_t(cell_verts[c * nl + i])] != 0;
      if (touch) (c < cell_begin ? lo : hi)[size_t(t)].push_back(int32_t(c));
    }
  });
  for (int t = 0; t < nt; ++t) {
    halo_lo.insert(halo_lo.end(), lo[size_t(t)].begin(), lo[size_t(t)].end());
    halo_hi.insert(halo_hi.end(), hi[size_t(t)].begin(), hi[size_t(t)].end());
  }
}

void compute_halo_local(int nl, const int32_t* cell_verts, const int32_t* cell_neigh, int64_t n_cells, int64_t cell_begin,
                        int64_t cell_end, std::vector<int32_t>& halo_lo, std::vector<int32_t>& halo_hi,
                        std::vector<int32_t>& boundary_owned) {
  halo_lo.clear();
  halo_hi.clear();
  boundary_owned.clear();
  if (cell_begin == 0 && cell_end == n_cells) return;
  static const int fv3[3][2] = {{0, 1}, {0, 2}, {1, 2}};
  static const int fv4[4][2] = {{0, 2}, {1, 3}, {0, 1}, {2, 3}};
  const int nf = nl;
  auto owned = [&](int64_t c) { return c >= cell_begin && c < cell_end; };
  // partition faces: threaded sweep over the owned cells only
  const int64_t n_own = cell_end - cell_begin;
  const int nt = worker_count(n_own);
  std::vector<std::vector<int32_t>> t_verts, t_seed_far, t_seed_near;
  t_verts.resize(size_t(nt));
  t_seed_far.resize(size_t(nt));
  t_seed_near.resize(size_t(nt));
  parallel_for_indexed(n_own, nt, [&](int t, int64_t a, int64_t b) {
    for (int64_t c = cell_begin + a; c < cell_begin + b; ++c)
      for (int f = 0; f < nf; ++f) {
        const int32_t g = cell_neigh[c * nf + f];
        if (g < 0 || g >= n_cells || owned(g)) continue;
        const int* fv = nl == 3 ? fv3[f] : fv4[f];
        t_verts[size_t(t)].push_back(cell_verts[c * nl + fv[0]]);
        t_verts[size_t(t)].push_back(cell_verts[c * nl + fv[1]]);
        t_seed_far[size_t(t)].push_back(g);
        t_seed_near[size_t(t)].push_back(int32_t(c));
      }
  });
  std::vector<int32_t> pverts, far, near;
  for (int t = 0; t < nt; ++t) {
    pverts.insert(pverts.end(), t_verts[size_t(t)].begin(), t_verts[size_t(t)].end());
    far.insert(far.end(), t_seed_far[size_t(t)].begin(), t_seed_far[size_t(t)].end());
    near.insert(near.end(), t_seed_near[size_t(t)].begin(), t_seed_near[size_t(t)].end());
  }
  std::sort(pverts.begin(), pverts.end());
  pverts.erase(std::unique(pverts.begin(), pverts.end()), pverts.end());
  auto touches = [&](int64_t c) {
    for (int i = 0; i < nl; ++i)
      if (std::binary_search(pverts.begin(), pverts.end(), cell_verts[c * nl + i])) return true;
    return false;
  };
  // walk: from the seeds through face neighbours on the same side of the partition, as long as the cells touch a
  // partition vertex
  auto walk = [&](std::vector<int32_t>& seeds, bool want_owned, std::vector<int32_t>& found) {
    std::sort(seeds.begin(), seeds.end());
    seeds.erase(std::unique(seeds.begin(), seeds.end()), seeds.end());
    std::vector<int32_t> visited(seeds), frontier(seeds);
    found = seeds;  // a seed shares a whole partition face
    while (!frontier.empty()) {
      std::vector<int32_t> next;
      for (int32_t c : frontier)
        for (int f = 0; f < nf; ++f) {
          const int32_t g = cell_neigh[int64_t(c) * nf + f];
          if (g < 0 || g >= n_cells || owned(g) != want_owned) continue;
          next.push_back(g);
        }
      std::sort(next.begin(), next.end());
      next.erase(std::unique(next.begin(), next.end()), next.end());
      std::vector<int32_t> fresh;
      std::set_difference(next.begin(), next.end(), visited.begin(), visited.end(), std::back_inserter(fresh));
      std::vector<int32_t> merged;
      std::merge(visited.begin(), visited.end(), fresh.begin(), fresh.end(), std::back_inserter(merged));
      visited.swap(merged);
      frontier.clear();
      for (int32_t c : fresh)
        if (touches(c)) {
          frontier.push_back(c);
          found.push_back(c);
        }
    }
    std::sort(found.begin(), found.end());
  };
  std::vector<int32_t> halo;
  walk(far, false, halo);
  walk(near, true, boundary_owned);
  for (int32_t c : halo) (c < cell_begin ? halo_lo : halo_hi).push_back(c);
}

int owner_of(const std::vector<int64_t>& rank_cell_offsets, int64_t g) {
  return int(std::upper_bound(rank_cell_offsets.begin(), rank_cell_offsets.end(), g) - rank_cell_offsets.begin()) - 1;
}

void compute_send_cells(int nl, int64_t n_verts, const int32_t* cell_verts, int64_t own_begin, int64_t own_end,
                        const std::vector<int32_t>& halo_cells, const std::vector<int>& halo_owner,
                        std::map<int, std::vector<int32_t>>& send_cells) {
  send_cells.clear();
  std::map<int, std::vector<int32_t>> halo_by_owner;
  for (size_t k = 0; k < halo_cells.size(); ++k) halo_by_owner[halo_owner[k]].push_back(halo_cells[k]);
  for (auto& kv : halo_by_owner) {
    std::vector<uint8_t> vmark(size_t(n_verts), 0);
    for (int32_t g : kv.second)
      for (int i = 0; i < nl; ++i) vmark[size_t(cell_verts[int64_t(g) * nl + i])] = 1;
    std::vector<int32_t>& out = send_cells[kv.first];
    const int nt = worker_count(own_end - own_begin);
    std::vector<std::vector<int32_t>> part;
    part.resize(size_t(nt));
    parallel_for_indexed(own_end - own_begin, nt, [&](int t, int64_t a, int64_t b) {
      for (int64_t c = own_begin + a; c < own_begin + b; ++c) {
        bool touch = false;
        for (int i = 0; i < nl; ++i) touch |= vmark[size_t(cell_verts[c * nl + i])] != 0;
        if (touch) part[size_t(t)].push_back(int32_t(c));
      }
    });
    for (int t = 0; t < nt; ++t) out.insert(out.end(), part[size_t(t)].begin(), part[size_t(t)].end());
  }
}

void compute_send_cells_local(int nl, const std::vector<int32_t>& bowned_cells, const int32_t* bowned_verts,
                              const std::vector<int32_t>& halo_cells, const int32_t* halo_verts,
                              const std::vector<int>& halo_owner, std::map<int, std::vector<int32_t>>& send_cells) {
  send_cells.clear();
  // (vertex, owner) pairs of the halo cells, sorted: which peers sit around which vertex
  std::vector<std::pair<int32_t, int>> vo;
  vo.reserve(halo_cells.size() * size_t(nl));
  for (size_t k = 0; k < halo_cells.size(); ++k)
    for (int i = 0; i < nl; ++i) vo.emplace_back(halo_verts[k * nl + i], halo_owner[k]);
  std::sort(vo.begin(), vo.end());
  vo.erase(std::unique(vo.begin(), vo.end()), vo.end());
  for (size_t k = 0; k < bowned_cells.size(); ++k) {
    int seen[8];
    int n_seen = 0;
    for (int i = 0; i < nl; ++i) {
      const int32_t v = bowned_verts[k * nl + i];
      auto it = std::lower_bound(vo.begin(), vo.end(), std::make_pair(v, -1));
      for (; it != vo.end() && it->first == v; ++it) {
        bool dup = false;
        for (int q = 0; q < n_seen; ++q) dup |= seen[q] == it->second;
        if (!dup && n_seen < 8) {
          seen[n_seen++] = it->second;
          send_cells[it->second].push_back(bowned_cells[k]);
        }
      }
    }
  }
  for (auto& kv : send_cells) std::sort(kv.second.begin(), kv.second.end());
}

}  // namespace hdd

extern "C" {

int hdd_partition_plan(int kind, int64_t n_cells, int64_t n_verts, const int32_t* cell_verts, int world_size,
                       const int64_t* rank_cell_offsets, int rank, int32_t** halo_cells, int64_t* n_halo,
                       int32_t** send_cells, int64_t* send_offsets) {
  return hdd::guarded([&] {
    if (kind != HDD_SIMPLEX2D && kind != HDD_CUBE2D) HDD_THROW(HDD_ERR_WRONG_INPUT, "unknown element kind " << kind);
    if (!cell_verts || !rank_cell_offsets || !halo_cells || !n_halo || !send_cells || !send_offsets)
      HDD_THROW(HDD_ERR_WRONG_INPUT, "NULL argument");
    if (world_size < 1 || rank < 0 || rank >= world_size) HDD_THROW(HDD_ERR_WRONG_INPUT, "bad rank / world size");
    std::vector<int64_t> off(rank_cell_offsets, rank_cell_offsets + world_size + 1);
    if (off.front() != 0 || off.back() != n_cells || !std::is_sorted(off.begin(), off.end()))
      HDD_THROW(HDD_ERR_WRONG_INPUT, "ranks must own consecutive, gap-free cell ranges in rank order");
    const int nl = kind == HDD_SIMPLEX2D ? 3 : 4;
    std::vector<int32_t> lo, hi;
    hdd::compute_halo(nl, n_cells, n_verts, cell_verts, off[size_t(rank)], off[size_t(rank) + 1], lo, hi);
    std::vector<int32_t> halo(lo);
    halo.insert(halo.end(), hi.begin(), hi.end());
    std::vector<int> owner(halo.size());
    for (size_t k = 0; k < halo.size(); ++k) owner[k] = hdd::owner_of(off, halo[k]);
    std::map<int, std::vector<int32_t>> send;
    hdd::compute_send_cells(nl, n_verts, cell_verts, off[size_t(rank)], off[size_t(rank) + 1], halo, owner, send);
    *n_halo = int64_t(halo.size());
    *halo_cells = static_cast<int32_t*>(std::malloc(std::max<size_t>(halo.size(), 1) * sizeof(int32_t)));
    if (!halo.empty()) std::memcpy(*halo_cells, halo.data(), halo.size() * sizeof(int32_t));
    std::vector<int32_t> flat;
    send_offsets[0] = 0;
    for (int r = 0; r < world_size; ++r) {
      auto it = send.find(r);
      if (it != send.end()) flat.insert(flat.end(), it->second.begin(), it->second.end());
      send_offsets[r + 1] = int64_t(flat.size());
    }
    *send_cells = static_cast<int32_t*>(std::malloc(std::max<size_t>(flat.size(), 1) * sizeof(int32_t)));
    if (!flat.empty()) std::memcpy(*send_cells, flat.data(), flat.size() * sizeof(int32_t));
  });
}

int hdd_partition_plan_local(int kind, int64_t n_cells, const int32_t* cell_verts, const int32_t* cell_neigh, int world_size,
                             const int64_t* rank_cell_offsets, int rank, int32_t** halo_cells, int64_t* n_halo,
                             int32_t** send_cells, int64_t* send_offsets) {
  return hdd::guarded([&] {
    if (kind != HDD_SIMPLEX2D && kind != HDD_CUBE2D) HDD_THROW(HDD_ERR_WRONG_INPUT, "unknown element kind " << kind);
    if (!cell_verts || !cell_neigh || !rank_cell_offsets || !halo_cells || !n_halo || !send_cells || !send_offsets)
      HDD_THROW(HDD_ERR_WRONG_INPUT, "NULL argument");
    if (world_size < 1 || rank < 0 || rank >= world_size) HDD_THROW(HDD_ERR_WRONG_INPUT, "bad rank / world size");
    std::vector<int64_t> off(rank_cell_offsets, rank_cell_offsets + world_size + 1);
    if (off.front() != 0 || off.back() != n_cells || !std::is_sorted(off.begin(), off.end()))
      HDD_THROW(HDD_ERR_WRONG_INPUT, "ranks must own consecutive, gap-free cell ranges in rank order");
    const int nl = kind == HDD_SIMPLEX2D ? 3 : 4;
    std::vector<int32_t> lo, hi, bowned;
    hdd::compute_halo_local(nl, cell_verts, cell_neigh, n_cells, off[size_t(rank)], off[size_t(rank) + 1], lo, hi, bowned);
    std::vector<int32_t> halo(lo);
    halo.insert(halo.end(), hi.begin(), hi.end());
    std::vector<int> owner(halo.size());
    std::vector<int32_t> hv(halo.size() * size_t(nl)), bv(bowned.size() * size_t(nl));
    for (size_t k = 0; k < halo.size(); ++k) {
      owner[k] = hdd::owner_of(off, halo[k]);
      std::memcpy(&hv[k * nl], cell_verts + int64_t(halo[k]) * nl, nl * sizeof(int32_t));
    }
    for (size_t k = 0; k < bowned.size(); ++k) std::memcpy(&bv[k * nl], cell_verts + int64_t(bowned[k]) * nl, nl * sizeof(int32_t));
    std::map<int, std::vector<int32_t>> send;
    hdd::compute_send_cells_local(nl, bowned, bv.data(), halo, hv.data(), owner, send);
    *n_halo = int64_t(halo.size());
    *halo_cells = static_cast<int32_t*>(std::malloc(std::max<size_t>(halo.size(), 1) * sizeof(int32_t)));
    if (!halo.empty()) std::memcpy(*halo_cells, halo.data(), halo.size() * sizeof(int32_t));
    std::vector<int32_t> flat;
    send_offsets[0] = 0;
    for (int r = 0; r < world_size; ++r) {
      auto it = send.find(r);
      if (it != send.end()) flat.insert(flat.end(), it->second.begin(), it->second.end());
      send_offsets[r + 1] = int64_t(flat.size());
    }
    *send_cells = static_cast<int32_t*>(std::malloc(std::max<size_t>(flat.size(), 1) * sizeof(int32_t)));
    if (!flat.empty()) std::memcpy(*send_cells, flat.data(), flat.size() * sizeof(int32_t));
  });
}

int hdd_free(void* p) {
  std::free(p);
  return HDD_OK;
}

}  // extern "C"
