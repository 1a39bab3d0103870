// hdd_mesh: localisation of the host grid (owned cells + vertex-adjacent halo, sorted by global id), upload,
// block offsets (K1 part 1), vertex incidence for the Oswald pass, halo-exchange plan and NCCL plumbing.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>

#include "handles.hpp"
#include "partition.hpp"

namespace hdd {

std::atomic<int64_t> g_kernel_launches{0};
std::atomic<int64_t> g_h2d_bytes{0}, g_d2h_bytes{0};

SolvePhases& phase_timer() {
  static SolvePhases t = [] {
    SolvePhases x;
    const char* e = std::getenv("HDD_CG_PHASES");
    x.on = e && e[0] == '1';
    return x;
  }();
  return t;
}

void SolvePhases::report(int rank, int iterations) {
  if (!on || marks.size() < 2) return;
  cudaEventSynchronize(marks.back().second);
  std::vector<std::pair<const char*, double>> sums;
  double total = 0.0;
  for (size_t i = 1; i < marks.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
    total += ms;
    bool found = false;
    for (auto& kv : sums)
      if (std::strcmp(kv.first, marks[i].first) == 0) { kv.second += ms; found = true; break; }
    if (!found) sums.emplace_back(marks[i].first, double(ms));
  }
  std::fprintf(stderr, "[hdd phases] rank %d: %d iterations, %.3f ms\n", rank, iterations, total);
  for (auto& kv : sums)
    std::fprintf(stderr, "[hdd phases] rank %d   %-28s %9.3f ms  %7.1f us/iteration  %5.1f %%\n", rank, kv.first, kv.second,
                 1e3 * kv.second / std::max(iterations, 1), 100.0 * kv.second / total);
  for (auto& m : marks) cudaEventDestroy(m.second);
  marks.clear();
}
static thread_local std::string t_last_error;
void set_last_error(const std::string& msg) { t_last_error = msg; }

// ---- caching device allocator (see common.hpp) ---------------------------------------------------------------------------
namespace {
constexpr double kDevCacheFraction = 0.45;
struct DevCache {
  struct Block { void* p; size_t bytes; int device; uint64_t stamp; };
  std::mutex mu;
  std::vector<Block> free_blocks;
  size_t cached_bytes = 0;
  uint64_t clock = 0;
  bool enabled = [] { const char* e = std::getenv("HDD_DEV_CACHE"); return !(e && e[0] == '0'); }();
};
DevCache& dev_cache() {
  static DevCache* c = new DevCache;  // never destroyed: blocks may be freed during static destruction
  return *c;
}
}  // namespace

void dev_cache_release_all() {
  DevCache& c = dev_cache();
  std::lock_guard<std::mutex> lock(c.mu);
  int cur = 0;
  cudaGetDevice(&cur);
  for (const auto& b : c.free_blocks) {
    cudaSetDevice(b.device);
    cudaFree(b.p);
  }
  cudaSetDevice(cur);
  c.free_blocks.clear();
  c.cached_bytes = 0;
}

void* dev_alloc(size_t bytes) {
  DevCache& c = dev_cache();
  int dev = 0;
  cudaGetDevice(&dev);
  if (c.enabled) {
    std::lock_guard<std::mutex> lock(c.mu);
    for (size_t k = 0; k < c.free_blocks.size(); ++k)
      if (c.free_blocks[k].bytes == bytes && c.free_blocks[k].device == dev) {
        void* p = c.free_blocks[k].p;
        c.cached_bytes -= bytes;
        c.free_blocks.erase(c.free_blocks.begin() + long(k));
        return p;
      }
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {  // make room: give everything cached back to the driver and try once more
    cudaGetLastError();
    dev_cache_release_all();
    e = cudaMalloc(&p, bytes);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    HDD_THROW(HDD_ERR_DEVICE, "cudaMalloc of " << bytes << " bytes failed: " << cudaGetErrorString(e));
  }
  return p;
}

void dev_free(void* p, size_t bytes) {
  if (!p) return;
  DevCache& c = dev_cache();
  if (!c.enabled) {
    cudaFree(p);
    return;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  // The block may be handed to another stream next: what cudaFree does implicitly - wait until nothing on the device uses
  // it any more - is done here explicitly for large blocks (small ones are scratch of calls that synchronise themselves)
  if (bytes >= (size_t(1) << 20)) cudaDeviceSynchronize();
  size_t total = 0, avail = 0;
  cudaMemGetInfo(&avail, &total);
  std::lock_guard<std::mutex> lock(c.mu);
  c.free_blocks.push_back({p, bytes, dev, ++c.clock});
  c.cached_bytes += bytes;
  const size_t cap = size_t(kDevCacheFraction * double(total));
  while (c.cached_bytes > cap && !c.free_blocks.empty()) {
    size_t oldest = 0;
    for (size_t k = 1; k < c.free_blocks.size(); ++k)
      if (c.free_blocks[k].stamp < c.free_blocks[oldest].stamp) oldest = k;
    int cur = 0;
    cudaGetDevice(&cur);
    cudaSetDevice(c.free_blocks[oldest].device);
    cudaFree(c.free_blocks[oldest].p);
    cudaSetDevice(cur);
    c.cached_bytes -= c.free_blocks[oldest].bytes;
    c.free_blocks.erase(c.free_blocks.begin() + long(oldest));
  }
}

// ---- NCCL, resolved at run time ----------------------------------------------------------------------------
enum { F_UID, F_INIT, F_DESTROY, F_ALLREDUCE, F_GSTART, F_GEND, F_SEND, F_RECV, F_ERRSTR, F_ALLGATHER };

Nccl::Nccl() {
  handle_ = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!handle_) return;
  const char* names[] = {"ncclGetUniqueId", "ncclCommInitRank", "ncclCommDestroy", "ncclAllReduce", "ncclGroupStart",
                         "ncclGroupEnd",    "ncclSend",         "ncclRecv",        "ncclGetErrorString",
                         "ncclAllGather"};
  for (int k = 0; k < 10; ++k) {
    fn_[k] = dlsym(handle_, names[k]);
    if (!fn_[k]) { handle_ = nullptr; return; }
  }
}
Nccl& Nccl::get() {
  static Nccl n;
  return n;
}
void Nccl::check(int result, const char* what) {
  if (result != 0) {
    auto errstr = reinterpret_cast<const char* (*)(ncclResult_t)>(fn_[F_ERRSTR]);
    HDD_THROW(HDD_ERR_DEVICE, "NCCL error in " << what << ": " << errstr(ncclResult_t(result)));
  }
}
void Nccl::unique_id(void* id128) {
  if (!available()) HDD_THROW(HDD_ERR_DEVICE, "libnccl.so.2 not found");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  check(reinterpret_cast<ncclResult_t (*)(ncclUniqueId*)>(fn_[F_UID])(static_cast<ncclUniqueId*>(id128)), "ncclGetUniqueId");
}
ncclComm* Nccl::init_rank(const void* id128, int rank, int world) {
  if (!available()) HDD_THROW(HDD_ERR_DEVICE, "libnccl.so.2 not found");
  ncclComm_t c = nullptr;
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  check(reinterpret_cast<ncclResult_t (*)(ncclComm_t*, int, ncclUniqueId, int)>(fn_[F_INIT])(&c, world, id, rank),
        "ncclCommInitRank");
  return c;
}
void Nccl::destroy(ncclComm* c) {
  if (c && available()) reinterpret_cast<ncclResult_t (*)(ncclComm_t)>(fn_[F_DESTROY])(c);
}
void Nccl::all_reduce_sum(double* buf, size_t count, ncclComm* c, cudaStream_t s) {
  check(reinterpret_cast<ncclResult_t (*)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t)>(
            fn_[F_ALLREDUCE])(buf, buf, count, ncclDouble, ncclSum, c, s),
        "ncclAllReduce");
}
void Nccl::all_reduce_min(double* buf, size_t count, ncclComm* c, cudaStream_t s) {
  check(reinterpret_cast<ncclResult_t (*)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t)>(
            fn_[F_ALLREDUCE])(buf, buf, count, ncclDouble, ncclMin, c, s),
        "ncclAllReduce");
}
void Nccl::all_gather_bytes(const void* send, void* recv, size_t bytes_per_rank, ncclComm* c, cudaStream_t s) {
  check(reinterpret_cast<ncclResult_t (*)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t)>(fn_[F_ALLGATHER])(
            send, recv, bytes_per_rank, ncclUint8, c, s),
        "ncclAllGather");
}
void Nccl::group_start() { check(reinterpret_cast<ncclResult_t (*)()>(fn_[F_GSTART])(), "ncclGroupStart"); }
void Nccl::group_end() { check(reinterpret_cast<ncclResult_t (*)()>(fn_[F_GEND])(), "ncclGroupEnd"); }
void Nccl::send(const double* buf, size_t count, int peer, ncclComm* c, cudaStream_t s) {
  check(reinterpret_cast<ncclResult_t (*)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)>(fn_[F_SEND])(
            buf, count, ncclDouble, peer, c, s),
        "ncclSend");
}
void Nccl::recv(double* buf, size_t count, int peer, ncclComm* c, cudaStream_t s) {
  check(reinterpret_cast<ncclResult_t (*)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)>(fn_[F_RECV])(
            buf, count, ncclDouble, peer, c, s),
        "ncclRecv");
}

namespace {

const int kFaceVertsSimplex[3][2] = {{0, 1}, {0, 2}, {1, 2}};
const int kFaceVertsCube[4][2] = {{0, 2}, {1, 3}, {0, 1}, {2, 3}};

// exact diameter of a point set: convex hull (monotone chain) + all pairs on the hull.
// Equals the all-pairs vertex loop of LocalResidualOS2014::finalize (estimators/block-swipdg.hh:294-303).
double point_set_diameter(std::vector<std::pair<double, double>>& pts) {
  std::sort(pts.begin(), pts.end());
  pts.erase(std::unique(pts.begin(), pts.end()), pts.end());
  const size_t n = pts.size();
  if (n < 2) return 0.0;
  std::vector<std::pair<double, double>> h(2 * n);
  size_t k = 0;
  auto cross = [](const std::pair<double, double>& o, const std::pair<double, double>& a,
                  const std::pair<double, double>& b) {
    return (a.first - o.first) * (b.second - o.second) - (a.second - o.second) * (b.first - o.first);
  };
  for (size_t i = 0; i < n; ++i) {
    while (k >= 2 && cross(h[k - 2], h[k - 1], pts[i]) <= 0) --k;
    h[k++] = pts[i];
  }
  for (size_t i = n - 1, t = k + 1; i > 0; --i) {
    while (k >= t && cross(h[k - 2], h[k - 1], pts[i - 1]) <= 0) --k;
    h[k++] = pts[i - 1];
  }
  h.resize(k > 1 ? k - 1 : k);
  double d = 0.0;
  for (size_t i = 0; i < h.size(); ++i)
    for (size_t j = i + 1; j < h.size(); ++j)
      d = std::max(d, std::hypot(h[i].first - h[j].first, h[i].second - h[j].second));
  return d;
}

}  // namespace
}  // namespace hdd

using namespace hdd;

hdd_mesh::~hdd_mesh() {
  if (stream) cudaStreamDestroy(stream);  // the communicator belongs to its hdd_comm
}

void hdd_mesh::halo_exchange(double* v_local, int nd) {
  if (world == 1 || peers.empty()) return;
  if (send_buf_capacity < n_send_cells * nd) {
    HDD_CUDA(cudaStreamSynchronize(stream));
    send_buf.release();
    send_buf.alloc(size_t(n_send_cells) * nd);
    send_buf_capacity = n_send_cells * nd;
  }
  launch_pack(v_local, send_idx.p, n_send_cells, nd, send_buf.p, stream);
  Nccl& nc = Nccl::get();
  nc.group_start();
  for (const auto& p : peers) {
    if (p.send_count) nc.send(send_buf.p + p.send_offset * nd, size_t(p.send_count) * nd, p.rank, comm, stream);
    if (p.recv_count) nc.recv(v_local + p.recv_offset * nd, size_t(p.recv_count) * nd, p.rank, comm, stream);
  }
  nc.group_end();
}

// Shared end of hdd_mesh_create / hdd_mesh_create_cube, once n_subdomains / sub_cell_offsets / sub_neighbours are known:
// owned subdomains, their diameters (simplex grids; needs the caller's arrays), the reduction segments and the block
// offsets of the matrix rows (K1 part 1).
static void finish_mesh(hdd_mesh* m, const double* xy, const int32_t* cell_verts, const int32_t* cell_neigh,
                        const int32_t* cell_subdomain, PhaseTimer& pt) {
  const int kind = m->kind, nl = m->nl, nf = m->nf;
  const int64_t cell_begin = m->cell_begin, cell_end = m->cell_end, n_own = cell_end - cell_begin;
  cudaStream_t s = m->stream;
  m->sub_dof_offsets.resize(m->sub_cell_offsets.size());
  for (size_t k = 0; k < m->sub_cell_offsets.size(); ++k) m->sub_dof_offsets[k] = nl * m->sub_cell_offsets[k];
  // owned subdomains: the owned range must consist of whole subdomains
  m->sub_first = int(std::lower_bound(m->sub_cell_offsets.begin(), m->sub_cell_offsets.end(), cell_begin) -
                     m->sub_cell_offsets.begin());
  m->sub_last = int(std::lower_bound(m->sub_cell_offsets.begin(), m->sub_cell_offsets.end(), cell_end) -
                    m->sub_cell_offsets.begin());
  if (n_own > 0 && (m->sub_cell_offsets[size_t(m->sub_first)] != cell_begin ||
                    m->sub_cell_offsets[size_t(m->sub_last)] != cell_end))
    HDD_THROW(HDD_ERR_WRONG_INPUT, "the owned cell range must consist of whole subdomains");
  if (n_own == 0) m->sub_last = m->sub_first;
  m->sub_diameter.assign(size_t(m->n_subdomains), 0.0);
  if (kind == HDD_SIMPLEX2D) {  // only the OS2014 estimators (simplex grids) use the diameters
    for (int sd = m->sub_first; sd < m->sub_last; ++sd) {
      std::vector<std::pair<double, double>> pts;
      // the farthest pair lies on the hull, whose vertices sit on faces leaving the subdomain
      for (int64_t c = m->sub_cell_offsets[size_t(sd)]; c < m->sub_cell_offsets[size_t(sd) + 1]; ++c)
        for (int f = 0; f < nf; ++f) {
          const int32_t g = cell_neigh[c * nf + f];
          if (g >= 0 && cell_subdomain && cell_subdomain[g] == sd) continue;
          if (g >= 0 && !cell_subdomain) continue;
          const int* fv = kFaceVertsSimplex[f];
          for (int e = 0; e < 2; ++e) {
            const int32_t v = cell_verts[c * nl + fv[e]];
            pts.emplace_back(xy[2 * v], xy[2 * v + 1]);
          }
        }
      m->sub_diameter[size_t(sd)] = point_set_diameter(pts);
    }
  }
  // chunked segments of the owned cells for deterministic two-level sums
  {
    const int64_t chunk = 8192;
    m->seg_ptr.clear();
    m->seg_sub.clear();
    for (int sd = m->sub_first; sd < m->sub_last; ++sd) {
      const int64_t b = m->sub_cell_offsets[size_t(sd)] - cell_begin, e = m->sub_cell_offsets[size_t(sd) + 1] - cell_begin;
      for (int64_t k = b; k < e; k += chunk) {
        m->seg_ptr.push_back(k);
        m->seg_sub.push_back(sd);
      }
    }
    m->seg_ptr.push_back(n_own);
    m->d_seg_ptr.upload(m->seg_ptr.data(), m->seg_ptr.size(), s);
  }

  pt.lap("subdomains");
  // ---- K1 part 1: number of blocks per owned cell and their exclusive prefix sum
  {
    DevBuf<int64_t> nblk;
    nblk.alloc(size_t(n_own) + 1);
    nblk.zero(s);
    m->blk_start.alloc(size_t(n_own) + 1);
    const MeshView v = m->view(nullptr);
    launch_count_blocks(v, nblk.p, s);
    exclusive_scan_i64(nblk.p, m->blk_start.p, n_own + 1, s);  // nblk[n_own] = 0 => total in blk_start[n_own]
    HDD_CUDA(cudaMemcpyAsync(&m->n_blocks, m->blk_start.p + n_own, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    HDD_CUDA(cudaStreamSynchronize(s));
  }
  pt.lap("block offsets");
}

extern "C" {

const char* hdd_last_error(void) { return t_last_error.c_str(); }
const char* hdd_version(void) { return "hdd_b200 0.1 (sm_100a)"; }
int64_t hdd_kernel_launches(void) { return g_kernel_launches.load(); }
int64_t hdd_h2d_bytes(void) { return g_h2d_bytes.load(); }

int hdd_mesh_create(int kind, int64_t n_cells, int64_t n_verts, const double* xy, const int32_t* cell_verts,
                    const int32_t* cell_neigh, const int32_t* cell_subdomain, const uint8_t* boundary_type,
                    int64_t cell_begin, int64_t cell_end, int device, hdd_mesh** out) {
  return guarded([&] {
    if (!out) HDD_THROW(HDD_ERR_WRONG_INPUT, "out is NULL");
    *out = nullptr;
    if (kind != HDD_SIMPLEX2D && kind != HDD_CUBE2D) HDD_THROW(HDD_ERR_WRONG_INPUT, "unknown element kind " << kind);
    if (n_cells < 1 || n_verts < 1 || !xy || !cell_verts || !cell_neigh)
      HDD_THROW(HDD_ERR_WRONG_INPUT, "empty grid or NULL array");
    if (n_cells > INT32_MAX / 8) HDD_THROW(HDD_ERR_WRONG_INPUT, "too many cells for 32-bit DoF indices");
    if (cell_begin < 0 || cell_end > n_cells || cell_begin > cell_end)
      HDD_THROW(HDD_ERR_WRONG_INPUT, "bad owned cell range [" << cell_begin << ", " << cell_end << ")");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
      HDD_THROW(HDD_ERR_DEVICE, "no CUDA device available (libhdd_b200 has no CPU fallback)");
    if (device < 0 || device >= n_dev) HDD_THROW(HDD_ERR_DEVICE, "CUDA device " << device << " out of range");

    std::unique_ptr<hdd_mesh> m(new hdd_mesh);
    m->kind = kind;
    m->nl = m->nf = (kind == HDD_SIMPLEX2D ? 3 : 4);
    m->device = device;
    m->n_global = n_cells;
    m->cell_begin = cell_begin;
    m->cell_end = cell_end;
    m->set_device();
    HDD_CUDA(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    const int nl = m->nl, nf = m->nf;
    const bool whole = (cell_begin == 0 && cell_end == n_cells);
    const int64_t n_own = cell_end - cell_begin;
    cudaStream_t s = m->stream;
    PhaseTimer pt("hdd_mesh_create", s);

    // ---- index validation: on the device (inside the geometry / neighbour kernels) for the cells this rank keeps.  A
    // distributed mesh walks the host arrays of its owned cells and of the cells across the partition boundary below, so
    // the ids of the owned cells are checked here first - every rank checks its own, together they check all
    int32_t own_vmin = INT32_MAX, own_vmax = -1;
    if (!whole) {
      std::atomic<int64_t> bad_vertex{-1}, bad_neigh{-1};
      const int nt = worker_count(n_own);
      std::vector<int32_t> mn(size_t(nt), INT32_MAX), mx(size_t(nt), -1);
      parallel_for_indexed(n_own, nt, [&](int t, int64_t c0, int64_t c1) {
        int32_t lo = INT32_MAX, hi = -1;
        for (int64_t c = cell_begin + c0; c < cell_begin + c1; ++c)
          for (int i = 0; i < nl; ++i) {
            const int32_t v = cell_verts[c * nl + i], g = cell_neigh[c * nf + i];
            if (v < 0 || v >= n_verts) bad_vertex = c;
            if (g >= n_cells) bad_neigh = c;
            lo = std::min(lo, v);
            hi = std::max(hi, v);
          }
        mn[size_t(t)] = lo;
        mx[size_t(t)] = hi;
      });
      if (bad_vertex >= 0) HDD_THROW(HDD_ERR_INDEX_OUT_OF_RANGE, "vertex id out of range in cell " << bad_vertex.load());
      if (bad_neigh >= 0) HDD_THROW(HDD_ERR_INDEX_OUT_OF_RANGE, "neighbour id out of range in cell " << bad_neigh.load());
      own_vmin = *std::min_element(mn.begin(), mn.end());
      own_vmax = *std::max_element(mx.begin(), mx.end());
    }
    pt.lap("validate indices");
    // ---- halo: every non-owned cell sharing a vertex with an owned cell (superset of the face neighbours the
    // SpMV needs; the Oswald interpolation needs all cells around a vertex), found from the owned side: work
    // proportional to the partition boundary, no sweep over the other ranks' cells
    std::vector<int32_t> halo_lo, halo_hi, bowned;
    compute_halo_local(nl, cell_verts, cell_neigh, n_cells, cell_begin, cell_end, halo_lo, halo_hi, bowned);
    for (const std::vector<int32_t>* hv : {&halo_lo, &halo_hi})
      for (int32_t g : *hv)
        for (int i = 0; i < nl; ++i) {
          const int32_t v = cell_verts[int64_t(g) * nl + i];
          if (v < 0 || v >= n_verts) HDD_THROW(HDD_ERR_INDEX_OUT_OF_RANGE, "vertex id out of range in cell " << g);
        }
    m->own0 = int32_t(halo_lo.size());
    m->n_own = int32_t(n_own);
    m->n_loc = int32_t(halo_lo.size() + n_own + halo_hi.size());
    m->whole = whole;
    if (!whole) {
      m->cgid = halo_lo;
      m->cgid.insert(m->cgid.end(), halo_hi.begin(), halo_hi.end());
      // what the halo plan of hdd_mesh_attach_comm needs later (the caller's arrays are gone by then): the vertex ids of
      // the halo cells and of the owned cells along the partition boundary
      m->h_halo_verts.resize((halo_lo.size() + halo_hi.size()) * size_t(nl));
      size_t k = 0;
      for (const std::vector<int32_t>* hv : {&halo_lo, &halo_hi})
        for (int32_t g : *hv) {
          std::memcpy(&m->h_halo_verts[k * nl], cell_verts + int64_t(g) * nl, nl * sizeof(int32_t));
          ++k;
        }
      m->h_bowned.resize(bowned.size());
      m->h_bowned_verts.resize(bowned.size() * size_t(nl));
      for (size_t q = 0; q < bowned.size(); ++q) {
        m->h_bowned[q] = int32_t(bowned[q] - cell_begin) + m->own0;  // local id
        std::memcpy(&m->h_bowned_verts[q * nl], cell_verts + int64_t(bowned[q]) * nl, nl * sizeof(int32_t));
      }
    }

    pt.lap("halo + cell ids");
    // ---- the vertices the local cells touch: only the id range [v_begin, v_end) is uploaded (the whole array for a whole
    // mesh; the rows of a strip for the slabs of a structured grid)
    int64_t v_begin = 0, v_end = n_verts;
    if (!whole) {
      int32_t lo = own_vmin, hi = own_vmax;
      for (int32_t v : m->h_halo_verts) {
        lo = std::min(lo, v);
        hi = std::max(hi, v);
      }
      if (hi >= lo) {
        v_begin = lo;
        v_end = int64_t(hi) + 1;
      } else {
        v_begin = 0;
        v_end = 1;
      }
    }
    m->v_begin = v_begin;
    m->v_end = v_end;
    // ---- device-side localisation: the owned slices of the host arrays are uploaded as they are (no host copy);
    // kernels build the per-cell geometry records and translate neighbour ids to local numbering
    {
      DevBuf<double> d_xy_slice;
      d_xy_slice.upload(xy + 2 * v_begin, size_t(2) * (v_end - v_begin), s);
      const double* d_xy = d_xy_slice.p - 2 * v_begin;  // addressed by global vertex id, valid in [v_begin, v_end)
      DevBuf<int32_t> d_cv;
      d_cv.alloc(size_t(m->n_loc) * nl);
      if (!halo_lo.empty())
        HDD_CUDA(h2d_async(d_cv.p, m->h_halo_verts.data(), halo_lo.size() * nl * sizeof(int32_t), s));
      if (n_own)
        HDD_CUDA(h2d_async(d_cv.p + halo_lo.size() * nl, cell_verts + cell_begin * nl, size_t(n_own) * nl * sizeof(int32_t), s));
      if (!halo_hi.empty())
        HDD_CUDA(h2d_async(d_cv.p + (halo_lo.size() + n_own) * nl, m->h_halo_verts.data() + halo_lo.size() * nl,
                                 halo_hi.size() * nl * sizeof(int32_t), s));
      const int ngeo = kind == HDD_SIMPLEX2D ? 6 : 4;
      m->cgeo.alloc(size_t(m->n_loc) * ngeo);
      DevBuf<int32_t> d_flag;
      d_flag.alloc(1);
      d_flag.zero(s);
      launch_build_geometry(kind, m->n_loc, int32_t(v_begin), int32_t(v_end), d_xy, d_cv.p, m->cgeo.p, d_flag.p, s);
      m->neigh.alloc(size_t(n_own) * nf);
      if (n_own)
        HDD_CUDA(h2d_async(m->neigh.p, cell_neigh + cell_begin * nf, size_t(n_own) * nf * sizeof(int32_t), s));
      DevBuf<int32_t> d_halo;
      if (!whole) {
        std::vector<int32_t> halo(halo_lo);
        halo.insert(halo.end(), halo_hi.begin(), halo_hi.end());
        d_halo.upload(halo.data(), halo.size(), s);
        launch_localize_neighbours(m->neigh.p, int64_t(n_own) * nf, int32_t(cell_begin), int32_t(cell_end), d_halo.p,
                                   int32_t(halo_lo.size()), int32_t(halo_hi.size()), d_flag.p, s);
        m->d_cgid.alloc(size_t(m->n_loc));
        launch_iota_from(m->d_cgid.p + halo_lo.size(), int32_t(n_own), int32_t(cell_begin), s);
        if (!halo_lo.empty()) HDD_CUDA(h2d_async(m->d_cgid.p, halo_lo.data(), halo_lo.size() * sizeof(int32_t), s));
        if (!halo_hi.empty())
          HDD_CUDA(h2d_async(m->d_cgid.p + halo_lo.size() + n_own, halo_hi.data(), halo_hi.size() * sizeof(int32_t), s));
      } else {
        launch_validate_neighbours(m->neigh.p, int64_t(n_own) * nf, int32_t(n_cells), d_flag.p, s);
        m->d_cgid.alloc(size_t(m->n_loc));
        launch_iota(m->d_cgid.p, m->n_loc, s);
      }
      mg_detect_structure(m.get(), xy, d_xy, v_begin, v_end, d_cv.p, n_verts);
      if (kind == HDD_SIMPLEX2D && m->lx > 0) m->cell_gv = std::move(d_cv);  // lattice ids of the local cells' vertices
      if (boundary_type) {
        m->has_btype = true;
        m->btype.upload(boundary_type + cell_begin * nf, size_t(n_own) * nf, s);
        std::atomic<int> dirichlet_found{0};
        parallel_for(n_cells, [&](int64_t a, int64_t b) {
          for (int64_t c = a; c < b; ++c)
            for (int f = 0; f < nf; ++f)
              if (cell_neigh[c * nf + f] < 0 && boundary_type[c * nf + f] == 1) dirichlet_found = 1;
        });
        m->purely_neumann = dirichlet_found == 0;
      }
      int32_t flag = 0;
      HDD_CUDA(cudaMemcpyAsync(&flag, d_flag.p, sizeof(flag), cudaMemcpyDeviceToHost, s));
      HDD_CUDA(cudaStreamSynchronize(s));
      if (flag & 4) HDD_THROW(HDD_ERR_INDEX_OUT_OF_RANGE, "a vertex id of a cell is out of range");
      if (flag & 8) HDD_THROW(HDD_ERR_INDEX_OUT_OF_RANGE, "a neighbour id of a cell is out of range");
      if (flag & 1) HDD_THROW(HDD_ERR_NOT_IMPLEMENTED, "HDD_CUBE2D cells must be axis-parallel rectangles");
      if (flag & 2) HDD_THROW(HDD_ERR_INTERNAL, "a neighbour of an owned cell is missing from the halo");
    }

    pt.lap("upload + geometry kernels");
    // ---- local vertices + incidence (vertex -> local DoFs), boundary flags
    // (the incidence feeds the Oswald pass, which exists on simplices only; the local vertex ids also drive the
    // halo plan of a distributed mesh)
    const bool need_verts = kind == HDD_SIMPLEX2D;
    std::vector<int32_t> cvl(need_verts ? size_t(m->n_loc) * nl : 0);
    m->n_verts_loc = int32_t(v_end - v_begin);
    if (need_verts) {
      // global vertex -> local vertex: ascending global id over the vertices of [v_begin, v_end) the local cells touch
      // (threaded mark, serial prefix sum over that range, threaded map)
      std::vector<int32_t> dense(size_t(v_end - v_begin), 0);
      int32_t* dn = dense.data() - v_begin;
      parallel_for(m->n_loc, [&](int64_t a, int64_t b) {
        for (int64_t lc = a; lc < b; ++lc) {
          const int32_t* gv = cell_verts + int64_t(m->gid(int32_t(lc))) * nl;
          for (int i = 0; i < nl; ++i) dn[gv[i]] = 1;  // same value from every thread
        }
      });
      int32_t nvl = 0;
      for (int64_t v = v_begin; v < v_end; ++v) {
        const int32_t touched = dn[v];
        dn[v] = touched ? nvl : -1;
        nvl += touched;
      }
      parallel_for(m->n_loc, [&](int64_t a, int64_t b) {
        for (int64_t lc = a; lc < b; ++lc) {
          const int32_t* gv = cell_verts + int64_t(m->gid(int32_t(lc))) * nl;
          for (int i = 0; i < nl; ++i) cvl[size_t(lc) * nl + i] = dn[gv[i]];
        }
      });
      m->n_verts_loc = nvl;
      if (m->lx > 0) {  // local vertex -> lattice vertex
        std::vector<int32_t> gidv;
        gidv.resize(size_t(nvl));
        for (int64_t v = v_begin; v < v_end; ++v)
          if (dn[v] >= 0) gidv[size_t(dn[v])] = int32_t(v);
        m->lvert_gid.upload(gidv.data(), gidv.size(), s);
        HDD_CUDA(cudaStreamSynchronize(s));
      }
    }
    if (kind == HDD_SIMPLEX2D) {
      const int32_t nvl = m->n_verts_loc;
      std::vector<int64_t> vptr(size_t(nvl) + 1, 0);
      for (int32_t v : cvl) ++vptr[size_t(v) + 1];
      for (int32_t v = 0; v < nvl; ++v) vptr[size_t(v) + 1] += vptr[size_t(v)];
      std::vector<int32_t> vdof(cvl.size());
      std::vector<int64_t> fill(vptr.begin(), vptr.end() - 1);
      for (int32_t lc = 0; lc < m->n_loc; ++lc)
        for (int i = 0; i < nl; ++i) vdof[size_t(fill[size_t(cvl[size_t(lc) * nl + i])]++)] = lc * nl + i;
      std::vector<uint8_t> vb(size_t(nvl), 0);
      for (int32_t lc = 0; lc < m->n_loc; ++lc) {
        const int64_t g = m->gid(lc);
        for (int f = 0; f < nf; ++f)
          if (cell_neigh[g * nf + f] < 0) {
            const int* fv = kFaceVertsSimplex[f];
            vb[size_t(cvl[size_t(lc) * nl + fv[0]])] = 1;
            vb[size_t(cvl[size_t(lc) * nl + fv[1]])] = 1;
          }
      }
      m->vptr.upload(vptr.data(), vptr.size(), s);
      m->vdof.upload(vdof.data(), vdof.size(), s);
      m->vboundary.upload(vb.data(), vb.size(), s);
      m->cell_verts.upload(cvl.data() + size_t(m->own0) * nl, size_t(n_own) * nl, s);
      HDD_CUDA(cudaStreamSynchronize(s));
    }

    pt.lap("vertex incidence");
    // ---- subdomains (grid::Multiscale view): contiguous, subdomain-major cell ranges
    const int n_sub_hint = cell_subdomain ? cell_subdomain[n_cells - 1] + 1 : 1;
    if (cell_subdomain && whole && n_sub_hint >= 1 && n_sub_hint <= 2048) {
      // one GPU owns every cell: offsets and the neighbouring-subdomain relation come from one kernel over the cells
      if (cell_subdomain[0] != 0) HDD_THROW(HDD_ERR_WRONG_INPUT, "subdomain numbering must start at 0");
      const int ns = n_sub_hint;
      m->n_subdomains = ns;
      DevBuf<int32_t> d_sub, d_flag2;
      DevBuf<int64_t> d_off;
      DevBuf<uint8_t> d_adj;
      d_sub.upload(cell_subdomain, size_t(n_cells), s);
      d_off.alloc(size_t(ns) + 1);
      d_off.zero(s);
      d_adj.alloc(size_t(ns) * ns);
      d_adj.zero(s);
      d_flag2.alloc(1);
      d_flag2.zero(s);
      launch_subdomain_structure(d_sub.p, m->neigh.p, nf, int32_t(n_cells), ns, d_off.p, d_adj.p, d_flag2.p, s);
      std::vector<uint8_t> adj(size_t(ns) * ns);
      m->sub_cell_offsets.assign(size_t(ns) + 1, 0);
      int32_t f2 = 0;
      HDD_CUDA(cudaMemcpyAsync(adj.data(), d_adj.p, adj.size(), cudaMemcpyDeviceToHost, s));
      HDD_CUDA(cudaMemcpyAsync(m->sub_cell_offsets.data(), d_off.p, (size_t(ns) + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
      HDD_CUDA(cudaMemcpyAsync(&f2, d_flag2.p, sizeof(f2), cudaMemcpyDeviceToHost, s));
      HDD_CUDA(cudaStreamSynchronize(s));
      m->sub_cell_offsets[size_t(ns)] = n_cells;
      bool empty_sub = false;
      for (int k = 1; k < ns; ++k) empty_sub |= m->sub_cell_offsets[size_t(k)] == 0;
      if (f2 != 0 || empty_sub)
        HDD_THROW(HDD_ERR_WRONG_INPUT, "cells must be numbered subdomain-major without empty subdomains");
      m->sub_neighbours.assign(size_t(ns), {});
      for (int a = 0; a < ns; ++a)
        for (int b = 0; b < ns; ++b)
          if (adj[size_t(a) * ns + b]) m->sub_neighbours[size_t(a)].push_back(b);
    } else if (cell_subdomain) {
      // A rank that owns a part of the cells looks at its own part only: the subdomain-major order is checked over
      // [cell_begin - 1, cell_end] (all ranks together check everything), the offsets of all subdomains come from binary
      // searches in the (monotone) array, the neighbouring-subdomain relation is known for the owned subdomains.
      if (cell_subdomain[0] != 0) HDD_THROW(HDD_ERR_WRONG_INPUT, "subdomain numbering must start at 0");
      std::atomic<int64_t> bad{-1};
      const int64_t lo = std::max<int64_t>(cell_begin - 1, 0), hi = std::min<int64_t>(cell_end, n_cells - 1);
      parallel_for(hi - lo, [&](int64_t a, int64_t b) {
        for (int64_t c = lo + a; c < lo + b; ++c) {
          const int32_t d = cell_subdomain[c + 1] - cell_subdomain[c];
          if (d < 0 || d > 1) bad = c + 1;
        }
      });
      if (bad >= 0)
        HDD_THROW(HDD_ERR_WRONG_INPUT, "cells must be numbered subdomain-major without empty subdomains (cell " << bad.load() << ")");
      m->n_subdomains = cell_subdomain[n_cells - 1] + 1;
      if (m->n_subdomains < 1) HDD_THROW(HDD_ERR_WRONG_INPUT, "bad subdomain numbering");
      m->sub_cell_offsets.assign(size_t(m->n_subdomains) + 1, 0);
      m->sub_cell_offsets[size_t(m->n_subdomains)] = n_cells;
      for (int sd = 1; sd < m->n_subdomains; ++sd) {
        const int32_t* it = std::lower_bound(cell_subdomain, cell_subdomain + n_cells, int32_t(sd));
        if (it == cell_subdomain + n_cells || *it != sd)
          HDD_THROW(HDD_ERR_WRONG_INPUT, "cells must be numbered subdomain-major without empty subdomains (subdomain " << sd << ")");
        m->sub_cell_offsets[size_t(sd)] = it - cell_subdomain;
      }
      // neighbouring subdomains of the owned ones: per-thread pair lists over the owned cells, merged
      const int nt = hdd::worker_count(n_own);
      std::vector<std::vector<std::pair<int32_t, int32_t>>> pairs;
      pairs.resize(size_t(nt));
      parallel_for_indexed(n_own, nt, [&](int t, int64_t a, int64_t b) {
        auto& out = pairs[size_t(t)];
        for (int64_t c = cell_begin + a; c < cell_begin + b; ++c)
          for (int f = 0; f < nf; ++f) {
            const int32_t g = cell_neigh[c * nf + f];
            if (g >= 0 && g < n_cells && cell_subdomain[g] != cell_subdomain[c]) {
              const std::pair<int32_t, int32_t> pr(cell_subdomain[c], cell_subdomain[g]);
              if (out.empty() || out.back() != pr) out.push_back(pr);
            }
          }
      });
      m->sub_neighbours.assign(size_t(m->n_subdomains), {});
      for (auto& v : pairs)
        for (auto& pr : v) m->sub_neighbours[size_t(pr.first)].push_back(pr.second);
      for (auto& v : m->sub_neighbours) {
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
      }
    } else {
      m->n_subdomains = 1;
      m->sub_cell_offsets = {0, n_cells};
      m->sub_neighbours.assign(1, {});
    }
    finish_mesh(m.get(), xy, cell_verts, cell_neigh, cell_subdomain, pt);
    *out = m.release();
  });
}

int hdd_mesh_create_cube(int64_t nx, int64_t ny, double x0, double x1, double y0, double y1, int px, int py,
                         int64_t cell_begin, int64_t cell_end, int device, hdd_mesh** out) {
  return guarded([&] {
    if (!out) HDD_THROW(HDD_ERR_WRONG_INPUT, "out is NULL");
    *out = nullptr;
    if (nx < 1 || ny < 1 || px < 1 || py < 1 || px > nx || py > ny || px * py > 4096)
      HDD_THROW(HDD_ERR_WRONG_INPUT, "bad grid size " << nx << " x " << ny << " or partition " << px << " x " << py);
    if (!(x1 > x0) || !(y1 > y0)) HDD_THROW(HDD_ERR_WRONG_INPUT, "upper_right must exceed lower_left");
    const int64_t n_cells = nx * ny;
    if (n_cells > INT32_MAX / 8) HDD_THROW(HDD_ERR_WRONG_INPUT, "too many cells for 32-bit DoF indices");
    if (cell_end < 0) cell_end = n_cells;
    if (cell_begin < 0 || cell_end > n_cells || cell_begin > cell_end)
      HDD_THROW(HDD_ERR_WRONG_INPUT, "bad owned cell range [" << cell_begin << ", " << cell_end << ")");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
      HDD_THROW(HDD_ERR_DEVICE, "no CUDA device available (libhdd_b200 has no CPU fallback)");
    if (device < 0 || device >= n_dev) HDD_THROW(HDD_ERR_DEVICE, "CUDA device " << device << " out of range");

    std::unique_ptr<hdd_mesh> m(new hdd_mesh);
    m->kind = HDD_CUBE2D;
    m->nl = m->nf = 4;
    m->device = device;
    m->n_global = n_cells;
    m->cell_begin = cell_begin;
    m->cell_end = cell_end;
    m->set_device();
    HDD_CUDA(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    cudaStream_t s = m->stream;
    PhaseTimer pt("hdd_mesh_create_cube", s);
    const bool whole = (cell_begin == 0 && cell_end == n_cells);
    const int64_t n_own = cell_end - cell_begin;
    // boxes: the column / row a box starts at, exactly as the host generator assigns cells to boxes (grids.cpp: box_of)
    auto starts = [](int64_t n, int parts) {
      std::vector<int32_t> st(size_t(parts) + 1, int32_t(n));
      st[0] = 0;
      for (int64_t i = 0; i < n; ++i) {
        int b = int((double(i) + 0.5) / double(n) * parts);
        b = std::min(std::max(b, 0), parts - 1);
        if (int32_t(i) < st[size_t(b)]) st[size_t(b)] = int32_t(i);
      }
      return st;
    };
    const std::vector<int32_t> X = starts(nx, px), Y = starts(ny, py);
    const int ns = px * py;
    std::vector<int64_t> off(size_t(ns) + 1, 0);
    for (int b = 0; b < ns; ++b)
      off[size_t(b) + 1] = off[size_t(b)] + int64_t(X[size_t(b % px) + 1] - X[size_t(b % px)]) * (Y[size_t(b / px) + 1] - Y[size_t(b / px)]);
    for (int b = 0; b < ns; ++b)
      if (off[size_t(b) + 1] == off[size_t(b)]) HDD_THROW(HDD_ERR_WRONG_INPUT, "the partition has an empty box");
    auto box_at = [&](const std::vector<int32_t>& st, int64_t i) { return int(std::upper_bound(st.begin(), st.end(), int32_t(i)) - st.begin()) - 1; };
    auto cell_id = [&](int64_t i, int64_t j) -> int64_t {
      if (i < 0 || j < 0 || i >= nx || j >= ny) return -1;
      const int bx = box_at(X, i), by = box_at(Y, j);
      return off[size_t(by * px + bx)] + (j - Y[size_t(by)]) * int64_t(X[size_t(bx) + 1] - X[size_t(bx)]) + (i - X[size_t(bx)]);
    };
    m->n_subdomains = ns;
    m->sub_cell_offsets = off;
    m->sub_neighbours.assign(size_t(ns), {});
    for (int b = 0; b < ns; ++b) {
      const int bx = b % px, by = b / px;
      if (by > 0) m->sub_neighbours[size_t(b)].push_back(b - px);
      if (bx > 0) m->sub_neighbours[size_t(b)].push_back(b - 1);
      if (bx + 1 < px) m->sub_neighbours[size_t(b)].push_back(b + 1);
      if (by + 1 < py) m->sub_neighbours[size_t(b)].push_back(b + px);
    }
    const int sub_first = int(std::lower_bound(off.begin(), off.end(), cell_begin) - off.begin());
    const int sub_last = int(std::lower_bound(off.begin(), off.end(), cell_end) - off.begin());
    if (n_own > 0 && (off[size_t(sub_first)] != cell_begin || off[size_t(sub_last)] != cell_end))
      HDD_THROW(HDD_ERR_WRONG_INPUT, "the owned cell range must consist of whole subdomains");
    // halo (cells of other ranks sharing a vertex with an owned cell) and the owned cells along the partition boundary:
    // the one-cell rings outside / inside the owned boxes, in closed form - work proportional to the boundary
    std::vector<int32_t> halo, bowned;
    if (!whole) {
      auto owned = [&](int64_t g) { return g >= cell_begin && g < cell_end; };
      for (int b = sub_first; b < sub_last; ++b) {
        const int bx = b % px, by = b / px;
        const int64_t i0 = X[size_t(bx)], i1 = X[size_t(bx) + 1], j0 = Y[size_t(by)], j1 = Y[size_t(by) + 1];
        auto visit = [&](int64_t i, int64_t j) {  // a cell of the outer ring and its inner neighbours
          const int64_t g = cell_id(i, j);
          if (g < 0 || owned(g)) return;
          halo.push_back(int32_t(g));
          for (int dj = -1; dj <= 1; ++dj)
            for (int di = -1; di <= 1; ++di) {
              const int64_t q = cell_id(i + di, j + dj);
              if (q >= 0 && owned(q)) bowned.push_back(int32_t(q));
            }
        };
        for (int64_t i = i0 - 1; i <= i1; ++i) { visit(i, j0 - 1); visit(i, j1); }
        for (int64_t j = j0; j < j1; ++j) { visit(i0 - 1, j); visit(i1, j); }
      }
      std::sort(halo.begin(), halo.end());
      halo.erase(std::unique(halo.begin(), halo.end()), halo.end());
      std::sort(bowned.begin(), bowned.end());
      bowned.erase(std::unique(bowned.begin(), bowned.end()), bowned.end());
    }
    const size_t n_lo = size_t(std::lower_bound(halo.begin(), halo.end(), int32_t(cell_begin)) - halo.begin());
    m->own0 = int32_t(n_lo);
    m->n_own = int32_t(n_own);
    m->n_loc = int32_t(halo.size() + n_own);
    m->whole = whole;
    m->n_verts_loc = int32_t((nx + 1) * (ny + 1));
    if (!whole) {
      m->cgid = halo;
      auto verts_of = [&](int64_t g, int32_t* v4) {  // inverse of cell_id, then the vertex ids of hdd_grid_cube
        const int b = int(std::upper_bound(off.begin(), off.end(), g) - off.begin()) - 1;
        const int bx = b % px, by = b / px;
        const int64_t w = X[size_t(bx) + 1] - X[size_t(bx)], r = g - off[size_t(b)];
        const int64_t i = X[size_t(bx)] + r % w, j = Y[size_t(by)] + r / w;
        v4[0] = int32_t(j * (nx + 1) + i);
        v4[1] = v4[0] + 1;
        v4[2] = int32_t((j + 1) * (nx + 1) + i);
        v4[3] = v4[2] + 1;
      };
      m->h_halo_verts.resize(halo.size() * 4);
      for (size_t k = 0; k < halo.size(); ++k) verts_of(halo[k], &m->h_halo_verts[4 * k]);
      m->h_bowned.resize(bowned.size());
      m->h_bowned_verts.resize(bowned.size() * 4);
      for (size_t k = 0; k < bowned.size(); ++k) {
        m->h_bowned[k] = int32_t(bowned[k] - cell_begin) + m->own0;
        verts_of(bowned[k], &m->h_bowned_verts[4 * k]);
      }
    }
    pt.lap("boxes + halo");
    // ---- everything per cell is written by one kernel
    DevBuf<int32_t> d_X, d_Y, d_halo, d_flag;
    DevBuf<int64_t> d_off;
    d_X.upload(X.data(), X.size(), s);
    d_Y.upload(Y.data(), Y.size(), s);
    d_off.upload(off.data(), off.size(), s);
    d_halo.upload(halo.data(), halo.size(), s);
    d_flag.alloc(1);
    d_flag.zero(s);
    m->cgeo.alloc(size_t(m->n_loc) * 4);
    m->neigh.alloc(size_t(n_own) * 4);
    m->d_cgid.alloc(size_t(m->n_loc));
    m->cell_v0.alloc(size_t(m->n_loc));
    m->lex_cell.alloc(size_t(n_cells));
    HDD_CUDA(cudaMemsetAsync(m->lex_cell.p, 0xFF, size_t(n_cells) * sizeof(int32_t), s));  // -1: not on this rank
    m->tgeo.alloc(4 * size_t(nx + ny));
    CubeGridDesc g{int(nx), int(ny), px, py, x0, x1, y0, y1, d_X.p, d_Y.p, d_off.p};
    launch_cube_fill(g, m->n_loc, m->own0, m->n_own, int32_t(cell_begin), d_halo.p, m->cgeo.p, m->cell_v0.p, m->lex_cell.p,
                     m->d_cgid.p, m->neigh.p, m->tgeo.p, s);
    if (!whole)
      launch_localize_neighbours(m->neigh.p, int64_t(n_own) * 4, int32_t(cell_begin), int32_t(cell_end), d_halo.p, int32_t(n_lo),
                                 int32_t(halo.size() - n_lo), d_flag.p, s);
    m->sx = int(nx);
    m->sy = int(ny);
    int32_t flag = 0;
    HDD_CUDA(cudaMemcpyAsync(&flag, d_flag.p, sizeof(flag), cudaMemcpyDeviceToHost, s));
    HDD_CUDA(cudaStreamSynchronize(s));
    if (flag & 2) HDD_THROW(HDD_ERR_INTERNAL, "a neighbour of an owned cell is missing from the halo");
    pt.lap("fill kernels");
    finish_mesh(m.get(), nullptr, nullptr, nullptr, nullptr, pt);
    *out = m.release();
  });
}

int hdd_host_alloc(size_t bytes, void** ptr) {
  return guarded([&] {
    if (!ptr) HDD_THROW(HDD_ERR_WRONG_INPUT, "ptr is NULL");
    *ptr = nullptr;
    HDD_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable));
  });
}

int hdd_host_free(void* ptr) {
  return guarded([&] {
    if (ptr) HDD_CUDA(cudaFreeHost(ptr));
  });
}

int hdd_mesh_destroy(hdd_mesh* mesh) {
  return guarded([&] {
    if (mesh) {
      cudaSetDevice(mesh->device);
      delete mesh;
    }
  });
}

int hdd_mesh_num_cells(const hdd_mesh* mesh, int64_t* n_global, int64_t* n_owned, int64_t* n_halo) {
  return guarded([&] {
    if (!mesh) HDD_THROW(HDD_ERR_WRONG_INPUT, "mesh is NULL");
    if (n_global) *n_global = mesh->n_global;
    if (n_owned) *n_owned = mesh->n_own;
    if (n_halo) *n_halo = mesh->n_loc - mesh->n_own;
  });
}

int hdd_comm_unique_id(void* id128) {
  return guarded([&] { Nccl::get().unique_id(id128); });
}

int hdd_comm_create(const void* id128, int rank, int world_size, int device, hdd_comm** out) {
  return guarded([&] {
    if (!out || !id128) HDD_THROW(HDD_ERR_WRONG_INPUT, "NULL argument");
    *out = nullptr;
    if (world_size < 1 || rank < 0 || rank >= world_size) HDD_THROW(HDD_ERR_WRONG_INPUT, "bad rank / world size");
    HDD_CUDA(cudaSetDevice(device));
    std::unique_ptr<hdd_comm> c(new hdd_comm);
    c->rank = rank;
    c->world = world_size;
    c->device = device;
    if (world_size > 1) c->comm = Nccl::get().init_rank(id128, rank, world_size);
    *out = c.release();
  });
}

int hdd_comm_destroy(hdd_comm* c) {
  return guarded([&] {
    if (!c) return;
    if (c->comm) Nccl::get().destroy(c->comm);
    delete c;
  });
}

int hdd_mesh_attach_comm(hdd_mesh* m, hdd_comm* c) {
  return guarded([&] {
    if (!m || !c) HDD_THROW(HDD_ERR_WRONG_INPUT, "mesh or comm is NULL");
    if (c->device != m->device) HDD_THROW(HDD_ERR_WRONG_INPUT, "communicator and mesh live on different devices");
    const int rank = c->rank, world_size = c->world;
    m->set_device();
    m->rank = rank;
    m->world = world_size;
    if (world_size == 1) return;
    Nccl& nc = Nccl::get();
    m->comm = c->comm;
    PhaseTimer pt("hdd_mesh_attach", m->stream);
    // every rank learns every rank's owned range (2 doubles per rank through one all-reduce)
    DevBuf<double> ranges;
    ranges.alloc(size_t(world_size) + 1);
    std::vector<double> h(size_t(world_size) + 1, 0.0);
    h[size_t(rank)] = double(m->cell_begin);
    if (rank == world_size - 1) h[size_t(world_size)] = double(m->cell_end);
    ranges.upload(h.data(), h.size(), m->stream);
    nc.all_reduce_sum(ranges.p, h.size(), m->comm, m->stream);
    HDD_CUDA(cudaMemcpyAsync(h.data(), ranges.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    HDD_CUDA(cudaStreamSynchronize(m->stream));
    m->rank_cell_offsets.resize(h.size());
    for (size_t r = 0; r < h.size(); ++r) m->rank_cell_offsets[r] = int64_t(h[r]);
    if (m->rank_cell_offsets[size_t(rank) + 1] != m->cell_end || m->rank_cell_offsets[0] != 0 ||
        m->rank_cell_offsets[size_t(world_size)] != m->n_global)
      HDD_THROW(HDD_ERR_WRONG_INPUT, "ranks must own consecutive, gap-free cell ranges in rank order");
    {  // subdomain diameters are computed by the owning rank only (zero elsewhere): make them global once
      DevBuf<double> dia;
      dia.upload(m->sub_diameter.data(), m->sub_diameter.size(), m->stream);
      nc.all_reduce_sum(dia.p, m->sub_diameter.size(), m->comm, m->stream);
      HDD_CUDA(cudaMemcpyAsync(m->sub_diameter.data(), dia.p, m->sub_diameter.size() * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
      HDD_CUDA(cudaStreamSynchronize(m->stream));
    }
    {  // logically structured grid: every rank checked its own cells and vertices; the verdict has to be the same everywhere
      DevBuf<double> bad;
      double f = (m->sx > 0 || m->lx > 0) ? 0.0 : 1.0;
      bad.upload(&f, 1, m->stream);
      nc.all_reduce_sum(bad.p, 1, m->comm, m->stream);
      HDD_CUDA(cudaMemcpyAsync(&f, bad.p, sizeof(double), cudaMemcpyDeviceToHost, m->stream));
      HDD_CUDA(cudaStreamSynchronize(m->stream));
      if (f != 0.0 && m->lx > 0) {
        m->lx = m->ly = 0;
        m->cell_gv.release();
        m->lvert_gid.release();
      }
      if (f != 0.0 && m->sx > 0) {
        m->sx = m->sy = 0;
        m->cell_v0.release();
        m->lex_cell.release();
        m->tgeo.release();
      }
    }
    pt.lap("ranges + diameters");
    // halo plan.  Receive: halo cells sorted by global id are grouped by owner => contiguous ranges of the local
    // vector.  Send: owned cells sharing a vertex with a halo cell owned by that peer, sorted by global id - this is
    // exactly that peer's receive range, no index exchange needed.
    const int nl = m->nl;
    std::map<int, HaloPeer> peers;
    std::vector<int32_t> halo_cells;
    std::vector<int> halo_owner;
    for (int32_t hk = 0; hk < m->n_loc - m->n_own; ++hk) {
      const int32_t lc = hk < m->own0 ? hk : hk + m->n_own;
      const int r = owner_of(m->rank_cell_offsets, m->gid(lc));
      halo_cells.push_back(lc);
      halo_owner.push_back(r);
      auto it = peers.find(r);
      if (it == peers.end()) {
        HaloPeer p{};
        p.rank = r;
        p.recv_offset = int64_t(lc);
        it = peers.emplace(r, p).first;
      }
      it->second.recv_count += 1;
    }
    std::map<int, std::vector<int32_t>> send_cells;
    compute_send_cells_local(nl, m->h_bowned, m->h_bowned_verts.data(), halo_cells, m->h_halo_verts.data(), halo_owner, send_cells);
    std::vector<int32_t> idx;
    m->peers.clear();
    for (auto& kv : peers) {
      HaloPeer p = kv.second;
      p.send_offset = int64_t(idx.size());
      for (int32_t lc : send_cells[kv.first]) idx.push_back(lc);
      p.send_count = int64_t(idx.size()) - p.send_offset;
      m->peers.push_back(p);
    }
    pt.lap("halo plan");
    m->send_idx.upload(idx.data(), idx.size(), m->stream);
    m->n_send_cells = int64_t(idx.size());
    m->send_buf.alloc(idx.size() * size_t(nl));
    m->send_buf_capacity = int64_t(idx.size()) * nl;
    // tables for the peer-memory SpMV: owner rank and the owner's row offset (in cells) of every halo cell
    {
      std::vector<int32_t> hp(halo_cells.size()), hr(halo_cells.size());
      for (size_t k = 0; k < halo_cells.size(); ++k) {
        hp[k] = halo_owner[k];
        hr[k] = int32_t(m->gid(halo_cells[k]) - m->rank_cell_offsets[size_t(halo_owner[k])]);
      }
      m->halo_peer.upload(hp.data(), hp.size(), m->stream);
      m->halo_rcell.upload(hr.data(), hr.size(), m->stream);
      // every rank's own0 (offset of its owned cells inside its local vectors)
      DevBuf<int32_t> own0_all;
      own0_all.alloc(size_t(world_size));
      DevBuf<int32_t> mine;
      mine.upload(&m->own0, 1, m->stream);
      nc.all_gather_bytes(mine.p, own0_all.p, sizeof(int32_t), m->comm, m->stream);
      m->rank_own0.resize(size_t(world_size));
      HDD_CUDA(cudaMemcpyAsync(m->rank_own0.data(), own0_all.p, size_t(world_size) * sizeof(int32_t), cudaMemcpyDeviceToHost, m->stream));
      HDD_CUDA(cudaStreamSynchronize(m->stream));
    }
    HDD_CUDA(cudaStreamSynchronize(m->stream));
    pt.lap("peer tables");
  });
}

}  // extern "C"
