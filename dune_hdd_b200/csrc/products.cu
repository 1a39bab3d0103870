// Product matrices (discretizations/swipdg.hh:359-508, block-swipdg.hh:392-548, base.hh:272-291),
// their application u^T P(mu) v on the device, and the error norms of the convergence studies (8f rank 3:
// test/linearelliptic-swipdg.hh:267-290).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "handles.hpp"

using namespace hdd;

namespace {

const char* const kProductIds[] = {"l2", "h1_semi", "elliptic", "boundary_l2", "penalty", "energy"};

int product_index(const std::string& id) {
  for (int k = 0; k < 6; ++k)
    if (id == kProductIds[k]) return k;
  return -1;
}

Product* find_product(hdd_swipdg* h, const char* id_c) {
  // get_product() (discretizations/base.hh:281-291)
  if (h->products.empty())
    HDD_THROW(HDD_ERR_USING_THIS_WRONG, "Do not call get_product() if available_products() is empty!");
  const std::string id = id_c ? id_c : "";
  for (auto& p : h->products)
    if (p->id == id) return p.get();
  HDD_THROW(HDD_ERR_WRONG_INPUT, id);
}

void assemble_part(hdd_swipdg* h, Product& P, MatrixPart& part) {
  const MeshView v = h->view();
  cudaStream_t s = h->mesh->stream;
  const size_t count = P.volume_pattern ? size_t(h->n_rows) * h->nl : size_t(h->nnz);
  part.values.alloc(count);
  static const DevFn one = [] { DevFn d{}; d.kind = HDD_FN_CONSTANT; d.value = 1.0; return d; }();
  switch (P.which) {
    case 0: case 1: case 3: launch_assemble_block_product(v, P.which, one, 0, h->polorder, part.values.p, s); break;
    case 2: launch_assemble_block_product(v, 2, h->fn_h(part.factor), part.factor.order, h->polorder, part.values.p, s); break;
    case 4: launch_assemble_penalty(v, h->fn_h(part.factor), part.factor.kind, part.factor.order, h->polorder, part.values.p, s); break;
    default: HDD_THROW(HDD_ERR_INTERNAL, "unknown product " << P.which);
  }
}

void assemble_product(hdd_swipdg* h, Product& P) {
  if (P.assembled) return;
  if (P.which == 5) {  // "energy": a copy of the system matrix (discretizations/swipdg.hh:505-508) - shares its parts
    P.assembled = true;
    return;
  }
  for (auto& part : P.comps) assemble_part(h, P, part);
  if (P.affine) assemble_part(h, P, *P.affine);
  P.assembled = true;
}

// parts of a product as (theta, values) lists; energy aliases the system matrix
struct PartList {
  int n = 0;
  const Program* prog[kMaxParts] = {};
  const double* values[kMaxParts] = {};
  const std::string* expr[kMaxParts] = {};
  bool is_affine[kMaxParts] = {};
};

PartList parts_of(hdd_swipdg* h, Product& P) {
  PartList L;
  auto add = [&](MatrixPart& mp, bool affine) {
    L.prog[L.n] = &mp.coef_prog;
    L.values[L.n] = mp.values.p;
    L.expr[L.n] = &mp.coef_expr;
    L.is_affine[L.n++] = affine;
  };
  if (P.which == 5) {
    for (auto& mp : h->lhs_comps) add(mp, false);
    if (h->lhs_affine) add(*h->lhs_affine, true);
  } else {
    for (auto& mp : P.comps) add(mp, false);
    if (P.affine) add(*P.affine, true);
  }
  return L;
}

void ensure_volume_pattern(hdd_swipdg* h) {
  if (h->vol_rowptr.p) return;
  h->vol_rowptr.alloc(size_t(h->n_rows) + 1);
  h->vol_col.alloc(size_t(h->n_rows) * h->nl);
  launch_fill_volume_csr(h->view(), h->vol_rowptr.p, h->vol_col.p, h->mesh->stream);
}

}  // namespace

namespace hdd {
void assemble_products(hdd_swipdg* h) {
  for (auto& p : h->products) assemble_product(h, *p);
}
}  // namespace hdd

extern "C" {

int hdd_swipdg_only_these_products(hdd_swipdg* h, const char* const* ids, int n) {
  return guarded([&] {
    if (!h) HDD_THROW(HDD_ERR_WRONG_INPUT, "discretization handle is NULL");
    if (n < 0 || (n > 0 && !ids)) HDD_THROW(HDD_ERR_WRONG_INPUT, "bad product list");
    h->mesh->set_device();
    for (int k = 0; k < n; ++k) {
      const std::string id = ids[k] ? ids[k] : "";
      const int which = product_index(id);
      if (which < 0) continue;  // unknown ids are ignored, like the std::find tests of discretizations/swipdg.hh:364-505
      bool have = false;
      for (auto& p : h->products) have |= p->id == id;
      if (have) continue;
      std::unique_ptr<Product> P(new Product);
      P->id = id;
      P->which = which;
      P->volume_pattern = which <= 3;
      if (which == 2 || which == 4) {  // one part per diffusion-factor part (:398-425, :444-479)
        for (size_t q = 0; q < h->factor.comps.size(); ++q) {
          MatrixPart mp;
          mp.coef_expr = h->factor.coef_expr[q];
          mp.coef_prog = h->factor.coef_prog[q];
          mp.factor = h->factor.comps[q];
          P->comps.push_back(std::move(mp));
        }
        if (h->factor.has_affine()) {
          P->affine.reset(new MatrixPart);
          P->affine->factor = h->factor.affine;
        }
      } else if (which != 5) {
        P->affine.reset(new MatrixPart);
      }
      h->products.push_back(std::move(P));
    }
    std::sort(h->products.begin(), h->products.end(),
              [](const std::unique_ptr<Product>& a, const std::unique_ptr<Product>& b) { return a->id < b->id; });  // std::map order
    h->product_ids.clear();
    for (auto& p : h->products) h->product_ids.push_back(p->id.c_str());
    if (h->initialized) {
      for (auto& p : h->products) assemble_product(h, *p);
      HDD_CUDA(cudaStreamSynchronize(h->mesh->stream));
    }
  });
}

int hdd_products_available(const hdd_swipdg* h, const char* const** ids, int* n) {
  return guarded([&] {
    if (!h) HDD_THROW(HDD_ERR_WRONG_INPUT, "discretization handle is NULL");
    if (ids) *ids = h->product_ids.data();
    if (n) *n = int(h->product_ids.size());
  });
}

int hdd_product_num_components(hdd_swipdg* h, const char* id, int* n_components, int* has_affine_part, int* volume_pattern) {
  return guarded([&] {
    require_init(h);
    Product* P = find_product(h, id);
    const PartList L = parts_of(h, *P);
    int nc = 0, aff = 0;
    for (int k = 0; k < L.n; ++k) (L.is_affine[k] ? aff : nc) += 1;
    if (n_components) *n_components = nc;
    if (has_affine_part) *has_affine_part = aff;
    if (volume_pattern) *volume_pattern = P->volume_pattern ? 1 : 0;
  });
}

int hdd_product_values(hdd_swipdg* h, const char* id, int q, const double** values_dev, int64_t* count) {
  return guarded([&] {
    require_init(h);
    Product* P = find_product(h, id);
    h->mesh->set_device();
    assemble_product(h, *P);
    const PartList L = parts_of(h, *P);
    const double* p = nullptr;
    int comp = 0;
    for (int k = 0; k < L.n; ++k) {
      if (q == -1 && L.is_affine[k]) p = L.values[k];
      if (!L.is_affine[k]) {
        if (comp == q) p = L.values[k];
        ++comp;
      }
    }
    if (!p) {
      if (q == -1) HDD_THROW(HDD_ERR_USING_THIS_WRONG, "product '" << P->id << "' has no affine part");
      HDD_THROW(HDD_ERR_INDEX_OUT_OF_RANGE, "component " << q);
    }
    if (values_dev) *values_dev = p;
    if (count) *count = P->volume_pattern ? h->n_rows * h->nl : h->nnz;
  });
}

int hdd_product_coefficient(hdd_swipdg* h, const char* id, int q, const char** expression) {
  return guarded([&] {
    require_init(h);
    Product* P = find_product(h, id);
    const PartList L = parts_of(h, *P);
    int comp = 0;
    for (int k = 0; k < L.n; ++k)
      if (!L.is_affine[k]) {
        if (comp == q) {
          if (expression) *expression = L.expr[k]->c_str();
          return;
        }
        ++comp;
      }
    HDD_THROW(HDD_ERR_INDEX_OUT_OF_RANGE, "component " << q);
  });
}

int hdd_pattern_volume(hdd_swipdg* h, int64_t* n_rows, int64_t* nnz, const int64_t** rowptr_dev, const int32_t** col_dev) {
  return guarded([&] {
    require_init(h);
    h->mesh->set_device();
    ensure_volume_pattern(h);
    if (n_rows) *n_rows = h->n_rows;
    if (nnz) *nnz = h->n_rows * h->nl;
    if (rowptr_dev) *rowptr_dev = h->vol_rowptr.p;
    if (col_dev) *col_dev = h->vol_col.p;
  });
}

int hdd_product_apply2(hdd_swipdg* h, const char* id, const double* mu, int mu_size, const double* u_host,
                       const double* v_host, double* result) {
  return guarded([&] {
    require_init(h);
    if (!u_host || !v_host || !result) HDD_THROW(HDD_ERR_WRONG_INPUT, "NULL argument");
    Product* P = find_product(h, id);
    hdd_mesh* m = h->mesh;
    m->set_device();
    cudaStream_t s = m->stream;
    assemble_product(h, *P);
    const PartList L = parts_of(h, *P);
    bool parametric = false;
    for (int k = 0; k < L.n; ++k) parametric |= !L.is_affine[k];
    if (parametric) check_mu(h, mu, mu_size, "mu");
    // freeze
    const size_t count = P->volume_pattern ? size_t(h->n_rows) * h->nl : size_t(h->nnz);
    const double* vals = nullptr;
    if (L.n == 1 && L.is_affine[0]) {
      vals = L.values[0];
    } else {
      FreezeArgs a{};
      for (int k = 0; k < L.n; ++k) {
        a.theta[a.n] = L.is_affine[k] ? 1.0 : eval_coef(*L.prog[k], mu, mu_size);
        a.part[a.n++] = L.values[k];
      }
      if (h->prod_frozen.n < count) h->prod_frozen.alloc(count);
      launch_freeze(a, h->prod_frozen.p, int64_t(count), s);
      vals = h->prod_frozen.p;
    }
    const int nl = h->nl;
    const size_t loc = size_t(m->n_loc) * nl, rows = size_t(h->n_rows);
    if (!h->tmp_local.p) { h->tmp_local.alloc(loc); h->tmp_local.zero(s); }
    if (h->prod_tmp.n < 2 * rows) h->prod_tmp.alloc(2 * rows);
    double* y = h->prod_tmp.p;          // P v
    double* u = h->prod_tmp.p + rows;   // u
    HDD_CUDA(h2d_async(h->tmp_local.p + size_t(m->own0) * nl, v_host, rows * sizeof(double), s));
    HDD_CUDA(h2d_async(u, u_host, rows * sizeof(double), s));
    if (P->volume_pattern) {
      launch_block_spmv(h->view(), vals, h->tmp_local.p + size_t(m->own0) * nl, y, s);
    } else {
      m->halo_exchange(h->tmp_local.p, nl);
      launch_spmv(h->view(), vals, h->tmp_local.p, y, s);
    }
    const int n_seg = int(m->seg_sub.size());
    DevBuf<double> seg;
    seg.alloc(size_t(std::max(n_seg, 1)));
    launch_segment_dot(u, y, m->d_seg_ptr.p, n_seg, nl, seg.p, s);
    std::vector<double> hs(size_t(std::max(n_seg, 1)), 0.0);
    HDD_CUDA(cudaMemcpyAsync(hs.data(), seg.p, size_t(n_seg) * sizeof(double), cudaMemcpyDeviceToHost, s));
    HDD_CUDA(cudaStreamSynchronize(s));
    double total = 0.0;
    for (int g = 0; g < n_seg; ++g) total += hs[size_t(g)];
    if (m->world > 1) {
      DevBuf<double> buf;
      buf.upload(&total, 1, s);
      Nccl::get().all_reduce_sum(buf.p, 1, m->comm, s);
      HDD_CUDA(cudaMemcpyAsync(&total, buf.p, sizeof(double), cudaMemcpyDeviceToHost, s));
      HDD_CUDA(cudaStreamSynchronize(s));
    }
    *result = total;
  });
}

int hdd_error_norms(hdd_swipdg* h, const double* u_host, const char* exact, const char* exact_dx, const char* exact_dy,
                    int order, const double* mu, int mu_size, double* out3) {
  return guarded([&] {
    require_init(h);
    if (!exact || !exact_dx || !exact_dy || !out3) HDD_THROW(HDD_ERR_WRONG_INPUT, "NULL argument");
    if (order < 0 || order > 12) HDD_THROW(HDD_ERR_WRONG_INPUT, "unsupported quadrature order " << order);
    hdd_mesh* m = h->mesh;
    m->set_device();
    cudaStream_t s = m->stream;
    if (h->factor.parametric()) check_mu(h, mu, mu_size, "mu");
    auto fn_of = [](const char* expr) {
      DevFn d{};
      d.kind = HDD_FN_EXPRESSION;
      d.prog = compile_expression(expr, "x");
      return d;
    };
    const DevFn fe = fn_of(exact), fx = fn_of(exact_dx), fy = fn_of(exact_dy);
    const DevCombo fac = make_combo(h, h->factor, mu, mu_size);
    const size_t rows = size_t(h->n_rows), n_own = size_t(m->n_own);
    const double* u_dev = nullptr;
    DevBuf<double> u_tmp;
    if (u_host) {
      u_tmp.upload(u_host, rows, s);
      u_dev = u_tmp.p;
    } else {
      if (!h->have_solution) HDD_THROW(HDD_ERR_USING_THIS_WRONG, "no vector given and no solution available");
      u_dev = h->x.p;
    }
    DevBuf<double> cellwise, seg;
    cellwise.alloc(3 * std::max<size_t>(n_own, 1));
    launch_error_norms(h->view(), h->polorder, fe, fx, fy, fac, h->fn_dev.p, order, u_dev, cellwise.p, s);
    const int n_seg = int(m->seg_sub.size());
    seg.alloc(3 * size_t(std::max(n_seg, 1)));
    for (int r = 0; r < 3; ++r)
      launch_segment_sums(cellwise.p + size_t(r) * n_own, m->d_seg_ptr.p, n_seg, seg.p + size_t(r) * n_seg, s);
    std::vector<double> hs(3 * size_t(std::max(n_seg, 1)), 0.0);
    HDD_CUDA(cudaMemcpyAsync(hs.data(), seg.p, 3 * size_t(n_seg) * sizeof(double), cudaMemcpyDeviceToHost, s));
    HDD_CUDA(cudaStreamSynchronize(s));
    double tot[3] = {0.0, 0.0, 0.0};
    for (int r = 0; r < 3; ++r)
      for (int g = 0; g < n_seg; ++g) tot[r] += hs[size_t(r) * n_seg + g];
    if (m->world > 1) {
      DevBuf<double> buf;
      buf.upload(tot, 3, s);
      Nccl::get().all_reduce_sum(buf.p, 3, m->comm, s);
      HDD_CUDA(cudaMemcpyAsync(tot, buf.p, 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
      HDD_CUDA(cudaStreamSynchronize(s));
    }
    for (int r = 0; r < 3; ++r) out3[r] = std::sqrt(tot[r]);
  });
}

int hdd_prolong(hdd_swipdg* coarse, const double* u_coarse_host, hdd_swipdg* fine, const int32_t* father_host,
                double* u_fine_host) {
  return guarded([&] {
    if (!coarse || !fine || !u_coarse_host || !father_host || !u_fine_host) HDD_THROW(HDD_ERR_WRONG_INPUT, "NULL argument");
    hdd_mesh* mc = coarse->mesh;
    hdd_mesh* mf = fine->mesh;
    if (mc->kind != mf->kind) HDD_THROW(HDD_ERR_WRONG_INPUT, "coarse and fine grid have different element types");
    if (!mc->whole || mc->n_own != mc->n_global)
      HDD_THROW(HDD_ERR_REQUIREMENTS_NOT_MET, "the coarse discretization must own its whole grid (the fine one may be distributed)");
    if (mc->device != mf->device) HDD_THROW(HDD_ERR_REQUIREMENTS_NOT_MET, "coarse and fine discretization live on different devices");
    mf->set_device();
    cudaStream_t s = mf->stream;
    HDD_CUDA(cudaStreamSynchronize(mc->stream));  // the coarse geometry is written on the coarse mesh's stream
    const size_t n_own = size_t(mf->n_own), rows_f = n_own * size_t(fine->nl);
    const size_t rows_c = size_t(mc->n_own) * size_t(coarse->nl);
    DevBuf<double> uc, uf;
    DevBuf<int32_t> fa;
    DevBuf<int> flag;
    uc.upload(u_coarse_host, rows_c, s);
    fa.upload(father_host, n_own, s);
    uf.alloc(std::max<size_t>(rows_f, 1));
    flag.alloc(1);
    flag.zero(s);
    launch_prolong(fine->view(), fine->polorder, mc->cgeo.p, mc->n_own, coarse->polorder, fa.p, uc.p, uf.p, flag.p, s);
    int f = 0;
    HDD_CUDA(cudaMemcpyAsync(&f, flag.p, sizeof(f), cudaMemcpyDeviceToHost, s));
    HDD_CUDA(cudaMemcpyAsync(u_fine_host, uf.p, rows_f * sizeof(double), cudaMemcpyDeviceToHost, s));
    HDD_CUDA(cudaStreamSynchronize(s));
    if (f != 0) HDD_THROW(HDD_ERR_INDEX_OUT_OF_RANGE, "a father id is not a cell of the coarse grid");
  });
}

}  // extern "C"
