// K7-K10: ESV2007 / OS2014 local indicators on 2-d simplices for sm_100a.
//
// Two passes instead of the reference's four grid walks (estimators/swipdg.hh:668-687):
//   pass 1 (per vertex)  Oswald value = mean of the DG values around the vertex, 0 on the boundary
//                        (GDT::Operators::OswaldInterpolation, estimators/swipdg.hh:149-150)
//   pass 2 (per cell)    everything else, one thread per element.  The RT0 flux reconstruction
//                        (Operators::DiffusiveFluxReconstruction, estimators/swipdg.hh:590-595) is evaluated
//                        owner-computes: each cell integrates the SWIPDG flux through its own three faces with its
//                        own outward normal - the flux is antisymmetric under swapping the two cells, so no face
//                        array and no face pass are needed - and immediately consumes t_h in eta_DF, eta_DF*, eta_R*.
#include "kernels.hpp"

namespace hdd {

namespace {

__global__ void k_vertex_means(const int64_t* __restrict__ vptr, const int32_t* __restrict__ vdof,
                               const uint8_t* __restrict__ vboundary, int32_t n_verts, const double* __restrict__ u,
                               double* __restrict__ vm) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_verts) return;
  if (vboundary[v]) { vm[v] = 0.0; return; }
  const int64_t b = vptr[v], e = vptr[v + 1];
  double s = 0.0;
  for (int64_t k = b; k < e; ++k) s += u[vdof[k]];
  vm[v] = e > b ? s / double(e - b) : 0.0;
}

constexpr double kPi = 3.14159265358979323846264338327950288;

struct IndicatorRules {
  ElemRule nc, p0, res, df, cut, amin, amax;
  LineRule face;
};

__global__ void __launch_bounds__(128)
    k_indicators(MeshView m, const __grid_constant__ IndicatorArgs a, const IndicatorRules* __restrict__ rules, double s_in,
                 double s_bnd) {
  using G = Geo<HDD_SIMPLEX2D>;
  constexpr int NL = 3;
  // the data functions (expression programs included) are staged in shared memory once per block
  extern __shared__ __align__(16) unsigned char fn_smem[];
  DevFn* table = reinterpret_cast<DevFn*>(fn_smem);
  {
    const int words = a.n_fn * int(sizeof(DevFn) / 4);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.fn_table);
    uint32_t* dst = reinterpret_cast<uint32_t*>(fn_smem);
    for (int w = threadIdx.x; w < words; w += blockDim.x) dst[w] = src[w];
  }
  __syncthreads();
  const DevFn& force = table[a.force_idx];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m.n_own) return;
  const int c = m.own0 + k;
  const IndicatorRules& R = *rules;
  G g;
  g.load(m.cgeo, c);
  double K[4];
  load_tensor(m.tensor, c, K);
  const double tr = K[0] + K[3], det = K[0] * K[3] - K[1] * K[2];
  const double lam_min = 0.5 * tr - sqrt(fmax(0.0, 0.25 * tr * tr - det));
  double phi[NL], gx[NL], gy[NL];
  g.basis(1.0 / 3.0, 1.0 / 3.0, phi, gx, gy);
  double u[NL], ux = 0, uy = 0, dx = 0, dy = 0;
#pragma unroll
  for (int i = 0; i < NL; ++i) {
    u[i] = a.u_local[size_t(NL) * c + i];
    const double iu = a.vertex_mean[a.cell_verts[size_t(NL) * k + i]];
    ux += u[i] * gx[i]; uy += u[i] * gy[i];
    dx += (u[i] - iu) * gx[i]; dy += (u[i] - iu) * gy[i];
  }
  const size_t n = size_t(m.n_own);
  // eta_NC,T^2
  double v_nc, v_r, v_df;
  {
    double s = 0.0;
    const double e = (K[0] * dx + K[1] * dy) * dx + (K[2] * dx + K[3] * dy) * dy;
    for (int q = 0; q < R.nc.n; ++q) {
      double x, y;
      g.to_global(R.nc.x[q], R.nc.y[q], x, y);
      s += R.nc.w[q] * g.detj * combo_eval(a.a_bar, table, c, x, y) * e;
    }
    a.out[0 * n + k] = s;
    v_nc = s;
  }
  // P0 projection of f and int_T (f - P0 f)^2
  double f0 = 0.0;
  for (int q = 0; q < R.p0.n; ++q) {
    double x, y;
    g.to_global(R.p0.x[q], R.p0.y[q], x, y);
    f0 += R.p0.w[q] * fn_eval(force, c, x, y);
  }
  f0 /= 0.5;
  const double hT = g.diameter();
  double cT = 1e300;
  for (int q = 0; q < R.cut.n; ++q) {
    double x, y;
    g.to_global(R.cut.x[q], R.cut.y[q], x, y);
    cT = fmin(cT, combo_eval(a.a_cut, table, c, x, y) * lam_min);
  }
  const double cutoff = hT * hT / (kPi * kPi * cT);
  double f_res[kMaxElemPts];  // force at the residual rule's points: evaluated once, used for eta_R and eta_R*
  {
    double rs = 0.0;
    for (int q = 0; q < R.res.n; ++q) {
      double x, y;
      g.to_global(R.res.x[q], R.res.y[q], x, y);
      f_res[q] = fn_eval(force, c, x, y);
      const double d = f_res[q] - f0;
      rs += R.res.w[q] * g.detj * d * d;
    }
    a.out[1 * n + k] = rs;
    a.out[2 * n + k] = cutoff * rs;
    v_r = cutoff * rs;
  }
  {
    double mn = 1e300;
    for (int q = 0; q < R.amin.n; ++q) {
      double x, y;
      g.to_global(R.amin.x[q], R.amin.y[q], x, y);
      mn = fmin(mn, combo_eval(a.a_min, table, c, x, y));
    }
    for (int q = 0; q < R.amax.n; ++q) {
      double x, y;
      g.to_global(R.amax.x[q], R.amax.y[q], x, y);
      mn = fmin(mn, combo_eval(a.a_max, table, c, x, y));
    }
    a.out[6 * n + k] = mn * lam_min;
  }
  // outward RT0 fluxes G_f = int_f ( -{{A grad u . n}}_omega + pen [[u]] )  resp. ( -A grad u . n + pen u )
  double Gf[3];
#pragma unroll 1
  for (int f = 0; f < 3; ++f) {
    const FaceGeo e = make_face(g, f);
    const int nbc = m.neigh[size_t(3) * k + f];
    const double dm = e.nx * (K[0] * e.nx + K[1] * e.ny) + e.ny * (K[2] * e.nx + K[3] * e.ny);
    double s = 0.0;
    if (nbc < 0) {
      for (int q = 0; q < R.face.n; ++q) {
        const double x = e.ax + R.face.x[q] * (e.bx - e.ax), y = e.ay + R.face.x[q] * (e.by - e.ay);
        double xi, eta, ph[NL], hx[NL], hy[NL];
        g.to_local(x, y, xi, eta);
        g.basis(xi, eta, ph, hx, hy);
        const double uv = u[0] * ph[0] + u[1] * ph[1] + u[2] * ph[2];
        const double am = combo_eval(a.a_mu, table, c, x, y);
        const double pen = s_bnd * dm * am / e.h;
        const double flux = am * ((K[0] * ux + K[1] * uy) * e.nx + (K[2] * ux + K[3] * uy) * e.ny);
        s += R.face.w[q] * e.h * (-flux + pen * uv);
      }
    } else {
      G gn;
      gn.load(m.cgeo, nbc);
      double Kn[4];
      load_tensor(m.tensor, nbc, Kn);
      double pn[NL], nx_[NL], ny_[NL], un[NL], vx = 0, vy = 0;
      gn.basis(1.0 / 3.0, 1.0 / 3.0, pn, nx_, ny_);
#pragma unroll
      for (int i = 0; i < NL; ++i) {
        un[i] = a.u_local[size_t(NL) * nbc + i];
        vx += un[i] * nx_[i]; vy += un[i] * ny_[i];
      }
      const double dp = e.nx * (Kn[0] * e.nx + Kn[1] * e.ny) + e.ny * (Kn[2] * e.nx + Kn[3] * e.ny);
      const double gamma = dp * dm / (dp + dm), wm = dp / (dp + dm), wp = dm / (dp + dm);
      for (int q = 0; q < R.face.n; ++q) {
        const double x = e.ax + R.face.x[q] * (e.bx - e.ax), y = e.ay + R.face.x[q] * (e.by - e.ay);
        double xi, eta, ph[NL], qh[NL], hx[NL], hy[NL];
        g.to_local(x, y, xi, eta);
        g.basis(xi, eta, ph, hx, hy);
        gn.to_local(x, y, xi, eta);
        gn.basis(xi, eta, qh, hx, hy);
        const double um = u[0] * ph[0] + u[1] * ph[1] + u[2] * ph[2];
        const double up = un[0] * qh[0] + un[1] * qh[1] + un[2] * qh[2];
        const double am = combo_eval(a.a_mu, table, c, x, y), ap = combo_eval(a.a_mu, table, nbc, x, y);
        const double pen = s_in * gamma * 0.5 * (am + ap) / e.h;
        const double fm = am * ((K[0] * ux + K[1] * uy) * e.nx + (K[2] * ux + K[3] * uy) * e.ny);
        const double fp = ap * ((Kn[0] * vx + Kn[1] * vy) * e.nx + (Kn[2] * vx + Kn[3] * vy) * e.ny);
        s += R.face.w[q] * e.h * (-(wm * fm + wp * fp) + pen * (um - up));
      }
    }
    Gf[f] = s;
  }
  const double area = 0.5 * g.detj;
  // t_h(x) = sum_f G_f (x - p_f) / (2 |T|), p_f the vertex opposite face f: {0,1}->2, {0,2}->1, {1,2}->0
  {
    double s = 0.0, ss = 0.0;
    const double k00 = K[3] / det, k01 = -K[1] / det, k10 = -K[2] / det, k11 = K[0] / det;
    const double kux = K[0] * ux + K[1] * uy, kuy = K[2] * ux + K[3] * uy;
    for (int q = 0; q < R.df.n; ++q) {
      double x, y;
      g.to_global(R.df.x[q], R.df.y[q], x, y);
      const double t0 = (Gf[0] * (x - g.vx[2]) + Gf[1] * (x - g.vx[1]) + Gf[2] * (x - g.vx[0])) / (2.0 * area);
      const double t1 = (Gf[0] * (y - g.vy[2]) + Gf[1] * (y - g.vy[1]) + Gf[2] * (y - g.vy[0])) / (2.0 * area);
      const double ah = combo_eval(a.a_hat, table, c, x, y), am = combo_eval(a.a_mu, table, c, x, y);
      const double w = R.df.w[q] * g.detj;
      double v0 = ah * kux + t0, v1 = ah * kuy + t1;
      s += w * (v0 * (k00 * v0 + k01 * v1) + v1 * (k10 * v0 + k11 * v1)) / ah;
      v0 = am * kux + t0;
      v1 = am * kuy + t1;
      ss += w * (v0 * (k00 * v0 + k01 * v1) + v1 * (k10 * v0 + k11 * v1)) / ah;
    }
    a.out[3 * n + k] = s;
    a.out[4 * n + k] = ss;
    v_df = s;
  }
  {
    const double div = (Gf[0] + Gf[1] + Gf[2]) / area;
    double s = 0.0;
    for (int q = 0; q < R.res.n; ++q) {
      const double d = f_res[q] - div;
      s += R.res.w[q] * g.detj * d * d;
    }
    a.out[5 * n + k] = cutoff * s;
    a.out[7 * n + k] = s;
  }
  // eta_T^2 of eta_ESV2007 (estimators/swipdg.hh:683-684)
  const double t = sqrt(v_r) + sqrt(v_df);
  a.out[8 * n + k] = v_nc + t * t;
}

// one block per segment, fixed tree => deterministic
template <bool MIN>
__global__ void __launch_bounds__(256)
    k_segment_reduce(const double* __restrict__ in, const int64_t* __restrict__ seg, double* __restrict__ out) {
  __shared__ double sm[256];
  const int64_t b = seg[blockIdx.x], e = seg[blockIdx.x + 1];
  double v = MIN ? 1e300 : 0.0;
  for (int64_t i = b + threadIdx.x; i < e; i += blockDim.x) v = MIN ? fmin(v, in[i]) : v + in[i];
  sm[threadIdx.x] = v;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (int(threadIdx.x) < o) sm[threadIdx.x] = MIN ? fmin(sm[threadIdx.x], sm[threadIdx.x + o]) : sm[threadIdx.x] + sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = sm[0];
}

}  // namespace

void launch_oswald_vertex_means(const int64_t* vptr, const int32_t* vdof, const uint8_t* vboundary, int32_t n_verts,
                                const double* u_local, double* vertex_mean, cudaStream_t s) {
  if (n_verts == 0) return;
  k_vertex_means<<<(n_verts + 255) / 256, 256, 0, s>>>(vptr, vdof, vboundary, n_verts, u_local, vertex_mean);
  count_launch();
  HDD_CUDA(cudaGetLastError());
}

void launch_indicators(const MeshView& m, const IndicatorArgs& a, int polorder, cudaStream_t s) {
  if (m.n_own == 0) return;
  if (m.kind != HDD_SIMPLEX2D)
    HDD_THROW(HDD_ERR_USING_THIS_WRONG, "estimators are only available on 2d simplex grids (estimators/swipdg.hh:71)");
  const int p = polorder, over = 2;  // over_integrate, estimators/swipdg.hh:47
  IndicatorRules R;
  R.nc = triangle_rule(a.a_bar.order + 2 * (p - 1) + over);
  R.p0 = triangle_rule(a.force_order + over);
  R.res = triangle_rule(2 * a.force_order + over);
  R.df = triangle_rule(a.a_hat.order + 2 * p + over);
  R.cut = triangle_rule(a.a_cut.order + over);
  R.amin = triangle_rule(a.a_min.order);
  R.amax = triangle_rule(a.a_max.order);
  R.face = line_rule(a.a_mu.order + 2 * p + over);
  IndicatorRules* dR = nullptr;
  HDD_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&dR), sizeof(R), s));
  HDD_CUDA(cudaMemcpyAsync(dR, &R, sizeof(R), cudaMemcpyHostToDevice, s));
  static_assert(sizeof(DevFn) % 4 == 0, "DevFn must be word sized");
  const size_t fn_bytes = size_t(a.n_fn) * sizeof(DevFn);
  if (fn_bytes > 48 * 1024) HDD_THROW(HDD_ERR_NOT_IMPLEMENTED, "too many data functions for the estimator kernel");
  k_indicators<<<(m.n_own + 127) / 128, 128, fn_bytes, s>>>(m, a, dR, sigma_inner(p), sigma_boundary(p));
  count_launch();
  HDD_CUDA(cudaGetLastError());
  HDD_CUDA(cudaStreamSynchronize(s));  // R lives on this stack frame until the copy has happened
  HDD_CUDA(cudaFreeAsync(dR, s));
}

void launch_segment_sums(const double* in, const int64_t* seg_ptr_dev, int n_seg, double* out, cudaStream_t s) {
  if (n_seg == 0) return;
  k_segment_reduce<false><<<n_seg, 256, 0, s>>>(in, seg_ptr_dev, out);
  count_launch();
  HDD_CUDA(cudaGetLastError());
}

void launch_segment_min(const double* in, const int64_t* seg_ptr_dev, int n_seg, double* out, cudaStream_t s) {
  if (n_seg == 0) return;
  k_segment_reduce<true><<<n_seg, 256, 0, s>>>(in, seg_ptr_dev, out);
  count_launch();
  HDD_CUDA(cudaGetLastError());
}

}  // namespace hdd
