// hdd_b200.hpp - header-only C++ facade over the C-ABI (hdd_b200.h) with the reference's names.
//
// Host code written against Dune::HDD::LinearElliptic::Discretizations::SWIPDG / BlockSWIPDG
// (dune/hdd/linearelliptic/discretizations/{interfaces,base,swipdg,block-swipdg}.hh) and
// Dune::HDD::LinearElliptic::Estimators::{SWIPDG,BlockSWIPDG} (estimators/{swipdg,block-swipdg}.hh) finds the same
// method names, argument meaning, solution cache and exception types here.  A DUNE host fills `Grid` from its grid
// view (cell index / reference-element vertex order / intersection neighbours) and `Problem` from its ProblemInterface
// object; INTEGRATION.md shows that glue.
#ifndef HDD_B200_HPP
#define HDD_B200_HPP

#include <cmath>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "hdd_b200.h"

namespace hdd {

// ---- exception types of the reference (Stuff::Exceptions::*, Pymor::Exceptions::*, Dune::NotImplemented) ----------
namespace Exceptions {
struct Exception : std::runtime_error { using std::runtime_error::runtime_error; };
struct wrong_input_given : Exception { using Exception::Exception; };
struct you_are_using_this_wrong : Exception { using Exception::Exception; };
struct wrong_parameter_type : Exception { using Exception::Exception; };
struct index_out_of_range : Exception { using Exception::Exception; };
struct NotImplemented : Exception { using Exception::Exception; };
struct requirements_not_met : Exception { using Exception::Exception; };
struct internal_error : Exception { using Exception::Exception; };
struct device_error : Exception { using Exception::Exception; };
struct linear_solver_failed : Exception { using Exception::Exception; };
}  // namespace Exceptions

inline void check(int status) {
  if (status == HDD_OK) return;
  const std::string msg = hdd_last_error();
  switch (status) {
    case HDD_ERR_WRONG_INPUT: throw Exceptions::wrong_input_given(msg);
    case HDD_ERR_USING_THIS_WRONG: throw Exceptions::you_are_using_this_wrong(msg);
    case HDD_ERR_WRONG_PARAMETER_TYPE: throw Exceptions::wrong_parameter_type(msg);
    case HDD_ERR_INDEX_OUT_OF_RANGE: throw Exceptions::index_out_of_range(msg);
    case HDD_ERR_NOT_IMPLEMENTED: throw Exceptions::NotImplemented(msg);
    case HDD_ERR_REQUIREMENTS_NOT_MET: throw Exceptions::requirements_not_met(msg);
    case HDD_ERR_DEVICE: throw Exceptions::device_error(msg);
    case HDD_ERR_NOT_CONVERGED: throw Exceptions::linear_solver_failed(msg);
    default: throw Exceptions::internal_error(msg);
  }
}

typedef std::vector<double> Vector;     // VectorType of the la backend
typedef std::vector<double> Parameter;  // Pymor::Parameter of the single key the problems use ("mu")

// ---- grid: what a grid view + boundary info + ms_grid boil down to -----------------------------------------------
struct Grid {
  int kind = HDD_SIMPLEX2D;
  std::vector<double> xy;
  std::vector<int32_t> cell_verts, cell_neigh, cell_subdomain;
  std::vector<uint8_t> boundary_type;  // empty = AllDirichlet
  int n_loc() const { return kind == HDD_SIMPLEX2D ? 3 : 4; }
  int64_t n_cells() const { return int64_t(cell_verts.size()) / n_loc(); }
  int64_t n_verts() const { return int64_t(xy.size()) / 2; }

  // Stuff::Grid::Providers::Cube< SGrid<2,2> >(lower_left, upper_right, num_elements [, num_partitions])
  static Grid cube(int64_t nx, int64_t ny, double x0, double x1, double y0, double y1, int px = 1, int py = 1) {
    Grid g;
    g.kind = HDD_CUBE2D;
    int64_t nc = 0, nv = 0;
    check(hdd_grid_cube_sizes(nx, ny, &nc, &nv));
    g.xy.resize(size_t(2 * nv));
    g.cell_verts.resize(size_t(4 * nc));
    g.cell_neigh.resize(size_t(4 * nc));
    g.cell_subdomain.resize(size_t(nc));
    check(hdd_grid_cube(nx, ny, x0, x1, y0, y1, px, py, g.xy.data(), g.cell_verts.data(), g.cell_neigh.data(),
                        g.cell_subdomain.data()));
    return g;
  }
  // the ALUGrid<2,2,simplex,conforming> ladder of the test cases (testcases/ESV2007.hh:123-134)
  static Grid simplex(int64_t squares_per_side, double x0, double x1, double y0, double y1, int px = 1, int py = 1) {
    Grid g;
    g.kind = HDD_SIMPLEX2D;
    int64_t nc = 0, nv = 0;
    check(hdd_grid_simplex_sizes(squares_per_side, &nc, &nv));
    g.xy.resize(size_t(2 * nv));
    g.cell_verts.resize(size_t(3 * nc));
    g.cell_neigh.resize(size_t(3 * nc));
    g.cell_subdomain.resize(size_t(nc));
    check(hdd_grid_simplex(squares_per_side, x0, x1, y0, y1, px, py, g.xy.data(), g.cell_verts.data(),
                           g.cell_neigh.data(), g.cell_subdomain.data()));
    return g;
  }
  // father cell of every cell of `fine` in this (coarser) grid: ALUGrid father() / Stuff::Grid::EntityInlevelSearch of
  // the studies (test/linearelliptic-swipdg.hh:186-194, test/linearelliptic-block-swipdg.hh:169-177)
  std::vector<int32_t> fathers_of(const Grid& fine) const {
    std::vector<int32_t> father(size_t(fine.n_cells()));
    check(hdd_grid_fathers(kind, n_cells(), n_verts(), xy.data(), cell_verts.data(), fine.n_cells(), fine.n_verts(),
                           fine.xy.data(), fine.cell_verts.data(), father.data()));
    return father;
  }
};

// ---- problem: ProblemInterface (problems/interfaces.hh:84-144) ----------------------------------------------------
struct Function {
  int kind = HDD_FN_CONSTANT;
  int order = 0;
  double value = 0.0;
  std::vector<double> cell_values;
  std::string expression;
  static Function Constant(double v) { Function f; f.value = v; return f; }
  static Function Cellwise(std::vector<double> v) { Function f; f.kind = HDD_FN_CELLWISE; f.cell_values = std::move(v); return f; }
  static Function Expression(const std::string& e, int order) {
    Function f; f.kind = HDD_FN_EXPRESSION; f.expression = e; f.order = order; return f;
  }
  hdd_function c() const {
    hdd_function f{};
    f.kind = kind; f.order = order; f.value = value;
    f.cell_values = cell_values.empty() ? nullptr : cell_values.data();
    f.expression = expression.empty() ? nullptr : expression.c_str();
    return f;
  }
};

struct AffinelyDecomposableFunction {  // Pymor::Functions::AffinelyDecomposableDefault
  std::vector<Function> components;
  std::vector<std::string> coefficients;
  std::unique_ptr<Function> affine;
  AffinelyDecomposableFunction() = default;
  explicit AffinelyDecomposableFunction(const Function& affine_part) : affine(new Function(affine_part)) {}
  void register_component(const Function& f, const std::string& coefficient) {
    components.push_back(f);
    coefficients.push_back(coefficient);
  }
  bool parametric() const { return !components.empty(); }
  bool has_affine_part() const { return bool(affine); }
};

struct Problem {
  AffinelyDecomposableFunction diffusion_factor, force, dirichlet{Function::Constant(0.0)}, neumann{Function::Constant(0.0)};
  std::vector<double> diffusion_tensor;  // 4 per cell or empty (identity)
  std::string parameter_name;
  int parameter_size = 0;
  bool parametric() const { return diffusion_factor.parametric() || force.parametric() || dirichlet.parametric(); }

  static Problem ESV2007(int integration_order = 3) {  // problems/ESV2007.hh:75-81
    Problem p;
    p.diffusion_factor = AffinelyDecomposableFunction(Function::Constant(1.0));
    p.force = AffinelyDecomposableFunction(Function::Expression("0.5*pi*pi*cos(0.5*pi*x[0])*cos(0.5*pi*x[1])", integration_order));
    return p;
  }
  static Problem OS2014ParametricESV2007(int integration_order = 3) {  // problems/OS2014.hh:63-113
    Problem p = ESV2007(integration_order);
    p.diffusion_factor = AffinelyDecomposableFunction(Function::Expression("1+0.75*(sin(4*pi*(x[0]+0.5*x[1])))", integration_order));
    p.diffusion_factor.register_component(Function::Expression("-0.75*(sin(4*pi*(x[0]+0.5*x[1])))", integration_order), "mu");
    p.parameter_name = "mu";
    p.parameter_size = 1;
    return p;
  }
};

namespace internal {
struct CAffine {
  std::vector<hdd_function> comps;
  std::vector<const char*> coefs;
  hdd_function affine{};
  hdd_affine_function c{};
  explicit CAffine(const AffinelyDecomposableFunction& f) {
    for (const auto& k : f.components) comps.push_back(k.c());
    for (const auto& s : f.coefficients) coefs.push_back(s.c_str());
    c.n_components = int(comps.size());
    c.components = comps.empty() ? nullptr : comps.data();
    c.coefficients = coefs.empty() ? nullptr : coefs.data();
    if (f.affine) { affine = f.affine->c(); c.affine_part = &affine; }
  }
};
}  // namespace internal

namespace LinearElliptic {
namespace Discretizations {

struct SolveInfo : hdd_solve_info {};

// AffinelyDecomposedContainer view returned by system_matrix() / rhs() / get_operator() / get_rhs()
class AffinelyDecomposedContainer {
 public:
  AffinelyDecomposedContainer(hdd_swipdg* h, int which) : h_(h), which_(which) {}
  int num_components() const { int n = 0, a = 0; check(hdd_num_components(h_, which_, &n, &a)); return n; }
  bool has_affine_part() const { int n = 0, a = 0; check(hdd_num_components(h_, which_, &n, &a)); return a != 0; }
  bool parametric() const { return num_components() > 0; }
  std::string coefficient(int q) const { const char* s = nullptr; check(hdd_component_coefficient(h_, which_, q, &s)); return s; }
  // device pointer to the values of component q (matrix: CSR values over pattern(); vector: entries)
  const double* component_dev(int q, int64_t* count = nullptr) const {
    const double* p = nullptr; check(hdd_component_values(h_, which_, q, &p, count)); return p;
  }
  const double* affine_part_dev(int64_t* count = nullptr) const { return component_dev(-1, count); }
  Vector component(int q) const {
    int64_t n = 0; const double* p = component_dev(q, &n);
    Vector v(static_cast<size_t>(n)); check(hdd_copy_to_host(h_, v.data(), p, v.size() * sizeof(double))); return v;
  }
  Vector affine_part() const { return component(-1); }
 private:
  hdd_swipdg* h_;
  int which_;
};

// get_product(id) (discretizations/base.hh:281-291): the affinely decomposed product matrix
class Product {
 public:
  Product(hdd_swipdg* h, std::string id) : h_(h), id_(std::move(id)) { info(nullptr, nullptr, nullptr); }
  int num_components() const { int n = 0; info(&n, nullptr, nullptr); return n; }
  bool has_affine_part() const { int a = 0; info(nullptr, &a, nullptr); return a != 0; }
  // true: values follow SWIPDG::pattern_volume() (one dense n_loc x n_loc block per cell), false: SWIPDG::pattern()
  bool volume_pattern() const { int v = 0; info(nullptr, nullptr, &v); return v != 0; }
  std::string coefficient(int q) const { const char* s = nullptr; check(hdd_product_coefficient(h_, id_.c_str(), q, &s)); return s; }
  const double* component_dev(int q, int64_t* count = nullptr) const {
    const double* p = nullptr; check(hdd_product_values(h_, id_.c_str(), q, &p, count)); return p;
  }
  Vector component(int q) const {
    int64_t n = 0; const double* p = component_dev(q, &n);
    Vector v(static_cast<size_t>(n)); check(hdd_copy_to_host(h_, v.data(), p, v.size() * sizeof(double))); return v;
  }
  Vector affine_part() const { return component(-1); }
  // freeze_parameter(mu).apply2(u, v) = u^T P(mu) v, evaluated on the device
  double apply2(const Vector& u, const Vector& v, const Parameter& mu = Parameter()) const {
    double r = 0.0;
    int64_t g = 0, o = 0;
    check(hdd_num_dofs(h_, &g, &o));
    if (int64_t(u.size()) != o || int64_t(v.size()) != o)
      throw Exceptions::wrong_input_given("apply2: the vectors must have num_owned_dofs() entries");  // shapes_do_not_match
    check(hdd_product_apply2(h_, id_.c_str(), mu.empty() ? nullptr : mu.data(), int(mu.size()), u.data(), v.data(), &r));
    return r;
  }
  double induced_norm(const Vector& u, const Parameter& mu = Parameter()) const {
    const double s = apply2(u, u, mu);
    return s > 0.0 ? std::sqrt(s) : 0.0;
  }
 private:
  void info(int* n, int* a, int* v) const { check(hdd_product_num_components(h_, id_.c_str(), n, a, v)); }
  hdd_swipdg* h_;
  std::string id_;
};

// Discretizations::SWIPDG< GridType, layer, double, 1, polOrder > and BlockSWIPDG share this class: a grid with
// cell_subdomain is the multiscale grid of BlockSWIPDG (boundary info forced to AllDirichlet there,
// discretizations/block-swipdg.hh:110,237).
class SWIPDG {
 public:
  static std::string static_id() { return "hdd.linearelliptic.discretizations.swipdg"; }

  SWIPDG(const Grid& grid, const Problem& problem, int polorder = 1, int device = 0, hdd_comm* comm = nullptr,
         int64_t cell_begin = 0, int64_t cell_end = -1, const std::vector<std::string>& only_these_products = {})
      : n_loc_(grid.kind == HDD_SIMPLEX2D ? (polorder + 1) * (polorder + 2) / 2 : (polorder + 1) * (polorder + 1)) {
    if (cell_end < 0) cell_end = grid.n_cells();
    check(hdd_mesh_create(grid.kind, grid.n_cells(), grid.n_verts(), grid.xy.data(), grid.cell_verts.data(),
                          grid.cell_neigh.data(), grid.cell_subdomain.empty() ? nullptr : grid.cell_subdomain.data(),
                          grid.boundary_type.empty() ? nullptr : grid.boundary_type.data(), cell_begin, cell_end, device,
                          &mesh_));
    finish_construction(problem, polorder, comm, only_these_products);
  }

 private:
  void finish_construction(const Problem& problem, int polorder, hdd_comm* comm, const std::vector<std::string>& only_these_products) {
    try {
      if (comm) check(hdd_mesh_attach_comm(mesh_, comm));
      internal::CAffine fac(problem.diffusion_factor), frc(problem.force), dir(problem.dirichlet), neu(problem.neumann);
      hdd_problem p{};
      p.diffusion_factor = fac.c; p.force = frc.c; p.dirichlet = dir.c; p.neumann = neu.c;
      p.diffusion_tensor = problem.diffusion_tensor.empty() ? nullptr : problem.diffusion_tensor.data();
      p.parameter_name = problem.parameter_name.empty() ? nullptr : problem.parameter_name.c_str();
      p.parameter_size = problem.parameter_size;
      check(hdd_swipdg_create(mesh_, polorder, &p, &h_));
      if (!only_these_products.empty()) {
        std::vector<const char*> ids;
        for (const auto& id : only_these_products) ids.push_back(id.c_str());
        check(hdd_swipdg_only_these_products(h_, ids.data(), int(ids.size())));
      }
    } catch (...) {
      if (h_) hdd_swipdg_destroy(h_);
      hdd_mesh_destroy(mesh_);
      throw;
    }
  }

 public:
  // Stuff::Grid::Providers::Cube(lower_left, upper_right, num_elements) with a [px py 1] multiscale partition
  // (testcases/ESV2007.hh:123-127, :150-163): the grid tables are written on the device (hdd_mesh_create_cube)
  struct CubeProvider {
    int64_t nx = 1, ny = 1;
    double lower_left[2] = {-1.0, -1.0}, upper_right[2] = {1.0, 1.0};
    int px = 1, py = 1;
  };
  SWIPDG(const CubeProvider& grid, const Problem& problem, int polorder = 1, int device = 0, hdd_comm* comm = nullptr,
         int64_t cell_begin = 0, int64_t cell_end = -1, const std::vector<std::string>& only_these_products = {})
      : n_loc_((polorder + 1) * (polorder + 1)) {
    check(hdd_mesh_create_cube(grid.nx, grid.ny, grid.lower_left[0], grid.upper_right[0], grid.lower_left[1], grid.upper_right[1],
                               grid.px, grid.py, cell_begin, cell_end, device, &mesh_));
    finish_construction(problem, polorder, comm, only_these_products);
  }
  SWIPDG(const SWIPDG&) = delete;
  SWIPDG& operator=(const SWIPDG&) = delete;
  ~SWIPDG() {
    if (h_) hdd_swipdg_destroy(h_);
    if (mesh_) hdd_mesh_destroy(mesh_);
  }

  void init() { check(hdd_swipdg_init(h_)); }  // idempotent, discretizations/swipdg.hh:206-512

  // space / containers
  int64_t num_dofs() const { int64_t g = 0, o = 0; check(hdd_num_dofs(h_, &g, &o)); return g; }
  int64_t num_owned_dofs() const { int64_t g = 0, o = 0; check(hdd_num_dofs(h_, &g, &o)); return o; }
  Vector create_vector() const { return Vector(static_cast<size_t>(num_owned_dofs()), 0.0); }
  void pattern(std::vector<int64_t>& rowptr, std::vector<int32_t>& col) const {
    int64_t n = 0, nnz = 0; const int64_t* rp = nullptr; const int32_t* cl = nullptr;
    check(hdd_pattern(h_, &n, &nnz, &rp, &cl));
    rowptr.resize(size_t(n) + 1); col.resize(size_t(nnz));
    check(hdd_copy_to_host(h_, rowptr.data(), rp, rowptr.size() * sizeof(int64_t)));
    check(hdd_copy_to_host(h_, col.data(), cl, col.size() * sizeof(int32_t)));
  }
  AffinelyDecomposedContainer system_matrix() const { return AffinelyDecomposedContainer(h_, HDD_LHS); }
  AffinelyDecomposedContainer rhs() const { return AffinelyDecomposedContainer(h_, HDD_RHS); }
  AffinelyDecomposedContainer get_operator() const { return system_matrix(); }
  AffinelyDecomposedContainer get_rhs() const { return rhs(); }
  // the C-ABI copies num_owned_dofs() doubles from / to the pointers it gets: a vector of another size is refused here,
  // where the reference throws Stuff::Exceptions::shapes_do_not_match
  void require_owned_size(const Vector& v, const char* what) const {
    if (int64_t(v.size()) != num_owned_dofs())
      throw Exceptions::wrong_input_given(std::string(what) + ": the vector has " + std::to_string(v.size()) + " entries, the space " +
                                          std::to_string(num_owned_dofs()) + " (owned) DoFs");
  }
  Vector apply(const Vector& x, const Parameter& mu = Parameter()) const {
    require_owned_size(x, "apply");
    Vector y(x.size());
    check(hdd_apply(h_, mu.empty() ? nullptr : mu.data(), int(mu.size()), x.data(), y.data()));
    return y;
  }

  // products (discretizations/base.hh:272-291)
  std::vector<std::string> available_products() const {
    const char* const* t = nullptr; int n = 0; check(hdd_products_available(h_, &t, &n));
    return std::vector<std::string>(t, t + n);
  }
  Product get_product(const std::string& id) const { return Product(h_, id); }
  void pattern_volume(std::vector<int64_t>& rowptr, std::vector<int32_t>& col) const {
    int64_t n = 0, nnz = 0; const int64_t* rp = nullptr; const int32_t* cl = nullptr;
    check(hdd_pattern_volume(h_, &n, &nnz, &rp, &cl));
    rowptr.resize(size_t(n) + 1); col.resize(size_t(nnz));
    check(hdd_copy_to_host(h_, rowptr.data(), rp, rowptr.size() * sizeof(int64_t)));
    check(hdd_copy_to_host(h_, col.data(), cl, col.size() * sizeof(int32_t)));
  }
  // {L2, H1_semi, energy} norms of vector - exact (test/linearelliptic-swipdg.hh:267-290); exact solution and its
  // gradient as Expression strings in x[0], x[1]
  std::map<std::string, double> error_norms(const Vector& vector, const std::string& exact, const std::string& exact_dx,
                                            const std::string& exact_dy, int order = 5,
                                            const Parameter& mu = Parameter()) const {
    double out[3] = {0.0, 0.0, 0.0};
    require_owned_size(vector, "error_norms");
    check(hdd_error_norms(h_, vector.data(), exact.c_str(), exact_dx.c_str(), exact_dy.c_str(), order,
                          mu.empty() ? nullptr : mu.data(), int(mu.size()), out));
    return {{"L2", out[0]}, {"H1_semi", out[1]}, {"energy", out[2]}};
  }

  // GDT::Operators::Prolongation(this grid view).apply(coarse function, fine function) (test/linearelliptic.hh:168-176);
  // father = coarse_grid.fathers_of(this grid), restricted to the owned cells
  Vector prolong(const SWIPDG& coarse, const Vector& coarse_vector, const std::vector<int32_t>& father) const {
    Vector fine = create_vector();
    check(hdd_prolong(coarse.h_, coarse_vector.data(), h_, father.data(), fine.data()));
    return fine;
  }

  // solve (CachedDefault::solve + ContainerBasedDefault::uncached_solve, discretizations/base.hh:151-178,327-367)
  std::vector<std::string> solver_types() const {
    const char* const* t = nullptr; int n = 0; check(hdd_solver_types(&t, &n));
    return std::vector<std::string>(t, t + n);
  }
  std::map<std::string, std::string> solver_options(const std::string& type = "") const {
    return {{"type", type.empty() ? solver_types()[0] : type}, {"precision", "1e-10"}, {"max_iter", "100000"}};
  }
  void solve(const std::map<std::string, std::string>& options, Vector& vector, const Parameter& mu = Parameter()) {
    const auto key = std::make_pair(options, mu);
    auto it = cache_.find(key);
    if (it == cache_.end()) {
      uncached_solve(options, vector, mu);
      cache_[key] = vector;
    } else {
      vector = it->second;
    }
  }
  void solve(Vector& vector, const Parameter& mu = Parameter()) { solve(solver_options(), vector, mu); }
  void uncached_solve(const std::map<std::string, std::string>& options, Vector& vector, const Parameter& mu = Parameter(),
                      SolveInfo* info = nullptr) {
    auto get = [&](const char* k, const char* dflt) { auto i = options.find(k); return i == options.end() ? std::string(dflt) : i->second; };
    vector.resize(static_cast<size_t>(num_owned_dofs()));
    SolveInfo local;
    check(hdd_solve(h_, get("type", "").c_str(), std::stod(get("precision", "1e-10")), std::stoi(get("max_iter", "100000")),
                    mu.empty() ? nullptr : mu.data(), int(mu.size()), vector.data(), info ? info : &local));
  }

  // BlockSWIPDG views (discretizations/block-swipdg.hh:553-690)
  int num_subdomains() const { int n = 0; check(hdd_num_subdomains(h_, &n)); return n; }
  std::vector<int> neighbouring_subdomains(int ss) const {
    const int32_t* p = nullptr; int n = 0; check(hdd_neighbouring_subdomains(h_, ss, &p, &n));
    return std::vector<int>(p, p + n);
  }
  std::vector<int64_t> subdomain_offsets() const {
    const int64_t* p = nullptr; check(hdd_subdomain_offsets(h_, &p));
    return std::vector<int64_t>(p, p + num_subdomains() + 1);
  }
  struct Csr { int64_t n_rows = 0, n_cols = 0; std::vector<int64_t> rowptr; std::vector<int32_t> col; std::vector<double> val; };
  Csr get_local_operator(int ss, int q = -1) const { return block(ss, ss, q); }
  Csr get_coupling_operator(int ss, int nn, int q = -1) const { return block(ss, nn, q); }
  Vector localize_vector(const Vector& global, int ss) const {
    const auto off = subdomain_offsets();
    if (ss < 0 || ss >= num_subdomains()) throw Exceptions::index_out_of_range("0 <= ss < num_subdomains() is not true");
    if (int64_t(global.size()) != off.back()) throw Exceptions::index_out_of_range("The size() of global_vector does not match");
    return Vector(global.begin() + off[size_t(ss)], global.begin() + off[size_t(ss) + 1]);
  }

  // estimators (Estimators::SWIPDG / BlockSWIPDG string dispatch)
  std::vector<std::string> available_estimators() const {
    const char* const* t = nullptr; int n = 0; check(hdd_estimators_available(h_, &t, &n));
    return std::vector<std::string>(t, t + n);
  }
  double estimate(const Vector& vector, const std::string& type, const std::map<std::string, Parameter>& parameters = {},
                  Vector* local = nullptr) const {
    hdd_parameters p{};
    auto set = [&](const char* key, const double*& dst) {
      auto it = parameters.find(key);
      if (it != parameters.end() && !it->second.empty()) { dst = it->second.data(); p.mu_size = int(it->second.size()); }
    };
    set("mu", p.mu); set("mu_hat", p.mu_hat); set("mu_bar", p.mu_bar);
    set("parameter_range_min", p.parameter_range_min); set("parameter_range_max", p.parameter_range_max);
    double eta = 0.0;
    require_owned_size(vector, "estimate");
    if (local) local->assign(size_t(type.find("OS2014") != std::string::npos ? num_subdomains() : num_owned_dofs() / n_loc_), 0.0);
    check(hdd_estimate(h_, type.c_str(), vector.data(), &p, &eta, local ? local->data() : nullptr));
    return eta;
  }

  hdd_swipdg* handle() const { return h_; }

 private:
  Csr block(int ss, int nn, int q) const {
    hdd_csr m{};
    check(hdd_block_extract(h_, ss, nn, q, &m));
    Csr out;
    out.n_rows = m.n_rows; out.n_cols = m.n_cols;
    out.rowptr.assign(m.rowptr, m.rowptr + m.n_rows + 1);
    out.col.assign(m.col, m.col + m.nnz);
    out.val.assign(m.val, m.val + m.nnz);
    hdd_csr_free(&m);
    return out;
  }
  hdd_mesh* mesh_ = nullptr;
  hdd_swipdg* h_ = nullptr;
  int n_loc_;
  std::map<std::pair<std::map<std::string, std::string>, Parameter>, Vector> cache_;  // (options, mu) -> solution
};

typedef SWIPDG BlockSWIPDG;

}  // namespace Discretizations

namespace Estimators {
// Estimators::SWIPDG<...>::estimate(space, vector, problem, type) / BlockSWIPDG<...>::estimate(..., type, parameters)
struct SWIPDG {
  static double estimate(const Discretizations::SWIPDG& d, const Vector& v, const std::string& type) { return d.estimate(v, type); }
};
struct BlockSWIPDG {
  static double estimate(const Discretizations::SWIPDG& d, const Vector& v, const std::string& type,
                         const std::map<std::string, Parameter>& parameters = {}) {
    return d.estimate(v, type, parameters);
  }
};
}  // namespace Estimators
}  // namespace LinearElliptic
}  // namespace hdd

#endif  // HDD_B200_HPP
