// C++ host driver over the facade (include/hdd_b200.hpp): the ESV2007 convergence study of the reference's
// test/linearelliptic-swipdg.cc:86-109 (SWIPDGStudy, ALUGrid<2,2,simplex,conforming>, polOrder 1) - build the level grid,
// init(), solve(), estimate - printing eta_* next to the committed expectations
// (test/linearelliptic-swipdg-expectations_esv2007_2daluconform.cxx:32-57).  Exit code 0 iff all 3-digit goldens match.
//
//   g++ -std=c++17 -Iinclude examples/swipdg_esv2007.cc -Ldune_hdd_b200 -lhdd_b200 -Wl,-rpath,$PWD/dune_hdd_b200 -o swipdg_esv2007
#include <cmath>
#include <cstdio>
#include <map>
#include <string>

#include "hdd_b200.hpp"

int main(int argc, char** argv) {
  using namespace hdd;
  using LinearElliptic::Discretizations::SWIPDG;
  const int levels = argc > 1 ? std::atoi(argv[1]) : 3;
  const std::map<std::string, std::vector<double>> golden = {
      {"eta_NC_ESV2007", {1.66e-1, 7.89e-2, 3.91e-2, 1.95e-2}}, {"eta_R_ESV2007", {7.23e-2, 1.82e-2, 4.54e-3, 1.14e-3}},
      {"eta_DF_ESV2007", {3.55e-1, 1.76e-1, 8.73e-2, 4.35e-2}}, {"eta_ESV2007", {4.49e-01, 2.07e-01, 9.91e-02, 4.85e-02}},
      {"eta_ESV2007_alt", {5.93e-01, 2.73e-01, 1.31e-01, 6.42e-02}}};
  int failures = 0;
  try {
    const Problem problem = Problem::ESV2007(3);
    for (int level = 0; level < levels; ++level) {
      const Grid grid = Grid::simplex(4 << level, -1.0, 1.0, -1.0, 1.0);
      SWIPDG discretization(grid, problem);
      discretization.init();
      Vector solution = discretization.create_vector();
      LinearElliptic::Discretizations::SolveInfo info;
      discretization.uncached_solve({{"type", "cg.diagonal"}, {"precision", "1e-12"}, {"max_iter", "100000"}}, solution,
                                    Parameter(), &info);
      std::printf("level %d: %lld cells, %lld DoFs, CG %d iterations (%.3g s)\n", level, (long long)grid.n_cells(),
                  (long long)discretization.num_dofs(), info.iterations, info.seconds);
      for (const auto& type : discretization.available_estimators()) {
        auto g = golden.find(type);
        if (g == golden.end()) continue;
        const double eta = LinearElliptic::Estimators::SWIPDG::estimate(discretization, solution, type);
        const double expected = g->second[size_t(level)];
        const bool ok = std::fabs(eta - expected) <= 0.006 * expected;
        failures += ok ? 0 : 1;
        std::printf("  %-16s %.3e   expected %.2e   %s\n", type.c_str(), eta, expected, ok ? "ok" : "MISMATCH");
      }
    }
    // ESV2007 on SGrid (test/linearelliptic-swipdg-expectations_esv2007_2dsgrid.cxx:31-36): Q1, solved with the
    // multigrid-preconditioned CG, error norms evaluated on the device, products requested through only_these_products
    const double l2_golden[4] = {1.13e-02, 2.90e-03, 7.41e-04, 1.88e-04}, h1_golden[4] = {2.77e-01, 1.39e-01, 6.98e-02, 3.50e-02};
    for (int level = 0; level < levels; ++level) {
      const int n = 8 << level;
      const Grid grid = Grid::cube(n, n, -1.0, 1.0, -1.0, 1.0);
      SWIPDG discretization(grid, problem, 1, 0, nullptr, 0, -1, {"l2", "h1_semi", "elliptic"});
      discretization.init();
      Vector solution = discretization.create_vector();
      LinearElliptic::Discretizations::SolveInfo info;
      discretization.uncached_solve({{"type", "cg.mg"}, {"precision", "1e-12"}, {"max_iter", "1000"}}, solution, Parameter(), &info);
      const auto err = discretization.error_norms(solution, "cos(0.5*pi*x[0])*cos(0.5*pi*x[1])",
                                                  "-0.5*pi*sin(0.5*pi*x[0])*cos(0.5*pi*x[1])",
                                                  "-0.5*pi*cos(0.5*pi*x[0])*sin(0.5*pi*x[1])", 7);
      const bool ok = std::fabs(err.at("L2") - l2_golden[level]) <= 0.006 * l2_golden[level] &&
                      std::fabs(err.at("H1_semi") - h1_golden[level]) <= 0.006 * h1_golden[level];
      failures += ok ? 0 : 1;
      const Vector one(solution.size(), 1.0);
      const double area = discretization.get_product("l2").apply2(one, one);
      failures += std::fabs(area - 4.0) <= 1e-10 ? 0 : 1;
      std::printf("sgrid level %d: %dx%d cells, cg.mg %d iterations, L2 %.3e (expected %.2e), H1_semi %.3e (expected %.2e), "
                  "|Omega| via l2 product %.12f  %s\n", level, n, n, info.iterations, err.at("L2"), l2_golden[level],
                  err.at("H1_semi"), h1_golden[level], area, ok ? "ok" : "MISMATCH");
    }
    // error behaviour of the reference: init() not called (discretizations/base.hh:370-377)
    try {
      const Grid grid = Grid::simplex(4, -1.0, 1.0, -1.0, 1.0);
      SWIPDG d(grid, problem);
      Vector v = d.create_vector();
      d.solve(v);
      ++failures;
    } catch (const Exceptions::you_are_using_this_wrong&) {
      std::printf("solve() before init() throws you_are_using_this_wrong: ok\n");
    }
  } catch (const Exceptions::Exception& e) {
    std::fprintf(stderr, "hdd exception: %s\n", e.what());
    return 2;
  }
  return failures == 0 ? 0 : 1;
}
