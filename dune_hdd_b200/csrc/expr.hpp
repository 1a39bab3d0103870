// Expression functions: the host compiles the strings the reference hands to
// Stuff::Functions::Expression("x", "<expr>", order) (problems/OS2014.hh:63-74) and to
// Pymor::ParameterFunctional("mu", 1, "<expr>") (problems/OS2014.hh:72, problems/spe10.hh:164-167,
// discretizations/swipdg.hh:319-321) into a small postfix program that the CUDA kernels evaluate at their own
// quadrature points - no coefficient arrays travel through HBM for analytic data.
#pragma once
#include <cmath>
#include <string>

#ifdef __CUDACC__
#define HDD_HD __host__ __device__
#define HDD_FORCEINLINE __forceinline__
#else
#define HDD_HD
#define HDD_FORCEINLINE inline
#endif

namespace hdd {

enum Op : unsigned char {
  OP_CONST = 0,
  OP_VAR0,  // x[0] or mu[0] / mu
  OP_VAR1,  // x[1] or mu[1]
  OP_VAR2,
  OP_VAR3,
  OP_ADD,
  OP_SUB,
  OP_MUL,
  OP_DIV,
  OP_NEG,
  OP_POW,
  OP_SIN,
  OP_COS,
  OP_TAN,
  OP_EXP,
  OP_LOG,
  OP_SQRT,
  OP_ABS,
  OP_ATAN,
  OP_MIN,
  OP_MAX
};

constexpr int kMaxOps = 48;
constexpr int kMaxConsts = 16;
constexpr int kMaxStack = 16;

constexpr int kRegStack = 6;  // programs at most this deep run on a register stack (no local memory on the device)

struct Program {
  int n_ops;
  int depth;  // maximal stack depth of the program
  unsigned char op[kMaxOps];
  unsigned char cidx[kMaxOps];  // constant slot for OP_CONST
  double cst[kMaxConsts];
};

// Evaluates a compiled program.  vars: x[0], x[1] (functions) or mu[0..3] (parameter functionals).
// Generic version: value stack in an indexed array (local memory on the device).
HDD_HD HDD_FORCEINLINE double eval_program_generic(const Program& p, const double* vars) {
  double st[kMaxStack];
  int sp = 0;
  for (int k = 0; k < p.n_ops; ++k) {
    switch (p.op[k]) {
      case OP_CONST: st[sp++] = p.cst[p.cidx[k]]; break;
      case OP_VAR0: st[sp++] = vars[0]; break;
      case OP_VAR1: st[sp++] = vars[1]; break;
      case OP_VAR2: st[sp++] = vars[2]; break;
      case OP_VAR3: st[sp++] = vars[3]; break;
      case OP_ADD: --sp; st[sp - 1] = st[sp - 1] + st[sp]; break;
      case OP_SUB: --sp; st[sp - 1] = st[sp - 1] - st[sp]; break;
      case OP_MUL: --sp; st[sp - 1] = st[sp - 1] * st[sp]; break;
      case OP_DIV: --sp; st[sp - 1] = st[sp - 1] / st[sp]; break;
      case OP_POW: --sp; st[sp - 1] = pow(st[sp - 1], st[sp]); break;
      case OP_MIN: --sp; st[sp - 1] = fmin(st[sp - 1], st[sp]); break;
      case OP_MAX: --sp; st[sp - 1] = fmax(st[sp - 1], st[sp]); break;
      case OP_NEG: st[sp - 1] = -st[sp - 1]; break;
      case OP_SIN: st[sp - 1] = sin(st[sp - 1]); break;
      case OP_COS: st[sp - 1] = cos(st[sp - 1]); break;
      case OP_TAN: st[sp - 1] = tan(st[sp - 1]); break;
      case OP_EXP: st[sp - 1] = exp(st[sp - 1]); break;
      case OP_LOG: st[sp - 1] = log(st[sp - 1]); break;
      case OP_SQRT: st[sp - 1] = sqrt(st[sp - 1]); break;
      case OP_ABS: st[sp - 1] = fabs(st[sp - 1]); break;
      case OP_ATAN: st[sp - 1] = atan(st[sp - 1]); break;
      default: break;
    }
  }
  return sp > 0 ? st[0] : 0.0;
}

// The same program on a stack of kRegStack named values (s0 = top): every access is statically indexed, so on the
// device the stack lives in registers - the array version spills every push and pop to local memory, which showed up as
// 9x the algorithmic DRAM writes in the estimator kernel.  Identical arithmetic, identical results.
HDD_HD HDD_FORCEINLINE double eval_program(const Program& p, const double* vars) {
  if (p.depth > kRegStack) return eval_program_generic(p, vars);
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, s4 = 0.0, s5 = 0.0;
#define HDD_PUSH(v) { const double t_ = (v); s5 = s4; s4 = s3; s3 = s2; s2 = s1; s1 = s0; s0 = t_; }
#define HDD_BIN(e) { const double a_ = s1, b_ = s0; s0 = (e); s1 = s2; s2 = s3; s3 = s4; s4 = s5; }
  for (int k = 0; k < p.n_ops; ++k) {
    switch (p.op[k]) {
      case OP_CONST: HDD_PUSH(p.cst[p.cidx[k]]); break;
      case OP_VAR0: HDD_PUSH(vars[0]); break;
      case OP_VAR1: HDD_PUSH(vars[1]); break;
      case OP_VAR2: HDD_PUSH(vars[2]); break;
      case OP_VAR3: HDD_PUSH(vars[3]); break;
      case OP_ADD: HDD_BIN(a_ + b_); break;
      case OP_SUB: HDD_BIN(a_ - b_); break;
      case OP_MUL: HDD_BIN(a_ * b_); break;
      case OP_DIV: HDD_BIN(a_ / b_); break;
      case OP_POW: HDD_BIN(pow(a_, b_)); break;
      case OP_MIN: HDD_BIN(fmin(a_, b_)); break;
      case OP_MAX: HDD_BIN(fmax(a_, b_)); break;
      case OP_NEG: s0 = -s0; break;
      case OP_SIN: s0 = sin(s0); break;
      case OP_COS: s0 = cos(s0); break;
      case OP_TAN: s0 = tan(s0); break;
      case OP_EXP: s0 = exp(s0); break;
      case OP_LOG: s0 = log(s0); break;
      case OP_SQRT: s0 = sqrt(s0); break;
      case OP_ABS: s0 = fabs(s0); break;
      case OP_ATAN: s0 = atan(s0); break;
      default: break;
    }
  }
#undef HDD_PUSH
#undef HDD_BIN
  return p.n_ops > 0 ? s0 : 0.0;
}

// ---- fast form: sum of products of elementary functions of affine arguments ---------------------------------------------
//   f(x) = sum_t c_t prod_k g_tk(a_tk x[0] + b_tk x[1] + d_tk),   g in {identity, cos, sin, exp}
// Every data function of the reference's test cases has this shape - the ESV2007 force and exact solution
// (problems/ESV2007.hh:78, testcases/ESV2007.hh:41), the OS2014 diffusion factor (problems/OS2014.hh:63-74) - and evaluating
// it costs two fused multiply-adds and one elementary function per factor, against ~25 instructions of dispatch per
// postfix operation in the interpreter.  Expressions that do not fit (abs, min, max, division by a variable, ...) keep
// n_terms = 0 and run through the interpreter.
constexpr int kFastTerms = 4;
constexpr int kFastFactors = 4;
enum FastKind : int { FAST_ID = 0, FAST_COS, FAST_SIN, FAST_EXP };

struct FastFn {
  int n_terms;  // 0: not available
  struct Term {
    double c;
    int n_fac;
    int kind[kFastFactors];
    double a[kFastFactors], b[kFastFactors], d[kFastFactors];
  } t[kFastTerms];
};

HDD_HD HDD_FORCEINLINE double eval_fast(const FastFn& f, double x, double y) {
  double s = 0.0;
  for (int t = 0; t < f.n_terms; ++t) {
    const FastFn::Term& T = f.t[t];
    double v = T.c;
    for (int k = 0; k < T.n_fac; ++k) {
      const double arg = fma(T.a[k], x, fma(T.b[k], y, T.d[k]));
      switch (T.kind[k]) {
        case FAST_COS: v *= cos(arg); break;
        case FAST_SIN: v *= sin(arg); break;
        case FAST_EXP: v *= exp(arg); break;
        default: v *= arg; break;
      }
    }
    s += v;
  }
  return s;
}

// ---- branch-free cosine + the "product of cosines" form --------------------------------------------------------------------
// cos(x) for |x| <= kFastCosMax without a branch: three-constant Cody-Waite reduction by pi/2 with fused multiply-adds
// (exact product, one rounding per step: the reduced argument is good to the last bit in that range), then ONE polynomial
// in r^2 whose coefficients are selected by the quadrant's parity - the fdlibm kernels: sin r = r + r^3 S(r^2),
// cos r = 1 - r^2/2 + r^4 C(r^2) on |r| <= pi/4, both below one ulp.  The library cos() carries a slow path for huge arguments
// behind a branch, which keeps the compiler from interleaving the evaluations at neighbouring quadrature points; this one is
// ~16 fp64 operations of straight-line code.  Differs from cos() in the last bit at most (tests/test_host_logic.py).
constexpr double kFastCosMax = 1.0e5;

HDD_HD HDD_FORCEINLINE double fast_cos(double x) {
  const double q = rint(x * 0.63661977236758138);  // x * 2 / pi
  double r = fma(-q, 1.5707963267948966e+00, x);
  r = fma(-q, 6.1232339957367574e-17, r);
  r = fma(-q, 8.4784276603688985e-32, r);
  const int j = int(q) + 1;  // cos x = sin(x + pi/2): quadrant of the sine
  const bool use_cos = j & 1;
  const double z = r * r;
  // coefficients of the selected kernel, highest order first
  double p = use_cos ? -1.13596475577881948265e-11 : 1.58969099521155010221e-10;
  p = fma(p, z, use_cos ? 2.08757232129817482790e-09 : -2.50507602534068634195e-08);
  p = fma(p, z, use_cos ? -2.75573143513906633035e-07 : 2.75573137070700676789e-06);
  p = fma(p, z, use_cos ? 2.48015872894767294178e-05 : -1.98412698298579493134e-04);
  p = fma(p, z, use_cos ? -1.38888888888741095749e-03 : 8.33333333332248946124e-03);
  p = fma(p, z, use_cos ? 4.16666666666666019037e-02 : -1.66666666666666324348e-01);
  const double base = use_cos ? fma(-0.5, z, 1.0) : r;      // 1 - r^2/2   |  r
  const double mult = use_cos ? z * z : r * z;              // r^4         |  r^3
  const double v = fma(mult, p, base);
  return (j & 2) ? -v : v;
}

// f(x) = c * prod_{k < 2} cos(a_k x[0] + b_k x[1] + d_k): the shape of the ESV2007 force and exact solution and of every
// other product of at most two sines / cosines of affine arguments (sin t = cos(t - pi/2); a missing second factor is
// cos(0)).  Kernels that evaluate one function at many points per cell (the estimator's residual sweeps) take it as plain
// scalars in registers: per cell the arguments become affine functions of the reference coordinates, per point the value
// costs two fused multiply-adds and one fast_cos per factor.
struct TrigProduct {
  int valid;
  double c, a[2], b[2], d[2];
};

inline TrigProduct as_trig_product(const FastFn& f) {
  TrigProduct t{};
  if (f.n_terms != 1 || f.t[0].n_fac < 1 || f.t[0].n_fac > 2) return t;
  const FastFn::Term& T = f.t[0];
  for (int k = 0; k < 2; ++k) {
    if (k >= T.n_fac) { t.a[k] = t.b[k] = t.d[k] = 0.0; continue; }
    if (T.kind[k] != FAST_COS && T.kind[k] != FAST_SIN) return t;
    t.a[k] = T.a[k];
    t.b[k] = T.b[k];
    t.d[k] = T.kind[k] == FAST_SIN ? T.d[k] - 1.5707963267948966 : T.d[k];
  }
  t.c = T.c;
  t.valid = 1;
  return t;
}

// Tries to bring `text` (a function of var[0], var[1]) into the fast form; false (and out.n_terms = 0) if it does not fit.
bool compile_fast(const std::string& text, const std::string& var, FastFn& out);

// Compiles `text` with the (vector) variable called `var` ("x" or "mu"): var[k], and bare `var` for var[0].
// Throws hdd::Error(HDD_ERR_WRONG_INPUT) on syntax errors or programs that are too long.
Program compile_expression(const std::string& text, const std::string& var);

// Tries to write the expression as fx(var[0]) * fy(var[1]) (top-level product whose factors depend on one coordinate
// each).  Returns false if it is not of that form.  Used for tensor-product quadrature on axis-parallel cells.
bool compile_separable(const std::string& text, const std::string& var, Program& fx, Program& fy);

}  // namespace hdd
