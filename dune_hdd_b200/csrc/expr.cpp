// Compiler for the expression strings of Stuff::Functions::Expression and Pymor::ParameterFunctional (see expr.hpp):
// recursive descent -> small AST -> constant folding -> postfix program.  For functions of x = (x[0], x[1]) it also
// tries to split the expression into g(x[0]) * h(x[1]) (a top-level product whose factors each depend on one
// coordinate only), which lets tensor-product quadrature on axis-parallel cells evaluate n + n instead of n * n
// transcendental factors per cell.
#include "expr.hpp"

#include <cctype>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.hpp"

namespace hdd {
namespace {

struct Node {
  Op op;
  double value = 0.0;
  int a = -1, b = -1;
};

bool is_binary(Op op) {
  return op == OP_ADD || op == OP_SUB || op == OP_MUL || op == OP_DIV || op == OP_POW || op == OP_MIN || op == OP_MAX;
}
bool is_leaf(Op op) { return op == OP_CONST || (op >= OP_VAR0 && op <= OP_VAR3); }

struct Parser {
  const std::string& s;
  const std::string& var;
  size_t pos = 0;
  std::vector<Node> nodes;

  Parser(const std::string& text, const std::string& v) : s(text), var(v) {}

  [[noreturn]] void fail(const std::string& why) {
    HDD_THROW(HDD_ERR_WRONG_INPUT, "cannot parse expression '" << s << "' at position " << pos << ": " << why);
  }
  void skip() {
    while (pos < s.size() && std::isspace(static_cast<unsigned char>(s[pos]))) ++pos;
  }
  bool accept(char c) {
    skip();
    if (pos < s.size() && s[pos] == c) { ++pos; return true; }
    return false;
  }
  int make(Op op, int a = -1, int b = -1, double value = 0.0) {
    Node n;
    n.op = op; n.a = a; n.b = b; n.value = value;
    nodes.push_back(n);
    return int(nodes.size()) - 1;
  }

  int expr() {
    int l = term();
    for (;;) {
      if (accept('+')) l = make(OP_ADD, l, term());
      else if (accept('-')) l = make(OP_SUB, l, term());
      else return l;
    }
  }
  int term() {
    int l = unary();
    for (;;) {
      if (accept('*')) l = make(OP_MUL, l, unary());
      else if (accept('/')) l = make(OP_DIV, l, unary());
      else return l;
    }
  }
  int unary() {
    if (accept('-')) return make(OP_NEG, unary());
    if (accept('+')) return unary();
    return power();
  }
  int power() {
    const int base = primary();
    if (accept('^')) return make(OP_POW, base, unary());
    return base;
  }
  int primary() {
    skip();
    if (pos >= s.size()) fail("unexpected end");
    const char c = s[pos];
    if (c == '(') {
      ++pos;
      const int e = expr();
      if (!accept(')')) fail("expected ')'");
      return e;
    }
    if (std::isdigit(static_cast<unsigned char>(c)) || c == '.') {
      char* end = nullptr;
      const double v = std::strtod(s.c_str() + pos, &end);
      if (end == s.c_str() + pos) fail("bad number");
      pos = size_t(end - s.c_str());
      return make(OP_CONST, -1, -1, v);
    }
    if (std::isalpha(static_cast<unsigned char>(c)) || c == '_') {
      const size_t b = pos;
      while (pos < s.size() && (std::isalnum(static_cast<unsigned char>(s[pos])) || s[pos] == '_')) ++pos;
      const std::string id = s.substr(b, pos - b);
      if (id == var) {
        int k = 0;
        if (accept('[')) {
          skip();
          char* end = nullptr;
          k = int(std::strtol(s.c_str() + pos, &end, 10));
          if (end == s.c_str() + pos) fail("bad index");
          pos = size_t(end - s.c_str());
          if (!accept(']')) fail("expected ']'");
        }
        if (k < 0 || k > 3) fail("variable index out of range");
        return make(Op(OP_VAR0 + k));
      }
      if (id == "pi" || id == "PI" || id == "M_PI") return make(OP_CONST, -1, -1, 3.14159265358979323846264338327950288);
      struct Fn { const char* name; Op op; int nargs; };
      static const Fn fns[] = {{"sin", OP_SIN, 1},   {"cos", OP_COS, 1},   {"tan", OP_TAN, 1}, {"exp", OP_EXP, 1},
                               {"log", OP_LOG, 1},   {"sqrt", OP_SQRT, 1}, {"abs", OP_ABS, 1}, {"atan", OP_ATAN, 1},
                               {"pow", OP_POW, 2},   {"min", OP_MIN, 2},   {"max", OP_MAX, 2}};
      for (const Fn& f : fns)
        if (id == f.name) {
          if (!accept('(')) fail("expected '(' after function name");
          const int a = expr();
          int b2 = -1;
          if (f.nargs == 2) {
            if (!accept(',')) fail("expected ','");
            b2 = expr();
          }
          if (!accept(')')) fail("expected ')'");
          return make(f.op, a, b2);
        }
      fail("unknown identifier '" + id + "'");
    }
    fail(std::string("unexpected character '") + c + "'");
  }

  // ---- analysis ------------------------------------------------------------------------------------------
  int deps(int n) const {  // bit k set <=> depends on var[k]
    const Node& x = nodes[size_t(n)];
    if (x.op == OP_CONST) return 0;
    if (x.op >= OP_VAR0 && x.op <= OP_VAR3) return 1 << (x.op - OP_VAR0);
    int d = deps(x.a);
    if (x.b >= 0) d |= deps(x.b);
    return d;
  }
  void emit(int n, Program& p, int& n_consts, int& depth, int& max_depth) {
    const Node& x = nodes[size_t(n)];
    auto push = [&](Op op, int delta) {
      if (p.n_ops >= kMaxOps) fail("expression too long");
      p.op[p.n_ops] = static_cast<unsigned char>(op);
      p.cidx[p.n_ops] = 0;
      ++p.n_ops;
      depth += delta;
      if (depth > max_depth) max_depth = depth;
      if (max_depth > kMaxStack) fail("expression too deeply nested");
    };
    if (x.op == OP_CONST) {
      int slot = -1;
      for (int k = 0; k < n_consts; ++k)
        if (std::memcmp(&p.cst[k], &x.value, sizeof(double)) == 0) slot = k;
      if (slot < 0) {
        if (n_consts >= kMaxConsts) fail("too many constants");
        slot = n_consts++;
        p.cst[slot] = x.value;
      }
      push(OP_CONST, +1);
      p.cidx[p.n_ops - 1] = static_cast<unsigned char>(slot);
      return;
    }
    if (is_leaf(x.op)) { push(x.op, +1); return; }
    emit(x.a, p, n_consts, depth, max_depth);
    if (x.b >= 0) emit(x.b, p, n_consts, depth, max_depth);
    push(x.op, x.b >= 0 ? -1 : 0);
  }
  Program program_of(int root) {
    Program p{};
    p.n_ops = 0;
    int n_consts = 0, depth = 0, max_depth = 0;
    emit(root, p, n_consts, depth, max_depth);
    p.depth = max_depth;
    return p;
  }
  // constant folding, bottom-up; returns the (possibly new) node
  int fold(int n) {
    Node x = nodes[size_t(n)];
    if (is_leaf(x.op)) return n;
    const int a = fold(x.a);
    const int b = x.b >= 0 ? fold(x.b) : -1;
    const int m = make(x.op, a, b);
    if (nodes[size_t(a)].op == OP_CONST && (b < 0 || nodes[size_t(b)].op == OP_CONST)) {
      const Program p = program_of(m);
      const double zero[4] = {0, 0, 0, 0};
      return make(OP_CONST, -1, -1, eval_program(p, zero));
    }
    return m;
  }
  // ---- fast form (see expr.hpp): normal form = list of terms, term = coefficient * list of (kind, a, b, d) factors
  struct NfFactor { int kind; double a, b, d; };
  struct NfTerm { double c; std::vector<NfFactor> f; };
  typedef std::vector<NfTerm> Nf;
  static bool nf_affine(const Nf& n, double& a, double& b, double& d) {
    a = b = d = 0.0;
    for (const NfTerm& t : n) {
      if (t.f.empty()) { d += t.c; continue; }
      if (t.f.size() != 1 || t.f[0].kind != FAST_ID) return false;
      a += t.c * t.f[0].a;
      b += t.c * t.f[0].b;
      d += t.c * t.f[0].d;
    }
    return true;
  }
  bool to_nf(int n, Nf& out) const {
    const Node& x = nodes[size_t(n)];
    out.clear();
    switch (x.op) {
      case OP_CONST: out.push_back({x.value, {}}); return true;
      case OP_VAR0: out.push_back({1.0, {{FAST_ID, 1.0, 0.0, 0.0}}}); return true;
      case OP_VAR1: out.push_back({1.0, {{FAST_ID, 0.0, 1.0, 0.0}}}); return true;
      case OP_ADD: case OP_SUB: {
        Nf a, b;
        if (!to_nf(x.a, a) || !to_nf(x.b, b)) return false;
        out = a;
        for (NfTerm t : b) { if (x.op == OP_SUB) t.c = -t.c; out.push_back(t); }
        return out.size() <= size_t(kFastTerms);
      }
      case OP_NEG: {
        if (!to_nf(x.a, out)) return false;
        for (NfTerm& t : out) t.c = -t.c;
        return true;
      }
      case OP_MUL: {
        Nf a, b;
        if (!to_nf(x.a, a) || !to_nf(x.b, b)) return false;
        for (const NfTerm& ta : a)
          for (const NfTerm& tb : b) {
            NfTerm t{ta.c * tb.c, ta.f};
            t.f.insert(t.f.end(), tb.f.begin(), tb.f.end());
            if (t.f.size() > size_t(kFastFactors)) return false;
            out.push_back(t);
          }
        return out.size() <= size_t(kFastTerms);
      }
      case OP_DIV: {
        Nf a, b;
        if (!to_nf(x.a, a) || !to_nf(x.b, b)) return false;
        if (b.size() != 1 || !b[0].f.empty() || b[0].c == 0.0) return false;  // division by a constant only
        out = a;
        for (NfTerm& t : out) t.c /= b[0].c;
        return true;
      }
      case OP_POW: {
        Nf a, b;
        if (!to_nf(x.a, a) || !to_nf(x.b, b)) return false;
        if (b.size() != 1 || !b[0].f.empty()) return false;
        const double e = b[0].c;
        if (e != 2.0 && e != 3.0 && e != 1.0) return false;
        out = a;
        for (int k = 1; k < int(e); ++k) {
          Nf prod;
          for (const NfTerm& ta : out)
            for (const NfTerm& tb : a) {
              NfTerm t{ta.c * tb.c, ta.f};
              t.f.insert(t.f.end(), tb.f.begin(), tb.f.end());
              if (t.f.size() > size_t(kFastFactors)) return false;
              prod.push_back(t);
            }
          if (prod.size() > size_t(kFastTerms)) return false;
          out.swap(prod);
        }
        return true;
      }
      case OP_COS: case OP_SIN: case OP_EXP: {
        Nf a;
        if (!to_nf(x.a, a)) return false;
        double A, B, D;
        if (!nf_affine(a, A, B, D)) return false;
        out.push_back({1.0, {{x.op == OP_COS ? FAST_COS : x.op == OP_SIN ? FAST_SIN : FAST_EXP, A, B, D}}});
        return true;
      }
      default: return false;
    }
  }

  // flattens a top-level product into factors (handles unary minus and division)
  void factors(int n, std::vector<int>& out) {
    const Node x = nodes[size_t(n)];
    if (x.op == OP_MUL) { factors(x.a, out); factors(x.b, out); return; }
    if (x.op == OP_NEG) { out.push_back(make(OP_CONST, -1, -1, -1.0)); factors(x.a, out); return; }
    if (x.op == OP_DIV) {
      factors(x.a, out);
      out.push_back(make(OP_DIV, make(OP_CONST, -1, -1, 1.0), x.b));
      return;
    }
    out.push_back(n);
  }
  int product(const std::vector<int>& f) {
    if (f.empty()) return make(OP_CONST, -1, -1, 1.0);
    int p = f[0];
    for (size_t k = 1; k < f.size(); ++k) p = make(OP_MUL, p, f[k]);
    return p;
  }
};

}  // namespace

Program compile_expression(const std::string& text, const std::string& var) {
  Parser p(text, var);
  const int root = p.expr();
  p.skip();
  if (p.pos != text.size()) p.fail("trailing characters");
  return p.program_of(p.fold(root));
}

bool compile_fast(const std::string& text, const std::string& var, FastFn& out) {
  std::memset(&out, 0, sizeof(out));
  Parser p(text, var);
  const int root = p.expr();
  p.skip();
  if (p.pos != text.size()) p.fail("trailing characters");
  Parser::Nf nf;
  if (!p.to_nf(p.fold(root), nf) || nf.empty() || nf.size() > size_t(kFastTerms)) return false;
  // constant terms first merged into one, identity factors with zero slope folded into the coefficient
  for (Parser::NfTerm& t : nf)
    for (size_t k = 0; k < t.f.size();)
      if (t.f[k].kind == FAST_ID && t.f[k].a == 0.0 && t.f[k].b == 0.0) {
        t.c *= t.f[k].d;
        t.f.erase(t.f.begin() + long(k));
      } else {
        ++k;
      }
  out.n_terms = int(nf.size());
  for (int t = 0; t < out.n_terms; ++t) {
    out.t[t].c = nf[size_t(t)].c;
    out.t[t].n_fac = int(nf[size_t(t)].f.size());
    for (int k = 0; k < out.t[t].n_fac; ++k) {
      out.t[t].kind[k] = nf[size_t(t)].f[size_t(k)].kind;
      out.t[t].a[k] = nf[size_t(t)].f[size_t(k)].a;
      out.t[t].b[k] = nf[size_t(t)].f[size_t(k)].b;
      out.t[t].d[k] = nf[size_t(t)].f[size_t(k)].d;
    }
  }
  return true;
}

bool compile_separable(const std::string& text, const std::string& var, Program& fx, Program& fy) {
  Parser p(text, var);
  const int root = p.expr();
  p.skip();
  if (p.pos != text.size()) p.fail("trailing characters");
  std::vector<int> all, fxs, fys;
  p.factors(p.fold(root), all);
  for (int f : all) {
    const int d = p.deps(f);
    if (d == 0 || d == 1) fxs.push_back(f);  // constants ride with the x factor
    else if (d == 2) fys.push_back(f);
    else return false;
  }
  try {
    fx = p.program_of(p.fold(p.product(fxs)));
    fy = p.program_of(p.fold(p.product(fys)));
  } catch (const Error&) {
    return false;
  }
  return true;
}

}  // namespace hdd

// host view of fast_cos for the CPU tests (include/hdd_b200.h)
extern "C" int hdd_fast_cos(const double* x, int64_t n, double* out) {
  if (!x || !out || n < 0) return HDD_ERR_WRONG_INPUT;
  for (int64_t i = 0; i < n; ++i) out[i] = hdd::fast_cos(x[i]);
  return HDD_OK;
}

// host view of the TrigProduct recognition for the CPU tests (include/hdd_b200.h): out[7] = c, a0, b0, d0, a1, b1, d1
extern "C" int hdd_trig_product(const char* expression, double* out, int* valid) {
  if (!expression || !out || !valid) return HDD_ERR_WRONG_INPUT;
  try {
    hdd::FastFn f;
    hdd::TrigProduct t{};
    if (hdd::compile_fast(expression, "x", f)) t = hdd::as_trig_product(f);
    *valid = t.valid;
    const double v[7] = {t.c, t.a[0], t.b[0], t.d[0], t.a[1], t.b[1], t.d[1]};
    for (int k = 0; k < 7; ++k) out[k] = v[k];
    return HDD_OK;
  } catch (const hdd::Error& e) {
    return e.status;
  }
}
