// Recursive-descent compiler for the expression strings of Stuff::Functions::Expression and
// Pymor::ParameterFunctional (see expr.hpp).
#include "expr.hpp"

#include <cctype>
#include <cstdlib>
#include <cstring>

#include "common.hpp"

namespace hdd {
namespace {

struct Parser {
  const std::string& s;
  const std::string& var;
  size_t pos = 0;
  Program prog{};
  int n_consts = 0;
  int depth = 0, max_depth = 0;

  Parser(const std::string& text, const std::string& v) : s(text), var(v) { prog.n_ops = 0; }

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
  void emit(Op op, int stack_delta) {
    if (prog.n_ops >= kMaxOps) fail("expression too long");
    prog.op[prog.n_ops] = op;
    prog.cidx[prog.n_ops] = 0;
    ++prog.n_ops;
    depth += stack_delta;
    if (depth > max_depth) max_depth = depth;
    if (max_depth > kMaxStack) fail("expression too deeply nested");
  }
  void emit_const(double v) {
    int slot = -1;
    for (int k = 0; k < n_consts; ++k)
      if (std::memcmp(&prog.cst[k], &v, sizeof(double)) == 0) slot = k;
    if (slot < 0) {
      if (n_consts >= kMaxConsts) fail("too many constants");
      slot = n_consts++;
      prog.cst[slot] = v;
    }
    emit(OP_CONST, +1);
    prog.cidx[prog.n_ops - 1] = static_cast<unsigned char>(slot);
  }

  void expr() {
    term();
    for (;;) {
      if (accept('+')) { term(); emit(OP_ADD, -1); }
      else if (accept('-')) { term(); emit(OP_SUB, -1); }
      else break;
    }
  }
  void term() {
    unary();
    for (;;) {
      if (accept('*')) { unary(); emit(OP_MUL, -1); }
      else if (accept('/')) { unary(); emit(OP_DIV, -1); }
      else break;
    }
  }
  void unary() {
    if (accept('-')) { unary(); emit(OP_NEG, 0); }
    else if (accept('+')) { unary(); }
    else power();
  }
  void power() {
    primary();
    if (accept('^')) { unary(); emit(OP_POW, -1); }
  }
  void primary() {
    skip();
    if (pos >= s.size()) fail("unexpected end");
    const char c = s[pos];
    if (c == '(') {
      ++pos;
      expr();
      if (!accept(')')) fail("expected ')'");
      return;
    }
    if (std::isdigit(static_cast<unsigned char>(c)) || c == '.') {
      char* end = nullptr;
      const double v = std::strtod(s.c_str() + pos, &end);
      if (end == s.c_str() + pos) fail("bad number");
      pos = size_t(end - s.c_str());
      emit_const(v);
      return;
    }
    if (std::isalpha(static_cast<unsigned char>(c)) || c == '_') {
      size_t b = pos;
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
        emit(Op(OP_VAR0 + k), +1);
        return;
      }
      if (id == "pi" || id == "PI" || id == "M_PI") { emit_const(3.14159265358979323846264338327950288); return; }
      struct Fn { const char* name; Op op; int nargs; };
      static const Fn fns[] = {{"sin", OP_SIN, 1},   {"cos", OP_COS, 1},   {"tan", OP_TAN, 1}, {"exp", OP_EXP, 1},
                               {"log", OP_LOG, 1},   {"sqrt", OP_SQRT, 1}, {"abs", OP_ABS, 1}, {"atan", OP_ATAN, 1},
                               {"pow", OP_POW, 2},   {"min", OP_MIN, 2},   {"max", OP_MAX, 2}};
      for (const Fn& f : fns)
        if (id == f.name) {
          if (!accept('(')) fail("expected '(' after function name");
          expr();
          for (int a = 1; a < f.nargs; ++a) {
            if (!accept(',')) fail("expected ','");
            expr();
          }
          if (!accept(')')) fail("expected ')'");
          emit(f.op, f.nargs == 2 ? -1 : 0);
          return;
        }
      fail("unknown identifier '" + id + "'");
    }
    fail(std::string("unexpected character '") + c + "'");
  }
};

}  // namespace

Program compile_expression(const std::string& text, const std::string& var) {
  Parser p(text, var);
  p.expr();
  p.skip();
  if (p.pos != text.size()) p.fail("trailing characters");
  return p.prog;
}

}  // namespace hdd
