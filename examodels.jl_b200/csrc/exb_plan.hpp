// Plan = host-only analysis of a pattern IR + generation of the model's CUDA C++ kernel module.
//
// What is restated here (reference file:line, paths relative to /root/reference/):
//   * the sparsity probe and `===` dedupe that fix o1step / o2step and the two Compressors
//     (src/simdfunction.jl:66-100), by running the reverse passes symbolically;
//   * the running counters that assign o0 / o1 / o2 in add order
//     (src/nlp.jl:1448-1482, 1551-1611, 1680-1738);
//   * the reverse passes themselves -- grpass / jrpass (src/gradient.jl:59-90,
//     src/jacobian.jl:16-40), hrpass0 / hrpass / hdrpass (src/hessian.jl:16-517) -- which are
//     UNROLLED AT GENERATION TIME into straight-line FP64 code per pattern: the recursion
//     over the (static) tree happens here on the host, every adjoint product becomes one SSA
//     temporary, and every leaf visit becomes an accumulation into a register-resident slot.
//     Multiplications by the structural constants 1 / -1 / 0 of the linear operators
//     (`+`, `-`, `c*x`, h12 of `*`) are folded away, which is where most of the
//     reference's per-point FLOPs go.
//
// The device never sees a tree: a pattern is a struct with static member functions
// (val / d1 / d2 / s1 / s2 / row) that the kernel templates of exb_device.cuh call.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "exb_ir.hpp"

namespace exb {

enum { K_REAL, K_NULL, K_VAR, K_N1, K_N2 };
enum { FX_NONE, FX_FIRST, FX_SECOND };

static const int EXB_MAXF_HOST = 16;  // must equal EXB_MAXF in exb_device.cuh
static const int EXB_MAXD_HOST = 4;
static const int EXB_TILE_MAX_NS_HOST = 96;
static const int EXB_CPAT_MAX = 256;  // patterns whose arguments fit the module's constant bank (224 B each)

struct NodeInfo { int kind = K_REAL; int fx = FX_NONE; bool is_int = false; };

struct PatternPlan {
  PatternIR ir;
  std::vector<NodeInfo> info;
  std::vector<int> comp1, comp2;
  int o1step = 0, o2step = 0;
  i64 o0 = 0, o1 = 0, o2 = 0;  // 0-based bases (SIMDFunction offsets)
  i64 oa = 0;                  // aug: first conbuffer slot (ConstraintAugmentation.oa)
  i64 ob = 0;                  // obj: first objbuffer slot
  int ppt0 = 1, ppt1 = 1, ppt2 = 1;   // points per thread in the value / first-order / second-order kernels
  int ppte = 1, ppte1 = 1;            // ... and in the fused evaluation kernels (exb_eval_g0; first-order form exb_eval1_g0)
  std::vector<int> leaf1;      // representative Var IR node per first-order slot
  std::vector<std::pair<int, int>> leaf2;
  // owner-computes gradient (see gen_pattern, g1): every first-order slot's variable index is `t + shift1[j]` with
  // t the value of a range iterator, so variable v receives slot j of point t = v - shift1[j] and nothing else
  bool gather1 = false;
  std::vector<i64> shift1;
  // x window (persistent Hessian kernel): every variable index of the pattern is `t + c` with t the value of a range
  // iterator and xlo <= c <= xhi, so a tile of consecutive points reads one contiguous window of x (and of y)
  bool win = false;
  i64 xlo = 0, xhi = 0;
  // part of x the pattern can read at all: shifts [rlo, rhi] relative to the range value plus fixed indices [flo, fhi]
  // (1-based); xr_ok = false: indices come from iterator data, anything may be read
  bool xr_ok = false, xr_shift = false, xr_fixed = false;
  i64 rlo = 0, rhi = 0, flo = 0, fhi = 0;
  // column-tile kernels (exb_tile_body, duplicate-free Hessian): every variable of the pattern is `t + const`, t the value of
  // a range iterator.  Second-order slot j is the lower-triangle entry (c + t2_d[j], c) in column c = t + t2_cb[j]; t2_r[j] is
  // the rank of t2_d[j] in the model's sorted set of distinct row - column distances (Plan::hd)
  bool tile_ok = false;
  std::vector<i64> t2_cb, t2_d;
  std::vector<int> t2_r;
  i64 t_cbmin = 0, t_cbmax = 0;
  // same for the gradient (exb_tile_body in gradient mode): first-order slot j of an objective pattern lands in variable t + shift1[j]
  bool tgrad = false;
  bool egrad = false;   // the model's only objective pattern with gradient slots, shift-indexed: exb_eval writes g from its own sweep
  i64 g_cbmin = 0, g_cbmax = 0;
};

struct Plan {
  ModelIR m;
  std::vector<PatternPlan> pats;
  i64 ncon = 0, nobj = 0, nconaug = 0, nnzg = 0, nnzj = 0, nnzh = 0;
  std::string source;  // generated module (without the device header)
  std::string error;
  // pattern lists per kernel (indices into pats), fixed at generation time
  std::vector<int> k_hess, k_jac, k_sgrad, k_ggrad, k_cons, k_obj, k_aug, k_eval, k_tgrad;
  // launch entries of exb_hess_g0: a pattern with many second-order slots per point (the 47-slot dynamics row of the COPS rocket:
  // 108 registers and a 48 KB tile per block -> 4 blocks per SM) CAN be split into two entries that evaluate the same points but
  // keep / stage / store one half of the slots each [w0, w1): the other half is dead code in that entry, registers and tile
  // shrink; the forward sweep is done twice.  Opt-in: measured slower (see the kernel-list loop in build_plan)
  std::vector<int> k_hess_l; std::vector<std::pair<int, int>> k_hess_w;
  bool hess_windowed = false;  // every Hessian pattern has an x window: the persistent kernel exb_hessp_g0 is generated too
  // duplicate-free Hessian emitted directly (exb_hessc_g0): possible when every pattern with second-order slots is tile_ok
  bool tile_ok = false;
  std::vector<i64> hd;         // distinct row - column distances of the model's Hessian entries, ascending
  std::vector<i64> h_lo, h_len;   // per distance: the ONE interval of (1-based) columns in which the entry exists
  int tile_halo = 0, tile_ppt = 3;   // tile_ppt: columns per thread (chosen in build_plan; EXB_TUNE_TILE_PPT overrides)
  int tgrad_halo = 0, tgrad_ppt = 4;   // gradient mode of the tile kernel (exb_gradt_g0)
  int egrad_pat = -1;                  // pattern whose gradient the fused evaluation kernels emit themselves (PatternPlan::egrad), or -1
  bool idx32 = false;          // every index (variables, points, slots) fits 31 bits: address arithmetic in 32 bits
  int block = 128, minb = 16;  // launch shape of the generated kernels (tuning knobs: EXB_TUNE_BLOCK / EXB_TUNE_MINB)
};

// ---------------------------------------------------------------------------------------
// static analysis
// ---------------------------------------------------------------------------------------
inline bool real_op_int(int tag, int op, bool a_int, bool b_int) {
  // Julia keeps Int arithmetic among Ints for these; everything else promotes to Float64
  if (tag == T_OP1) return a_int && (op == U_PLUS || op == U_MINUS || op == U_ABS || op == U_ABS2);
  return a_int && b_int && (op == B_ADD || op == B_SUB || op == B_MUL || op == B_MAX || op == B_MIN || op == B_POW);
}

inline void classify(PatternPlan& p) {
  const auto& N = p.ir.nodes;
  p.info.assign(N.size(), NodeInfo());
  for (size_t k = 0; k < N.size(); k++) {
    const IRNode& n = N[k];
    NodeInfo& q = p.info[k];
    switch (n.tag) {
      case T_CONST_I: case T_VAL: case T_DATA_SELF: q.kind = K_REAL; q.is_int = true; break;
      case T_CONST_F: case T_PAR: q.kind = K_REAL; break;
      case T_DATA_FIELD: { q.kind = K_REAL; i64 t = p.ir.fields[(size_t)n.a].type; q.is_int = (t == FT_I64 || t == FT_I32); } break;
      case T_VAR: q.kind = K_VAR; break;                               // graph.jl:397-400,491-494
      case T_NULL: q.kind = K_NULL; break;                             // graph.jl:499-502
      case T_OP1: {
        const NodeInfo& a = p.info[(size_t)n.a];
        if (a.kind == K_REAL) { q.kind = K_REAL; q.is_int = real_op_int(T_OP1, (int)n.payload, a.is_int, false); }
        else q.kind = K_N1;                                            // register.jl:65-71
      } break;
      case T_OP2: {
        const NodeInfo &a = p.info[(size_t)n.a], &b = p.info[(size_t)n.b];
        bool r1 = a.kind == K_REAL, r2 = b.kind == K_REAL;
        if (r1 && r2) { q.kind = K_REAL; q.is_int = real_op_int(T_OP2, (int)n.payload, a.is_int, b.is_int); }
        else if (r2) { q.kind = K_N1; q.fx = FX_SECOND; }              // register.jl:231-248
        else if (r1) { q.kind = K_N1; q.fx = FX_FIRST; }               // register.jl:249-266
        else q.kind = K_N2;                                            // register.jl:209-230
      } break;
    }
  }
}

inline bool ir_equal(const PatternIR& p, int a, int b) {   // Julia `===` on immutable node structs
  if (a == b) return true;
  const IRNode &x = p.nodes[(size_t)a], &y = p.nodes[(size_t)b];
  if (x.tag != y.tag) return false;
  switch (x.tag) {
    case T_CONST_I: case T_CONST_F: case T_NULL: case T_VAL: return x.payload == y.payload;
    case T_DATA_SELF: return true;
    case T_DATA_FIELD: return x.a == y.a;
    case T_VAR: case T_PAR: return ir_equal(p, (int)x.a, (int)y.a);
    case T_OP1: return x.payload == y.payload && ir_equal(p, (int)x.a, (int)y.a);
    case T_OP2: return x.payload == y.payload && ir_equal(p, (int)x.a, (int)y.a) && ir_equal(p, (int)x.b, (int)y.b);
  }
  return false;
}

// Integer index expression as an affine function of t: value = coef * t + cst, where t is the iterator VALUE of a range pattern
// (DATA_SELF) or the POINT NUMBER of an AoS pattern with iota columns (range_start is 0 there, so the two coincide in `pa.start + kg`).
// Only the integer `+ - *` that index expressions are made of (nlp.jl:900-926,2012-2015); anything else -> false.
inline bool affine_index(const PatternIR& p, int n, i64& coef, i64& cst) {
  const IRNode& q = p.nodes[(size_t)n];
  auto small = [](i64 v) { return v > -(1LL << 40) && v < (1LL << 40); };
  switch (q.tag) {
    case T_CONST_I: case T_VAL: coef = 0; cst = q.payload; return small(cst);
    case T_DATA_SELF: coef = 1; cst = 0; return true;
    case T_DATA_FIELD:   // an iota column (detect_iota): value = point number + iota0, like a range whose start is iota0
      if ((size_t)q.a < p.iota.size() && p.iota[(size_t)q.a]) { coef = 1; cst = p.iota0[(size_t)q.a]; return true; }
      return false;
    case T_OP1: {
      i64 a, b;
      if (!affine_index(p, (int)q.a, a, b)) return false;
      if (q.payload == U_PLUS) { coef = a; cst = b; return true; }
      if (q.payload == U_MINUS) { coef = -a; cst = -b; return true; }
      return false;
    }
    case T_OP2: {
      i64 a1, b1, a2, b2;
      if (!affine_index(p, (int)q.a, a1, b1) || !affine_index(p, (int)q.b, a2, b2)) return false;
      if (q.payload == B_ADD) { coef = a1 + a2; cst = b1 + b2; return small(coef) && small(cst); }
      if (q.payload == B_SUB) { coef = a1 - a2; cst = b1 - b2; return small(coef) && small(cst); }
      if (q.payload == B_MUL && (a1 == 0 || a2 == 0)) {
        if (a1 == 0) { coef = b1 * a2; cst = b1 * b2; } else { coef = a1 * b2; cst = b1 * b2; }
        return small(coef) && small(cst) && small(b1) && small(b2);
      }
      return false;
    }
  }
  return false;
}
// Can two variable index expressions be decided equal / unequal for EVERY data point?  1 always equal, 0 never, -1 unknown.
inline bool shiftable(const PatternIR& p) { return p.itr_kind == ITR_RANGE || p.has_iota(); }
inline int index_relation(const PatternIR& p, int a, int b) {
  if (ir_equal(p, a, b)) return 1;
  i64 c1, k1, c2, k2;
  if (shiftable(p) && affine_index(p, a, c1, k1) && affine_index(p, b, c2, k2)) {
    if (c1 == c2) return k1 == k2 ? 1 : 0;
    // (c1 - c2) t = k2 - k1 has at most one integer solution: only decidable when it has none
    const i64 dc = c1 - c2, dk = k2 - k1;
    if (dk % dc != 0) return 0;
    const i64 t = dk / dc;
    if (t < p.range_start || t >= p.range_start + p.nitr) return 0;
  }
  return -1;
}

// ---------------------------------------------------------------------------------------
// symbolic scalar: either a compile-time constant or the name of an SSA temporary
// ---------------------------------------------------------------------------------------
struct Ex { std::string s; bool c = false; double v = 0.0; };

inline std::string dlit(double v) {
  if (std::isnan(v)) return "exb_nan()";
  if (std::isinf(v)) return v > 0 ? "exb_inf()" : "(-exb_inf())";
  char buf[64];
  snprintf(buf, sizeof buf, "%.17g", v);
  std::string s = buf;
  if (s.find_first_of(".eE") == std::string::npos) s += ".0";
  if (v < 0 || (v == 0 && std::signbit(v))) s = "(" + s + ")";
  return s;
}
inline Ex K(double v) { Ex e; e.c = true; e.v = v; e.s = dlit(v); return e; }
inline Ex Sym(const std::string& s) { Ex e; e.s = s; return e; }

// A generated function body: SSA lines with hash-consing of identical right-hand sides.
struct Body {
  std::vector<std::string> lines;
  std::map<std::string, std::string> memo;
  int ntmp = 0;
  bool quiet = false;  // probe mode: no code is produced
  std::string tmp(const std::string& type, const std::string& rhs) {
    if (quiet) return "_";
    auto it = memo.find(rhs);
    if (it != memo.end()) return it->second;
    std::string name = "t" + std::to_string(ntmp++);
    lines.push_back("const " + type + " " + name + " = " + rhs + ";");
    memo[rhs] = name;
    return name;
  }
  void raw(const std::string& l) { if (!quiet) lines.push_back(l); }
  Ex mul(const Ex& a, const Ex& b) {
    if (a.c && b.c) return K(a.v * b.v);
    if (a.c) { if (a.v == 0) return K(0); if (a.v == 1) return b; if (a.v == -1) return neg(b); }
    if (b.c) { if (b.v == 0) return K(0); if (b.v == 1) return a; if (b.v == -1) return neg(a); }
    return Sym(tmp("double", a.s + " * " + b.s));
  }
  Ex add(const Ex& a, const Ex& b) {
    if (a.c && b.c) return K(a.v + b.v);
    if (a.c && a.v == 0) return b;
    if (b.c && b.v == 0) return a;
    return Sym(tmp("double", a.s + " + " + b.s));
  }
  Ex sub(const Ex& a, const Ex& b) {
    if (a.c && b.c) return K(a.v - b.v);
    if (b.c && b.v == 0) return a;
    if (a.c && a.v == 0) return neg(b);
    return Sym(tmp("double", a.s + " - " + b.s));
  }
  Ex neg(const Ex& a) {
    if (a.c) return K(-a.v);
    return Sym(tmp("double", "-" + a.s));
  }
  Ex sq(const Ex& a) { return mul(a, a); }
};

// forward image of one IR node inside one generated function
struct NV {
  bool done = false;
  // K_REAL
  bool is_int = false, lit = false; i64 iv = 0; std::string rs;  // rs: C++ expression (name or literal)
  // value kinds
  Ex x, y1, y2, h11, h12, h22;
  std::string idx;  // K_VAR: name of the evaluated (1-based) variable index
};

struct Gen {
  const PatternPlan& p;
  Body& B;
  int order;  // 0 value, 1 first-order tape, 2 second-order tape
  std::vector<NV> nv;
  // leaf sinks
  int cnt = 0;
  std::vector<Ex> slot;                           // symbolic accumulators
  std::vector<int>* raw1 = nullptr;               // probe outputs
  std::vector<std::pair<int, int>>* raw2 = nullptr;
  const std::vector<int>* comp = nullptr;

  Gen(const PatternPlan& pp, Body& b, int ord) : p(pp), B(b), order(ord), nv(pp.ir.nodes.size()) {}

  const IRNode& N(int n) const { return p.ir.nodes[(size_t)n]; }
  const NodeInfo& I(int n) const { return p.info[(size_t)n]; }
  int inner_of(int n) const { return (N(n).tag == T_OP2 && I(n).fx == FX_FIRST) ? (int)N(n).b : (int)N(n).a; }

  // ---- Real (variable-free) subtrees: graph.jl:305-318, register.jl:70,268-273 ----------
  static std::string ilit(i64 v) { return v < 0 ? "(" + std::to_string(v) + "LL)" : std::to_string(v) + "LL"; }
  std::string as_double(const NV& r) {
    if (r.is_int) return r.lit ? dlit((double)r.iv) : "(double)" + r.rs;
    return r.rs;
  }
  Ex real_ex(int n) {   // a Real node as a double-valued symbolic scalar
    NV& r = real(n);
    if (r.lit) return K(r.is_int ? (double)r.iv : r.x.v);
    return Sym(as_double(r));
  }
  NV& real(int n) {
    NV& r = nv[(size_t)n];
    if (r.done) return r;
    r.done = true;
    const IRNode& q = N(n);
    r.is_int = I(n).is_int;
    switch (q.tag) {
      case T_CONST_I: case T_VAL: r.lit = true; r.iv = q.payload; r.rs = ilit(q.payload); break;
      case T_CONST_F: { double v; std::memcpy(&v, &q.payload, 8); r.lit = true; r.x = K(v); r.rs = dlit(v); } break;
      case T_DATA_SELF: r.rs = B.tmp("long long", "pa.start + kg"); break;
      case T_DATA_FIELD:
        if (r.is_int && (size_t)q.a < p.ir.iota.size() && p.ir.iota[(size_t)q.a])   // iota column: never loaded
          r.rs = B.tmp("long long", "kg + " + ilit(p.ir.iota0[(size_t)q.a]));
        else if (r.is_int) r.rs = B.tmp("long long", "exb_ld_i(pa, " + std::to_string(q.a) + ", EXB_IX(kg))");
        else r.rs = B.tmp("double", "exb_ld_f(pa, " + std::to_string(q.a) + ", EXB_IX(kg))");
        break;
      case T_PAR: { NV& ix = real((int)q.a); r.rs = B.tmp("double", "__ldg(th + EXB_IX(" + ix.rs + " - 1))"); } break;  // graph.jl:310-311
      case T_OP1: {
        NV& a = real((int)q.a);
        int op = (int)q.payload;
        if (r.is_int) {
          const char* f = op == U_PLUS ? "+" : op == U_MINUS ? "-" : nullptr;
          if (f) r.rs = B.tmp("long long", std::string(f) + a.rs);
          else if (op == U_ABS) r.rs = B.tmp("long long", "(" + a.rs + " < 0 ? -" + a.rs + " : " + a.rs + ")");
          else r.rs = B.tmp("long long", a.rs + " * " + a.rs);
        } else {
          r.rs = B.tmp("double", "exb_f1<" + std::to_string(op) + ", SLOW>(" + as_double(a) + ", bad)");
        }
      } break;
      case T_OP2: {
        NV &a = real((int)q.a), &b = real((int)q.b);
        int op = (int)q.payload;
        if (r.is_int) {
          switch (op) {
            case B_ADD: r.rs = B.tmp("long long", a.rs + " + " + b.rs); break;
            case B_SUB: r.rs = B.tmp("long long", a.rs + " - " + b.rs); break;
            case B_MUL: r.rs = B.tmp("long long", a.rs + " * " + b.rs); break;
            case B_MAX: r.rs = B.tmp("long long", "exb_imax(" + a.rs + ", " + b.rs + ")"); break;
            case B_MIN: r.rs = B.tmp("long long", "exb_imin(" + a.rs + ", " + b.rs + ")"); break;
            default: r.rs = B.tmp("long long", "exb_ipow(" + a.rs + ", " + b.rs + ")");
          }
        } else if (op == B_POW && b.is_int) {   // Float64 ^ Int
          r.rs = B.tmp("double", "exb_powi(" + as_double(a) + ", " + b.rs + ")");
        } else {
          r.rs = B.tmp("double", "exb_f2<" + std::to_string(op) + ">(" + as_double(a) + ", " + as_double(b) + ")");
        }
      } break;
      default: r.rs = "exb_nan()";
    }
    return r;
  }

  // ---- forward sweep over value kinds (graph.jl:397-400,491-494; register.jl:65-68,174-266) ----
  std::string fresh(const char* pre, int n) { return std::string(pre) + std::to_string(n); }
  NV& fwd(int n) {
    NV& r = nv[(size_t)n];
    if (r.done) return r;
    const IRNode& q = N(n);
    const NodeInfo& in = I(n);
    if (in.kind == K_REAL) { NV& rr = real(n); rr.x = real_ex(n); return rr; }
    r.done = true;
    const std::string O = std::to_string(order);
    if (in.kind == K_NULL) { double v; std::memcpy(&v, &q.payload, 8); r.x = K(v); return r; }
    if (in.kind == K_VAR) {
      NV& ix = real((int)q.a);
      r.idx = ix.rs;
      // a variable at a FIXED index (e.g. a step length shared by every point) is outside any per-tile window of x
      i64 cf = 1, ct = 0;
      const bool fixed = affine_index(p.ir, (int)q.a, cf, ct) && cf == 0;
      r.x = Sym(B.tmp("double", std::string(fixed ? "x.ldc(" : "x.ld(") + "EXB_IX(" + ix.rs + " - 1))"));
      return r;
    }
    if (q.tag == T_OP1) {
      NV& a = fwd((int)q.a);
      int op = (int)q.payload;
      switch (op) {
        case U_PLUS: r.x = a.x; r.y1 = K(1); r.h11 = K(0); return r;
        case U_MINUS: r.x = B.neg(a.x); r.y1 = K(-1); r.h11 = K(0); return r;
        case U_ABS2: r.x = B.mul(a.x, a.x); r.y1 = B.mul(K(2), a.x); r.h11 = K(2); return r;
      }
      if (order == 0) { r.x = Sym(B.tmp("double", "exb_f1<" + std::to_string(op) + ", SLOW>(" + a.x.s + ", bad)")); return r; }
      std::string f = fresh("f", n), d = fresh("d", n), dd = fresh("dd", n);
      B.raw("double " + f + ", " + d + ", " + dd + "; exb_uni<" + std::to_string(op) + ", " + O + ", SLOW>(" + a.x.s + ", " + f + ", " + d + ", " + dd + ", bad);");
      r.x = Sym(f); r.y1 = Sym(d); r.h11 = Sym(dd);
      switch (op) {   // structurally constant derivatives
        case U_ABS: case U_DEG2RAD: case U_RAD2DEG: r.h11 = K(0); break;
        case U_SIGN: case U_SIGNBIT: case U_FLOOR: case U_CEIL: r.y1 = K(0); r.h11 = K(0); break;
      }
      return r;
    }
    // T_OP2
    int op = (int)q.payload;
    if (in.kind == K_N2) {
      NV &a = fwd((int)q.a), &b = fwd((int)q.b);
      switch (op) {
        case B_ADD: r.x = B.add(a.x, b.x); r.y1 = K(1); r.y2 = K(1); r.h11 = r.h12 = r.h22 = K(0); return r;
        case B_SUB: r.x = B.sub(a.x, b.x); r.y1 = K(1); r.y2 = K(-1); r.h11 = r.h12 = r.h22 = K(0); return r;
        case B_MUL: r.x = B.mul(a.x, b.x); r.y1 = b.x; r.y2 = a.x; r.h11 = K(0); r.h12 = K(1); r.h22 = K(0); return r;
      }
      if (order == 0) { r.x = Sym(B.tmp("double", "exb_f2<" + std::to_string(op) + ">(" + a.x.s + ", " + b.x.s + ")")); return r; }
      std::string f = fresh("f", n), u1 = fresh("ya", n), u2 = fresh("yb", n), g11 = fresh("haa", n), g12 = fresh("hab", n), g22 = fresh("hbb", n);
      B.raw("double " + f + ", " + u1 + ", " + u2 + ", " + g11 + ", " + g12 + ", " + g22 + "; exb_bi<" + std::to_string(op) + ", " + O + ">(" +
            a.x.s + ", " + b.x.s + ", " + f + ", " + u1 + ", " + u2 + ", " + g11 + ", " + g12 + ", " + g22 + ");");
      r.x = Sym(f); r.y1 = Sym(u1); r.y2 = Sym(u2); r.h11 = Sym(g11); r.h12 = Sym(g12); r.h22 = Sym(g22);
      if (op == B_DIV) r.h11 = K(0);
      if (op == B_MAX || op == B_MIN) r.h11 = r.h12 = r.h22 = K(0);
      return r;
    }
    // one Real operand: a Node1 carrying (f, df, ddf) w.r.t. the node operand only
    const bool second_fixed = in.fx == FX_SECOND;   // node OP real
    NV& a = fwd(second_fixed ? (int)q.a : (int)q.b);
    const int rn = second_fixed ? (int)q.b : (int)q.a;
    NV& cr = real(rn);
    Ex c = real_ex(rn);
    switch (op) {
      case B_ADD: r.x = second_fixed ? B.add(a.x, c) : B.add(c, a.x); r.y1 = K(1); r.h11 = K(0); return r;
      case B_SUB:
        if (second_fixed) { r.x = B.sub(a.x, c); r.y1 = K(1); } else { r.x = B.sub(c, a.x); r.y1 = K(-1); }
        r.h11 = K(0); return r;
      case B_MUL: r.x = second_fixed ? B.mul(a.x, c) : B.mul(c, a.x); r.y1 = c; r.h11 = K(0); return r;
    }
    if (op == B_DIV && second_fixed) {   // x / c: (1/c, 0)
      r.x = Sym(B.tmp("double", a.x.s + " / " + c.s));
      r.y1 = c.c ? K(1.0 / c.v) : Sym(B.tmp("double", "1.0 / " + c.s));
      r.h11 = K(0);
      return r;
    }
    if (op == B_POW && second_fixed) {
      if (order == 0) {
        r.x = cr.is_int ? Sym(B.tmp("double", "exb_powi(" + a.x.s + ", " + cr.rs + ")"))
                        : Sym(B.tmp("double", "pow(" + a.x.s + ", " + c.s + ")"));
        return r;
      }
      std::string f = fresh("f", n), d = fresh("d", n), dd = fresh("dd", n);
      if (cr.is_int) B.raw("double " + f + ", " + d + ", " + dd + "; exb_pow_int<" + O + ">(" + a.x.s + ", " + cr.rs + ", " + f + ", " + d + ", " + dd + ");");
      else B.raw("double " + f + ", " + d + ", " + dd + "; exb_pow_flt<" + O + ">(" + a.x.s + ", " + c.s + ", " + f + ", " + d + ", " + dd + ");");
      r.x = Sym(f); r.y1 = Sym(d); r.h11 = Sym(dd);
      if (cr.is_int && cr.lit && (cr.iv == 0 || cr.iv == 1)) r.h11 = K(0);
      return r;
    }
    // generic: evaluate the full bivariate table and keep the operand's entries
    {
      const std::string x1 = second_fixed ? a.x.s : c.s, x2 = second_fixed ? c.s : a.x.s;
      if (order == 0) { r.x = Sym(B.tmp("double", "exb_f2<" + std::to_string(op) + ">(" + x1 + ", " + x2 + ")")); return r; }
      std::string f = fresh("f", n), u1 = fresh("ya", n), u2 = fresh("yb", n), g11 = fresh("haa", n), g12 = fresh("hab", n), g22 = fresh("hbb", n);
      B.raw("double " + f + ", " + u1 + ", " + u2 + ", " + g11 + ", " + g12 + ", " + g22 + "; exb_bi<" + std::to_string(op) + ", " + O + ">(" +
            x1 + ", " + x2 + ", " + f + ", " + u1 + ", " + u2 + ", " + g11 + ", " + g12 + ", " + g22 + ");");
      r.x = Sym(f);
      r.y1 = Sym(second_fixed ? u1 : u2);
      r.h11 = Sym(second_fixed ? g11 : g22);
      if (op == B_MAX || op == B_MIN) r.h11 = K(0);
      return r;
    }
  }

  // ---- leaf sinks ------------------------------------------------------------------------
  void leaf1(int n, const Ex& adj) {
    if (raw1) { raw1->push_back(n); return; }
    int j = (*comp)[(size_t)cnt++] - 1;
    slot[(size_t)j] = B.add(slot[(size_t)j], adj);
  }
  void leafd(int n, const Ex& adj2) {   // hessian.jl:580-592 (values), :533-536 (probe)
    if (raw2) { raw2->push_back(std::make_pair(n, n)); return; }
    int j = (*comp)[(size_t)cnt++] - 1;
    slot[(size_t)j] = B.add(slot[(size_t)j], adj2);
  }
  void leaf2(int a, int b, const Ex& adj) {   // hessian.jl:251-268 (values), :520-532 (probe)
    if (raw2) { raw2->push_back(std::make_pair(a, b)); return; }
    int j = (*comp)[(size_t)cnt++] - 1;
    if (adj.c && adj.v == 0) return;
    Ex v;
    const int rel = index_relation(p.ir, (int)N(a).a, (int)N(b).a);
    if (rel == 1) v = B.mul(K(2), adj);     // i == j for every point (structurally, or equal affine index maps)
    else if (rel == 0) v = adj;             // affine index maps that never meet: the compare is decided at build time
    else v = Sym(B.tmp("double", "exb_twice_if_eq(EXB_IX(" + nv[(size_t)a].idx + "), EXB_IX(" + nv[(size_t)b].idx + "), " + adj.s + ")"));
    slot[(size_t)j] = B.add(slot[(size_t)j], v);
  }

  // ---- reverse passes, unrolled symbolically ----------------------------------------------
  const NV& T(int n) const { return nv[(size_t)n]; }
  void rpass1(int n, const Ex& adj) {   // gradient.jl:59-90, jacobian.jl:16-40
    switch (I(n).kind) {
      case K_REAL: case K_NULL: return;
      case K_N1: rpass1(inner_of(n), B.mul(adj, T(n).y1)); return;
      case K_N2:
        rpass1((int)N(n).a, B.mul(adj, T(n).y1));
        rpass1((int)N(n).b, B.mul(adj, T(n).y2));
        return;
      case K_VAR: leaf1(n, adj); return;
    }
  }
  void hdrpass(int a, int b, const Ex& adj) {   // hessian.jl:16-320
    const int k1 = I(a).kind, k2 = I(b).kind;
    if (k1 == K_NULL || k2 == K_NULL || k1 == K_REAL || k2 == K_REAL) return;   // :318-320
    if (k1 == K_VAR && k2 == K_VAR) { leaf2(a, b, adj); return; }                // :251-268
    if (k1 == K_N1 && k2 == K_N1) { hdrpass(inner_of(a), inner_of(b), B.mul(B.mul(adj, T(a).y1), T(b).y1)); return; }   // :16-28
    if (k1 == K_VAR && k2 == K_N1) { hdrpass(a, inner_of(b), B.mul(adj, T(b).y1)); return; }   // :44-56
    if (k1 == K_N1 && k2 == K_VAR) { hdrpass(inner_of(a), b, B.mul(adj, T(a).y1)); return; }   // :72-84
    if (k1 == K_N2 && k2 == K_N2) {                                              // :100-115
      hdrpass((int)N(a).a, (int)N(b).a, B.mul(B.mul(adj, T(a).y1), T(b).y1));
      hdrpass((int)N(a).a, (int)N(b).b, B.mul(B.mul(adj, T(a).y1), T(b).y2));
      hdrpass((int)N(a).b, (int)N(b).a, B.mul(B.mul(adj, T(a).y2), T(b).y1));
      hdrpass((int)N(a).b, (int)N(b).b, B.mul(B.mul(adj, T(a).y2), T(b).y2));
      return;
    }
    if (k1 == K_N1 && k2 == K_N2) {                                              // :134-147
      hdrpass(inner_of(a), (int)N(b).a, B.mul(B.mul(adj, T(a).y1), T(b).y1));
      hdrpass(inner_of(a), (int)N(b).b, B.mul(B.mul(adj, T(a).y1), T(b).y2));
      return;
    }
    if (k1 == K_N2 && k2 == K_N1) {                                              // :163-176
      hdrpass((int)N(a).a, inner_of(b), B.mul(B.mul(adj, T(a).y1), T(b).y1));
      hdrpass((int)N(a).b, inner_of(b), B.mul(B.mul(adj, T(a).y2), T(b).y1));
      return;
    }
    if (k1 == K_VAR && k2 == K_N2) {                                             // :192-205
      hdrpass(a, (int)N(b).a, B.mul(adj, T(b).y1));
      hdrpass(a, (int)N(b).b, B.mul(adj, T(b).y2));
      return;
    }
    if (k1 == K_N2 && k2 == K_VAR) {                                             // :221-234
      hdrpass((int)N(a).a, b, B.mul(adj, T(a).y1));
      hdrpass((int)N(a).b, b, B.mul(adj, T(a).y2));
      return;
    }
  }
  void hrpass(int n, const Ex& adj, const Ex& adj2) {   // hessian.jl:337-380
    switch (I(n).kind) {
      case K_REAL: case K_NULL: return;
      case K_N1:
        hrpass(inner_of(n), B.mul(adj, T(n).y1), B.add(B.mul(adj2, B.sq(T(n).y1)), B.mul(adj, T(n).h11)));
        return;
      case K_N2: {
        Ex cross = B.add(B.mul(B.mul(adj2, T(n).y1), T(n).y2), B.mul(adj, T(n).h12));
        hrpass((int)N(n).a, B.mul(adj, T(n).y1), B.add(B.mul(adj2, B.sq(T(n).y1)), B.mul(adj, T(n).h11)));
        hrpass((int)N(n).b, B.mul(adj, T(n).y2), B.add(B.mul(adj2, B.sq(T(n).y2)), B.mul(adj, T(n).h22)));
        hdrpass((int)N(n).a, (int)N(n).b, cross);
        return;
      }
      case K_VAR: leafd(n, adj2); return;
    }
  }
  void hrpass0(int n, const Ex& adj, const Ex& adj2) {   // hessian.jl:382-517
    const NodeInfo& in = I(n);
    if (in.kind == K_VAR) return;                                                // :494-517
    if (in.kind == K_N1) {
      int c = inner_of(n);
      int op = (int)N(n).payload;
      if (N(n).tag == T_OP2) {
        if (op == B_MUL) { hrpass0(c, B.mul(adj, T(n).y1), B.mul(adj2, B.sq(T(n).y1))); return; }   // :385-397
        if (op == B_ADD) { hrpass0(c, adj, adj2); return; }                                        // :398-410
        if (op == B_SUB && in.fx == FX_FIRST) { hrpass0(c, B.neg(adj), adj2); return; }            // :411-423
        if (op == B_SUB && in.fx == FX_SECOND) { hrpass0(c, adj, adj2); return; }                  // :424-436
      } else {
        if (op == U_PLUS) { hrpass0(c, adj, adj2); return; }                     // :438-450
        if (op == U_MINUS) { hrpass0(c, B.neg(adj), adj2); return; }             // :451-463
      }
    }
    if (in.kind == K_N2 && N(n).payload == B_ADD) {                              // :465-478
      hrpass0((int)N(n).a, adj, adj2); hrpass0((int)N(n).b, adj, adj2); return;
    }
    if (in.kind == K_N2 && N(n).payload == B_SUB) {                              // :480-493
      hrpass0((int)N(n).a, adj, adj2); hrpass0((int)N(n).b, B.neg(adj), adj2); return;
    }
    hrpass(n, adj, adj2);                                                        // :382
  }
};

// ---------------------------------------------------------------------------------------
// probe + counters
// ---------------------------------------------------------------------------------------
inline void probe(PatternPlan& p) {   // simdfunction.jl:66-100
  Body B; B.quiet = true;
  std::vector<int> raw1; std::vector<std::pair<int, int>> raw2;
  {
    Gen g(p, B, 1); g.raw1 = &raw1;
    g.rpass1(p.ir.root, K(1));
  }
  {
    Gen g(p, B, 2); g.raw2 = &raw2;
    g.hrpass0(p.ir.root, Sym("_"), K(0));
  }
  p.comp1.clear(); p.comp2.clear(); p.leaf1.clear(); p.leaf2.clear();
  for (int v : raw1) {   // _ident_unique, :66-76
    int found = -1;
    for (size_t q = 0; q < p.leaf1.size(); q++) if (ir_equal(p.ir, p.leaf1[q], v)) { found = (int)q; break; }
    if (found < 0) { p.leaf1.push_back(v); found = (int)p.leaf1.size() - 1; }
    p.comp1.push_back(found + 1);
  }
  p.o1step = (int)p.leaf1.size();
  for (auto& v : raw2) {
    int found = -1;
    for (size_t q = 0; q < p.leaf2.size(); q++)
      if (ir_equal(p.ir, p.leaf2[q].first, v.first) && ir_equal(p.ir, p.leaf2[q].second, v.second)) { found = (int)q; break; }
    if (found < 0) { p.leaf2.push_back(v); found = (int)p.leaf2.size() - 1; }
    p.comp2.push_back(found + 1);
  }
  p.o2step = (int)p.leaf2.size();
}

// ---------------------------------------------------------------------------------------
// code generation
// ---------------------------------------------------------------------------------------
inline void emit_fn(std::ostringstream& o, const std::string& sig, const Body& B, const std::vector<std::string>& tail) {
  o << "  __device__ static __forceinline__ " << sig << " {\n";
  o << "    constexpr bool SLOW = false; bool bad = false; (void)SLOW; (void)bad;\n";
  for (auto& l : B.lines) o << "    " << l << "\n";
  for (auto& l : tail) o << "    " << l << "\n";
  o << "  }\n";
}
// does the body call a routine with a range-limited fast path (sin, cos, exp: exb_sincos / exb_exp in exb_device.cuh)?
inline bool body_uses_fast(const Body& B) {
  for (auto& l : B.lines)
    for (int op : {(int)U_EXP, (int)U_SIN, (int)U_COS}) {
      const std::string k = "<" + std::to_string(op) + ", ";
      if (l.find("exb_uni" + k) != std::string::npos || l.find("exb_f1" + k) != std::string::npos) return true;
    }
  return false;
}
// A value / derivative function of a pattern: `NAME_t<SLOW>` holds the body; NAME evaluates the fast form and, when some
// transcendental argument was out of its fast range (rare: |a| >= 2^31, |x| >= 708, inf, nan), re-evaluates the point
// through the __noinline__ NAME_slow built on libdevice's full-range routines.  `ns` = 0: returns double; else writes s[ns].
inline void emit_eval_fn(std::ostringstream& o, const std::string& name, const std::string& params, const std::string& args, int ns,
                         const Body& B, const std::vector<std::string>& tail) {
  // XA = how x is read: ExbXG (global memory, __ldg) or ExbXS (the block's shared-memory window, exb_hessp_body)
  const std::string ret = ns == 0 ? "double" : "void";
  const std::string sp = ns == 0 ? "" : ", double (&s)[" + std::to_string(ns) + "]";
  o << "  template <bool SLOW, class XA> __device__ static __forceinline__ " << ret << " " << name << "_t(" << params << sp << ", bool& bad) {\n";
  for (auto& l : B.lines) o << "    " << l << "\n";
  for (auto& l : tail) o << "    " << l << "\n";
  o << "  }\n";
  const bool fast = body_uses_fast(B);
  if (fast) {
    if (ns == 0) o << "  template <class XA> __device__ static __noinline__ double " << name << "_slow(" << params << ") { bool bad = false; return " << name << "_t<true>(" << args << ", bad); }\n";
    else o << "  template <class XA> __device__ static __noinline__ void " << name << "_slow(" << params << ", double* __restrict__ so) { bool bad = false; double s[" << ns << "]; "
           << name << "_t<true>(" << args << ", s, bad); for (int j = 0; j < " << ns << "; j++) so[j] = s[j]; }\n";
  }
  o << "  template <class XA> __device__ static __forceinline__ " << ret << " " << name << "(" << params << sp << ") {\n    bool bad = false;\n";
  if (ns == 0) {
    o << "    double r = " << name << "_t<false>(" << args << ", bad);\n";
    if (fast) o << "    if (bad) r = " << name << "_slow(" << args << ");\n";
    o << "    return r;\n";
  } else {
    o << "    " << name << "_t<false>(" << args << ", s, bad);\n";
    if (fast) o << "    if (bad) { double q[" << ns << "]; " << name << "_slow(" << args << ", q); for (int j = 0; j < " << ns << "; j++) s[j] = q[j]; }\n";
  }
  o << "  }\n";
}
// Fused evaluation of one point: value, first-order slots (adjoint 1) and second-order slots (adjoint a0) from ONE forward
// sweep (exb_eval_block).  Same fast / slow split as emit_eval_fn.
inline void emit_eval3_fn(std::ostringstream& o, const std::string& params, const std::string& args, int n1, int n2, const Body& B,
                          const std::vector<std::string>& tail, const std::string& fn = "d012") {
  const std::string outs = ", double& v, double (&s1)[" + std::to_string(n1) + "], double (&s2)[" + std::to_string(n2) + "]";
  o << "  template <bool SLOW, class XA> __device__ static __forceinline__ void " << fn << "_t(" << params << outs << ", bool& bad) {\n";
  for (auto& l : B.lines) o << "    " << l << "\n";
  for (auto& l : tail) o << "    " << l << "\n";
  o << "  }\n";
  const bool fast = body_uses_fast(B);
  if (fast)
    o << "  template <class XA> __device__ static __noinline__ void " << fn << "_slow(" << params << ", double* __restrict__ so) { bool bad = false; double v; double s1[" << n1
      << "], s2[" << n2 << "]; " << fn << "_t<true>(" << args << ", v, s1, s2, bad); so[0] = v; for (int j = 0; j < " << n1 << "; j++) so[1 + j] = s1[j]; for (int j = 0; j < "
      << n2 << "; j++) so[" << 1 + n1 << " + j] = s2[j]; }\n";
  o << "  template <class XA> __device__ static __forceinline__ void " << fn << "(" << params << outs << ") {\n    bool bad = false;\n";
  o << "    " << fn << "_t<false>(" << args << ", v, s1, s2, bad);\n";
  if (fast)
    o << "    if (bad) { double q[" << 1 + n1 + n2 << "]; " << fn << "_slow(" << args << ", q); v = q[0]; for (int j = 0; j < " << n1 << "; j++) s1[j] = q[1 + j]; for (int j = 0; j < "
      << n2 << "; j++) s2[j] = q[" << 1 + n1 << " + j]; }\n";
  o << "  }\n";
}
// Points per thread: cheap bodies are dominated by the per-block prologue / tile-store epilogue, so they get
// several points per thread (also more loads in flight per thread); heavy bodies keep one.
inline int body_weight(const Body& B) {
  int w = 0;
  for (auto& l : B.lines) w += (l.find("exb_uni<") != std::string::npos || l.find("exb_bi<") != std::string::npos ||
                                l.find("exb_pow_") != std::string::npos || l.find("exb_f1<") != std::string::npos ||
                                l.find("exb_f2<") != std::string::npos) ? 30 : 1;
  return w;
}
inline int ppt_for(int weight, int ns) {
  static const int w4 = getenv("EXB_TUNE_PPT_W4") ? atoi(getenv("EXB_TUNE_PPT_W4")) : 60;    // dev knobs
  static const int w2 = getenv("EXB_TUNE_PPT_W2") ? atoi(getenv("EXB_TUNE_PPT_W2")) : 150;
  int p = weight <= w4 ? 4 : weight <= w2 ? 2 : 1;
  while (p > 1 && p * ns > 16) p >>= 1;   // slots live in registers
  return p;
}

// x window of a pattern (see PatternPlan::win)
inline void compute_window(PatternPlan& p) {
  p.win = shiftable(p.ir) && p.ir.kind != KIND_AUG && p.o2step > 0;
  bool first = true;
  for (size_t q = 0; q < p.ir.nodes.size() && p.win; q++) {
    if (p.ir.nodes[q].tag != T_VAR) continue;
    i64 cf, ct;
    if (!affine_index(p.ir, (int)p.ir.nodes[q].a, cf, ct) || (cf != 1 && cf != 0)) { p.win = false; break; }
    if (cf == 0) continue;   // fixed index: read from global memory (ExbXS::ldc)
    if (first || ct < p.xlo) p.xlo = ct;
    if (first || ct > p.xhi) p.xhi = ct;
    first = false;
  }
  if (first || p.xhi - p.xlo > 4096) p.win = false;
}

inline void compute_xrange(PatternPlan& p) {
  p.xr_ok = true; p.xr_shift = p.xr_fixed = false;
  for (size_t q = 0; q < p.ir.nodes.size() && p.xr_ok; q++) {
    if (p.ir.nodes[q].tag != T_VAR) continue;
    i64 cf, ct;
    if (!affine_index(p.ir, (int)p.ir.nodes[q].a, cf, ct)) { p.xr_ok = false; break; }
    if (cf == 1 && shiftable(p.ir)) {
      if (!p.xr_shift || ct < p.rlo) p.rlo = ct;
      if (!p.xr_shift || ct > p.rhi) p.rhi = ct;
      p.xr_shift = true;
    } else if (cf == 0) {
      if (!p.xr_fixed || ct < p.flo) p.flo = ct;
      if (!p.xr_fixed || ct > p.fhi) p.fhi = ct;
      p.xr_fixed = true;
    } else p.xr_ok = false;
  }
}

// shift analysis for the column-tile kernels (see PatternPlan::tile_ok)
inline void compute_tile(PatternPlan& p) {
  p.tile_ok = shiftable(p.ir) && p.ir.kind != KIND_AUG;
  p.t2_cb.clear(); p.t2_d.clear(); p.t2_r.clear();
  if (p.o2step == 0) return;
  bool first = true;
  for (int j = 0; j < p.o2step && p.tile_ok; j++) {
    i64 ca, ka, cb, kb;
    const int na = (int)p.ir.nodes[(size_t)p.leaf2[(size_t)j].first].a, nb = (int)p.ir.nodes[(size_t)p.leaf2[(size_t)j].second].a;
    if (!affine_index(p.ir, na, ca, ka) || !affine_index(p.ir, nb, cb, kb) || ca != 1 || cb != 1) { p.tile_ok = false; break; }
    const i64 lo = std::min(ka, kb), d = std::max(ka, kb) - lo;
    p.t2_cb.push_back(lo); p.t2_d.push_back(d);
    if (first || lo < p.t_cbmin) p.t_cbmin = lo;
    if (first || lo > p.t_cbmax) p.t_cbmax = lo;
    first = false;
  }
  if (p.tile_ok && p.t_cbmax - p.t_cbmin > 64) p.tile_ok = false;   // halo of re-evaluated points per tile
}

inline std::string gen_pattern(PatternPlan& p, int index, bool windowed, const std::vector<i64>* hd = nullptr, int tile_stride = 0, bool sole_obj = false) {
  std::ostringstream o;
  const int ns1 = p.o1step, ns2 = p.o2step;
  const int a1 = ns1 > 0 ? ns1 : 1, a2 = ns2 > 0 ? ns2 : 1;
  const std::string A = "const ExbPatArgs& pa, const long long kg";
  o << "struct P" << index << " {\n";
  o << "  static constexpr int INDEX = " << index << ", KIND = " << p.ir.kind << ", NS1 = " << ns1 << ", NS2 = " << ns2 << ";\n";
  o << "  static constexpr bool WIN = " << (windowed ? "true" : "false") << "; static constexpr long long XLO = " << (windowed ? p.xlo : 0)
    << ", XHI = " << (windowed ? p.xhi : 0) << ";\n";
  {  // row (offset0, nlp.jl:1980-2001; idxx :2012-2015)
    Body B; Gen g(p, B, 0);
    std::vector<std::string> tail;
    if (p.ir.kind == KIND_CON) tail.push_back("return pa.o0 + kg + 1;");
    else if (p.ir.kind == KIND_OBJ) tail.push_back("return 0;");
    else if (p.ir.idx_roots.size() == 1) {
      NV& r = g.real(p.ir.idx_roots[0]);
      tail.push_back("return pa.o0 + " + r.rs + ";");
    } else {
      std::string lin = "0LL", a = "1LL";
      for (size_t d = 0; d < p.ir.idx_roots.size(); d++) {
        NV& r = g.real(p.ir.idx_roots[d]);
        lin = B.tmp("long long", lin + " + " + a + " * (" + r.rs + " - 1)");
        a = B.tmp("long long", a + " * pa.dim[" + std::to_string(d) + "]");
      }
      tail.push_back("return pa.o0 + " + lin + " + 1;");
    }
    emit_fn(o, "long long row(" + A + ")", B, tail);
  }
  {  // val
    Body B; Gen g(p, B, 0);
    NV& r = g.fwd(p.ir.root);
    emit_eval_fn(o, "val", A + ", const XA x, const double* __restrict__ th", "pa, kg, x, th", 0, B, {"return " + r.x.s + ";"});
    p.ppt0 = ppt_for(body_weight(B), 1);
  }
  {  // d1
    Body B; Gen g(p, B, 1);
    std::vector<std::string> tail;
    if (ns1 > 0) {
      g.fwd(p.ir.root);
      g.comp = &p.comp1; g.slot.assign((size_t)ns1, K(0));
      g.rpass1(p.ir.root, K(1));
      for (int j = 0; j < ns1; j++) tail.push_back("s[" + std::to_string(j) + "] = " + g.slot[(size_t)j].s + ";");
    }
    emit_eval_fn(o, "d1", A + ", const XA x, const double* __restrict__ th", "pa, kg, x, th", a1, B, tail);
    p.ppt1 = ppt_for(body_weight(B), a1);
    // Owner-computes gradient: objective over a range iterator whose slots all address x[t + const].  Variable v
    // then receives exactly slot j of point t = v - shift1[j] (if that point exists), so ONE thread per variable
    // re-evaluates the (cheap) body once per slot and writes g[v] once: no gradbuffer, no sorted list, no atomics
    // (the reference writes nnzg slots and then segment-sums them, ext:310-336,691-697).  Bodies too heavy to be
    // re-evaluated o1step times keep the slot + compress path.
    p.gather1 = false; p.shift1.clear();
    static const int gmax = getenv("EXB_TUNE_GATHER_W") ? atoi(getenv("EXB_TUNE_GATHER_W")) : 300;
    // ... and only when a point costs nothing to fetch: a pattern that reads iterator data (beyond iota columns) would re-load it
    // once per slot, so it goes to the tile kernel instead (each point evaluated once; see `tgrad` below)
    bool reads_data = false;
    for (auto& nd : p.ir.nodes) if (nd.tag == T_DATA_FIELD && !((size_t)nd.a < p.ir.iota.size() && p.ir.iota[(size_t)nd.a])) reads_data = true;
    if (p.ir.kind == KIND_OBJ && shiftable(p.ir) && ns1 > 0 && !reads_data && body_weight(B) * ns1 <= gmax) {
      bool ok = true;
      for (int j = 0; j < ns1 && ok; j++) {
        i64 cf, ct;
        ok = affine_index(p.ir, (int)p.ir.nodes[(size_t)p.leaf1[(size_t)j]].a, cf, ct) && cf == 1;
        p.shift1.push_back(ct);
      }
      p.gather1 = ok;
    }
    {  // tile gradient: every first-order slot of an objective pattern addresses x[t + const] (any body weight)
      p.tgrad = false; p.egrad = false;
      std::vector<i64> sh;
      if (p.ir.kind == KIND_OBJ && shiftable(p.ir) && ns1 > 0) {
        bool ok = true;
        for (int j = 0; j < ns1 && ok; j++) {
          i64 cf, ct;
          ok = affine_index(p.ir, (int)p.ir.nodes[(size_t)p.leaf1[(size_t)j]].a, cf, ct) && cf == 1;
          sh.push_back(ct);
        }
        if (ok) {
          p.g_cbmin = *std::min_element(sh.begin(), sh.end()); p.g_cbmax = *std::max_element(sh.begin(), sh.end());
          // light bodies without data stay with the per-variable kernel (LV: 0.033-0.044 ms against 0.049 for the tile form:
          // its two barriers and the staging cost more than re-evaluating a 10-flop body twice); everything else is tiled
          p.tgrad = !p.gather1 && p.g_cbmax - p.g_cbmin <= 64 && getenv("EXB_NO_TGRAD") == nullptr;
          // the model's ONLY objective pattern with gradient slots: the fused evaluation kernels (exb_eval) have its first-order
          // slots in registers anyway -- they stage them and write g themselves (exb_eval_block), no gradient launch at all
          p.egrad = sole_obj && p.g_cbmax - p.g_cbmin <= 64 && getenv("EXB_NO_EGRAD") == nullptr;
          if (p.tgrad || p.egrad) p.shift1 = sh;
        }
      }
      if (p.tgrad || p.egrad) {
        const int ts1 = ns1 | 1;
        std::vector<int> ord((size_t)ns1);
        for (int j = 0; j < ns1; j++) ord[(size_t)j] = j;
        std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return sh[(size_t)a] > sh[(size_t)b]; });
        o << "  static constexpr bool TGRAD = " << (p.tgrad ? "true" : "false") << ", EGRAD = " << (p.egrad ? "true" : "false") << "; static constexpr int TS1 = " << ts1
          << "; static constexpr long long GBMIN = " << p.g_cbmin << ", GBMAX = " << p.g_cbmax << ";\n";
        o << "  template <bool CHECK> __device__ static __forceinline__ void ggather(const double* __restrict__ rp, const int ql, const int qlo, const int qhi, double (&acc)[1]) {\n";
        for (int j : ord) {
          const long long off = (long long)j - (long long)sh[(size_t)j] * ts1;
          o << "    if (!CHECK || (ql - (" << sh[(size_t)j] << ") >= qlo && ql - (" << sh[(size_t)j] << ") < qhi)) acc[0] += rp[" << off << "];\n";
        }
        o << "  }\n";
      } else {
        o << "  static constexpr bool TGRAD = false, EGRAD = false; static constexpr int TS1 = 1; static constexpr long long GBMIN = 0, GBMAX = 0;\n";
        o << "  template <bool CHECK> __device__ static __forceinline__ void ggather(const double* __restrict__, const int, const int, const int, double (&)[1]) {}\n";
      }
    }
    if (p.gather1) {
      // summation order of the reference for one variable: ascending global slot number = ascending point, then slot
      std::vector<int> ord((size_t)ns1);
      for (int j = 0; j < ns1; j++) ord[(size_t)j] = j;
      std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return p.shift1[(size_t)a] > p.shift1[(size_t)b]; });
      o << "  __device__ static __forceinline__ double g1(const ExbPatArgs& pa, const long long v, const ExbXG x, const double* __restrict__ th) {\n";
      // branch-free: an out-of-range point is replaced by the pattern's first local point and its slot discarded, so the
      // loads of every slot are issued up front (the kernel is latency-bound: ~10 flops per 8-byte word)
      // the kernel partitions VARIABLES (a sharded handle owns a contiguous range of them), so every point of the pattern
      // is eligible: [0, nfull), not the handle's point shard
      o << "    double acc = 0.0;\n    if (pa.nfull > 0) {\n";
      for (int j : ord) {
        o << "      { const long long kg = v - (" << Gen::ilit(p.shift1[(size_t)j]) << ") - pa.start; const bool in = kg >= 0 && kg < pa.nfull;\n"
          << "        double s[" << a1 << "]; d1(pa, in ? kg : 0, x, th, s); acc += in ? s[" << j << "] : 0.0; }\n";
      }
      o << "    }\n";
      o << "    return acc;\n  }\n";
    }
  }
  {  // d2
    Body B; Gen g(p, B, 2);
    std::vector<std::string> tail;
    if (ns2 > 0) {
      g.fwd(p.ir.root);
      g.comp = &p.comp2; g.slot.assign((size_t)ns2, K(0));
      g.hrpass0(p.ir.root, Sym("a0"), K(0));
      for (int j = 0; j < ns2; j++) tail.push_back("s[" + std::to_string(j) + "] = " + g.slot[(size_t)j].s + ";");
    }
    emit_eval_fn(o, "d2", A + ", const XA x, const double* __restrict__ th, const double a0", "pa, kg, x, th, a0", a2, B, tail);
    p.ppt2 = ppt_for(body_weight(B), a2);
  }
  {  // d012 (exb_eval: one sweep for cons! / obj + jac_coord! / grad! slots + hess_coord! slots; src/nlp.jl:1827-1940 evaluates the
     // same tree three to five times at the same x)
    Body B; Gen g(p, B, 2);
    NV& r = g.fwd(p.ir.root);
    std::vector<std::string> tail;
    tail.push_back("v = " + r.x.s + ";");
    if (ns1 > 0) {
      g.comp = &p.comp1; g.slot.assign((size_t)ns1, K(0)); g.cnt = 0;
      g.rpass1(p.ir.root, K(1));
      for (int j = 0; j < ns1; j++) tail.push_back("s1[" + std::to_string(j) + "] = " + g.slot[(size_t)j].s + ";");
    }
    if (ns2 > 0) {
      g.comp = &p.comp2; g.slot.assign((size_t)ns2, K(0)); g.cnt = 0;
      g.hrpass0(p.ir.root, Sym("a0"), K(0));
      for (int j = 0; j < ns2; j++) tail.push_back("s2[" + std::to_string(j) + "] = " + g.slot[(size_t)j].s + ";");
    }
    emit_eval3_fn(o, A + ", const XA x, const double* __restrict__ th, const double a0", "pa, kg, x, th, a0", a1, a2, B, tail);
    p.ppte = ppt_for(body_weight(B), a1 + a2);
    o << "  static constexpr int PPTE = " << p.ppte << "; static constexpr bool G1 = " << ((p.gather1 || p.tgrad) ? "true" : "false") << ";\n";
  }
  {  // d01: value + first-order slots from one order-1 sweep (exb_eval with mask obj | grad | cons | jac: the first-order evaluation
     // a solver does at a new iterate); same signature as d012, s2 / a0 unused
    Body B; Gen g(p, B, 1);
    NV& r = g.fwd(p.ir.root);
    std::vector<std::string> tail;
    tail.push_back("v = " + r.x.s + ";");
    if (ns1 > 0) {
      g.comp = &p.comp1; g.slot.assign((size_t)ns1, K(0)); g.cnt = 0;
      g.rpass1(p.ir.root, K(1));
      for (int j = 0; j < ns1; j++) tail.push_back("s1[" + std::to_string(j) + "] = " + g.slot[(size_t)j].s + ";");
    }
    emit_eval3_fn(o, A + ", const XA x, const double* __restrict__ th, const double a0", "pa, kg, x, th, a0", a1, a2, B, tail, "d01");
    p.ppte1 = ppt_for(body_weight(B), a1 + 1);
    o << "  static constexpr int PPTE1 = " << p.ppte1 << ";\n";
  }
  if (hd != nullptr && ns2 > 0 && p.tile_ok) {
    // Column-tile form (exb_tile_body): the block stages the second-order slots of the points around its tile of columns in
    // shared memory (`raw`, point-major, TSTRIDE words per point); the thread that owns column c then sums every slot that
    // lands in column c -- slot j of the point whose range value is c - CB_j -- into acc[rank of the entry's row - column
    // distance].  The additions run in ascending point, then slot order: the order in which `_compress!` (utils.jl:564-571)
    // meets the duplicates of one coordinate in the sorted list.
    std::vector<int> ord((size_t)ns2);
    for (int j = 0; j < ns2; j++) ord[(size_t)j] = j;
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return p.t2_cb[(size_t)a] > p.t2_cb[(size_t)b]; });
    o << "  static constexpr bool TILE = true; static constexpr int TSTRIDE = " << tile_stride << "; static constexpr long long CBMIN = " << p.t_cbmin
      << ", CBMAX = " << p.t_cbmax << ";\n";
    // rp = the staged slots of the point whose range value equals the column (local point ql): slot j of the point CB_j
    // earlier sits at the compile-time offset j - CB_j * TSTRIDE from it.  CHECK = false: interior tile, every point exists.
    o << "  template <bool CHECK> __device__ static __forceinline__ void hgather(const double* __restrict__ rp, const int ql, const int qlo, const int qhi, double (&acc)["
      << hd->size() << "]) {\n";
    for (int j : ord) {
      const int r = p.t2_r[(size_t)j];
      const long long off = (long long)j - (long long)p.t2_cb[(size_t)j] * tile_stride;
      o << "    if (!CHECK || (ql - (" << p.t2_cb[(size_t)j] << ") >= qlo && ql - (" << p.t2_cb[(size_t)j] << ") < qhi)) acc[" << r << "] += rp[" << off << "];\n";
    }
    o << "  }\n";
  } else {
    o << "  static constexpr bool TILE = false; static constexpr int TSTRIDE = 1; static constexpr long long CBMIN = 0, CBMAX = 0;\n";
    if (hd != nullptr)
      o << "  template <bool CHECK> __device__ static __forceinline__ void hgather(const double* __restrict__, const int, const int, const int, double (&)[" << hd->size() << "]) {}\n";
  }
  {  // s1: variable index per first-order slot (jacobian.jl:69-83)
    Body B; Gen g(p, B, 0);
    std::vector<std::string> tail;
    for (int j = 0; j < ns1; j++) {
      NV& ix = g.real((int)p.ir.nodes[(size_t)p.leaf1[(size_t)j]].a);
      tail.push_back("col[" + std::to_string(j) + "] = " + ix.rs + ";");
    }
    emit_fn(o, "void s1(" + A + ", long long (&col)[" + std::to_string(a1) + "])", B, tail);
  }
  {  // s2: (max, min) per second-order slot (hessian.jl:593-607,622-642)
    Body B; Gen g(p, B, 0);
    std::vector<std::string> tail;
    for (int j = 0; j < ns2; j++) {
      NV& ia = g.real((int)p.ir.nodes[(size_t)p.leaf2[(size_t)j].first].a);
      NV& ib = g.real((int)p.ir.nodes[(size_t)p.leaf2[(size_t)j].second].a);
      tail.push_back("r[" + std::to_string(j) + "] = exb_imax(" + ia.rs + ", " + ib.rs + "); c[" + std::to_string(j) + "] = exb_imin(" + ia.rs + ", " + ib.rs + ");");
    }
    emit_fn(o, "void s2(" + A + ", long long (&r)[" + std::to_string(a2) + "], long long (&c)[" + std::to_string(a2) + "])", B, tail);
  }
  {  // hp: the point's contribution to a Hessian-vector product (kersyspmv / kersyspmv2, ext:497-511: y[r] += h v[c], and the mirror
     // off the diagonal), grouped by DISTINCT index expression: one v load and one scattered add per distinct variable of the
     // point instead of two per slot (LV constraint: 3 instead of 9; rocket dynamics: 9 instead of ~110).  Expressions that
     // cannot be proved different are compared at run time (a self loop f_bus == t_bus is a diagonal entry: no mirror).
    Body B; Gen g(p, B, 0);
    std::vector<int> ex;                                 // distinct index expressions (node ids)
    std::vector<int> ta((size_t)ns2), tb((size_t)ns2);
    auto slot_of = [&](int e) {
      for (size_t t = 0; t < ex.size(); t++) if (index_relation(p.ir, ex[t], e) == 1) return (int)t;
      ex.push_back(e);
      return (int)ex.size() - 1;
    };
    for (int j = 0; j < ns2; j++) {
      ta[(size_t)j] = slot_of((int)p.ir.nodes[(size_t)p.leaf2[(size_t)j].first].a);
      tb[(size_t)j] = slot_of((int)p.ir.nodes[(size_t)p.leaf2[(size_t)j].second].a);
    }
    const int nt = (int)ex.size(), at = nt > 0 ? nt : 1;
    std::vector<std::string> tail;
    for (int t = 0; t < nt; t++) {
      NV& ix = g.real(ex[(size_t)t]);
      tail.push_back("idx[" + std::to_string(t) + "] = " + ix.rs + "; const double vv" + std::to_string(t) + " = __ldg(v + (idx[" + std::to_string(t) + "] - 1)); val[" + std::to_string(t) + "] = 0.0;");
    }
    for (int j = 0; j < ns2; j++) {
      const std::string a = std::to_string(ta[(size_t)j]), b = std::to_string(tb[(size_t)j]), sj = "s[" + std::to_string(j) + "]";
      if (ta[(size_t)j] == tb[(size_t)j]) { tail.push_back("val[" + a + "] += " + sj + " * vv" + a + ";"); continue; }
      tail.push_back("val[" + a + "] += " + sj + " * vv" + b + ";");
      const int rel = index_relation(p.ir, ex[(size_t)ta[(size_t)j]], ex[(size_t)tb[(size_t)j]]);
      tail.push_back(std::string(rel == 0 ? "" : "if (idx[" + a + "] != idx[" + b + "]) ") + "val[" + b + "] += " + sj + " * vv" + a + ";");
    }
    o << "  static constexpr int NT2 = " << nt << ";\n";
    emit_fn(o, "void hp(" + A + ", const double (&s)[" + std::to_string(a2) + "], const double* __restrict__ v, long long (&idx)[" + std::to_string(at) +
                   "], double (&val)[" + std::to_string(at) + "])", B, tail);
  }
  o << "  static constexpr int PPT0 = " << p.ppt0 << ", PPT1 = " << p.ppt1 << ", PPT2 = " << p.ppt2 << ";\n";
  o << "  static constexpr int W0 = 0, W1 = " << ns2 << ";   // window of second-order slots an exb_hess_g0 entry keeps (ExbSplit narrows it)\n";
  o << "};\n";
  return o.str();
}

inline std::string plist(const std::vector<int>& v) {
  std::string s;
  for (size_t k = 0; k < v.size(); k++) { if (k) s += ", "; s += "P" + std::to_string(v[k]); }
  return s;
}

inline bool build_plan(Plan& pl, const void* ir, size_t bytes, const void* const* host_data = nullptr, int n_data = 0) {
  if (!parse_ir(ir, bytes, pl.m, pl.error)) return false;
  if (getenv("EXB_NO_IOTA") == nullptr) detect_iota(pl.m, host_data, n_data);
  pl.pats.resize(pl.m.pats.size());
  for (size_t k = 0; k < pl.pats.size(); k++) {
    PatternPlan& p = pl.pats[k];
    p.ir = pl.m.pats[k];
    if ((int)p.ir.fields.size() > EXB_MAXF_HOST) { pl.error = "a pattern reads more than 16 distinct iterator fields"; return false; }
    if ((int)p.ir.idx_roots.size() > EXB_MAXD_HOST) { pl.error = "augmentation index has more than 4 dimensions"; return false; }
    classify(p);
    for (size_t q = 0; q < p.ir.nodes.size(); q++) {
      const IRNode& n = p.ir.nodes[q];
      if ((n.tag == T_VAR || n.tag == T_PAR) && !(p.info[(size_t)n.a].kind == K_REAL && p.info[(size_t)n.a].is_int)) {
        pl.error = "variable / parameter index expression is not an integer expression"; return false;
      }
    }
    for (int r : p.ir.idx_roots)
      if (!(p.info[(size_t)r].kind == K_REAL && p.info[(size_t)r].is_int)) { pl.error = "augmentation row index is not an integer expression"; return false; }
    probe(p);
    // comps supplied by the emitter (e.g. the Julia shim passing SIMDFunction.comp1/comp2) must agree
    if (!p.ir.comp1_given.empty()) {
      bool same = p.ir.comp1_given.size() == p.comp1.size();
      for (size_t q = 0; same && q < p.comp1.size(); q++) same = p.ir.comp1_given[q] == p.comp1[q];
      if (!same) { pl.error = "comp1 supplied in the IR disagrees with the probe"; return false; }
    }
    if (!p.ir.comp2_given.empty()) {
      bool same = p.ir.comp2_given.size() == p.comp2.size();
      for (size_t q = 0; same && q < p.comp2.size(); q++) same = p.ir.comp2_given[q] == p.comp2[q];
      if (!same) { pl.error = "comp2 supplied in the IR disagrees with the probe"; return false; }
    }
  }
  // running counters in add order (nlp.jl:1474-1482, 1597-1611, 1730-1738)
  for (auto& p : pl.pats) {
    const i64 n = p.ir.nitr;
    if (p.ir.kind == KIND_OBJ) {
      p.o0 = pl.nobj; p.ob = pl.nobj; p.o1 = pl.nnzg; p.o2 = pl.nnzh;          // nlp.jl:1450
      pl.nobj += n; pl.nnzg += n * p.o1step; pl.nnzh += n * p.o2step;
    } else if (p.ir.kind == KIND_CON) {
      p.o0 = pl.ncon; p.o1 = pl.nnzj; p.o2 = pl.nnzh;                          // nlp.jl:1587
      pl.ncon += n; pl.nnzj += n * p.o1step; pl.nnzh += n * p.o2step;
    } else {
      p.o0 = pl.pats[(size_t)p.ir.base].o0; p.oa = pl.nconaug; p.o1 = pl.nnzj; p.o2 = pl.nnzh;   // nlp.jl:1683
      pl.nconaug += n; pl.nnzj += n * p.o1step; pl.nnzh += n * p.o2step;
    }
    if (p.ir.o0 >= 0 && p.ir.kind != KIND_OBJ && p.ir.o0 != p.o0) { pl.error = "o0 supplied in the IR disagrees with the counters"; return false; }
    if (p.ir.o1 >= 0 && p.ir.o1 != p.o1) { pl.error = "o1 supplied in the IR disagrees with the counters"; return false; }
    if (p.ir.o2 >= 0 && p.ir.o2 != p.o2) { pl.error = "o2 supplied in the IR disagrees with the counters"; return false; }
  }
  // module source
  {
    i64 mx = std::max(std::max(pl.m.nvar, pl.m.npar), std::max(std::max(pl.nnzh, pl.nnzj), std::max(pl.nnzg, pl.ncon)));
    for (auto& p : pl.pats) mx = std::max(mx, p.ir.nitr + (p.ir.range_start > 0 ? p.ir.range_start : 0));
    // measured on LV N=1e7 (scripts/ab.py): 32-bit address chains are NOT faster (0.166 vs 0.162 ms), so this stays opt-in
    pl.idx32 = mx < 2000000000LL && getenv("EXB_TUNE_IDX32") != nullptr;
  }
  if (const char* e = getenv("EXB_TUNE_BLOCK")) { int b = atoi(e); if (b == 64 || b == 128 || b == 256 || b == 512) pl.block = b; }
  if (const char* e = getenv("EXB_TUNE_MINB")) { int b = atoi(e); if (b >= 1 && b <= 32) pl.minb = b; }
  std::ostringstream o;
  o << "// generated by exb_plan.hpp -- one struct per pattern, kernels per callback\n";
  // The persistent Hessian kernel needs an x window for EVERY pattern with second-order slots.  Opt-in (EXB_TUNE_PERSISTENT=1):
  // measured on LV N=1e7 it is SLOWER than the classic one-tile-per-block kernel (0.174 ms with the classic tiles and a single
  // staging buffer, 0.221 ms with one point per thread and double-buffered staging, against 0.158 ms): the hardware block
  // scheduler over 16 small resident blocks per SM hides the x / y latency better than a software pipeline that pays two
  // barriers per tile and loses occupancy to its shared-memory windows.  Kept as a tuner candidate for experiments.
  pl.hess_windowed = getenv("EXB_TUNE_PERSISTENT") != nullptr && atoi(getenv("EXB_TUNE_PERSISTENT")) != 0;
  {
    bool any = false;
    for (auto& p : pl.pats) { compute_window(p); compute_xrange(p); if (p.o2step > 0) { any = true; pl.hess_windowed = pl.hess_windowed && p.win; } }
    pl.hess_windowed = pl.hess_windowed && any;
  }
  {  // duplicate-free Hessian through the column-tile kernel: every pattern with second-order slots must be shift-indexed, and
     // every row - column distance must exist in ONE interval of columns (else: the sorted gather of exb_runtime.cpp)
    bool any = false; pl.tile_ok = true;
    for (auto& p : pl.pats) { compute_tile(p); if (p.o2step > 0) { any = true; pl.tile_ok = pl.tile_ok && p.tile_ok; } }
    pl.tile_ok = pl.tile_ok && any && getenv("EXB_NO_TILE") == nullptr;
    // columns per thread: few patterns (LV) -- larger tiles amortise the per-tile prologue (measured 2 -> 0.147, 3 -> 0.142, 4 ->
    // 0.150 ms); many patterns reading iterator data (32-pattern family) -- small tiles keep more blocks resident to hide the
    // load latency of each pattern's phase (measured 1 -> 0.291, 2 -> 0.552, 3 -> 0.710 ms)
    { int nh = 0; for (auto& p : pl.pats) if (p.o2step > 0) nh++; pl.tile_ppt = nh > 4 ? 1 : 3; }
    if (const char* e = getenv("EXB_TUNE_TILE_PPT")) { int v = atoi(e); if (v >= 1 && v <= 8) pl.tile_ppt = v; }
    if (pl.tile_ok) {
      for (auto& p : pl.pats) for (i64 d : p.t2_d) pl.hd.push_back(d);
      std::sort(pl.hd.begin(), pl.hd.end());
      pl.hd.erase(std::unique(pl.hd.begin(), pl.hd.end()), pl.hd.end());
      if (pl.hd.size() > 16) pl.tile_ok = false;
    }
    if (pl.tile_ok) {
      const size_t D = pl.hd.size();
      std::vector<std::vector<std::pair<i64, i64>>> iv(D);
      for (auto& p : pl.pats) {
        p.t2_r.clear();
        for (size_t j = 0; j < p.t2_d.size(); j++) {
          const int r = (int)(std::lower_bound(pl.hd.begin(), pl.hd.end(), p.t2_d[j]) - pl.hd.begin());
          p.t2_r.push_back(r);
          if (p.ir.nitr > 0) iv[(size_t)r].push_back({p.ir.range_start + p.t2_cb[j], p.ir.range_start + p.ir.nitr - 1 + p.t2_cb[j]});
        }
        if (p.o2step > 0) pl.tile_halo = std::max(pl.tile_halo, (int)(p.t_cbmax - p.t_cbmin));
      }
      pl.h_lo.assign(D, 1); pl.h_len.assign(D, 0);
      for (size_t r = 0; r < D && pl.tile_ok; r++) {
        auto& v = iv[r];
        if (v.empty()) continue;
        std::sort(v.begin(), v.end());
        i64 lo = v[0].first, hi = v[0].second;
        for (size_t q = 1; q < v.size(); q++) { if (v[q].first > hi + 1) { pl.tile_ok = false; break; } hi = std::max(hi, v[q].second); }
        if (lo < 1 || hi > pl.m.nvar) pl.tile_ok = false;
        pl.h_lo[r] = lo; pl.h_len[r] = hi - lo + 1;
      }
    }
    if (!pl.tile_ok) { pl.hd.clear(); pl.h_lo.clear(); pl.h_len.clear(); }
  }
  int nobj1 = 0;   // objective patterns with gradient slots
  for (auto& p : pl.pats) if (p.ir.kind == KIND_OBJ && p.o1step > 0) nobj1++;
  for (size_t k = 0; k < pl.pats.size(); k++) {
    int stride = pl.pats[k].o2step | 1;   // odd word stride: conflict-free 64-bit shared-memory accesses
    o << gen_pattern(pl.pats[k], (int)k, pl.hess_windowed && pl.pats[k].win, pl.tile_ok ? &pl.hd : nullptr, stride, nobj1 == 1);
    if (pl.pats[k].egrad) pl.egrad_pat = (int)k;
  }
  // kernel pattern lists
  for (size_t k = 0; k < pl.pats.size(); k++) {
    const PatternPlan& p = pl.pats[k];
    if (p.o2step > 0) {
      pl.k_hess.push_back((int)k);
      // opt-in (EXB_TUNE_SPLIT_NS = slots per point above which a pattern is split): MEASURED SLOWER on the rocket nh = 1e6 --
      // hess_coord! 0.117 ms split (90 registers, 5 blocks / SM) against 0.101 ms whole (108 registers, 4 blocks / SM); 0.110 /
      // 0.115 ms with 85- / 64-register builds: the second forward sweep and the row-wise stores cost more than the occupancy buys
      const int split_ns = getenv("EXB_TUNE_SPLIT_NS") ? atoi(getenv("EXB_TUNE_SPLIT_NS")) : 0;
      if (split_ns > 0 && p.o2step > split_ns) {
        const int h = (p.o2step + 1) / 2;
        pl.k_hess_l.push_back((int)k); pl.k_hess_w.push_back({0, h});
        pl.k_hess_l.push_back((int)k); pl.k_hess_w.push_back({h, p.o2step});
      } else { pl.k_hess_l.push_back((int)k); pl.k_hess_w.push_back({0, p.o2step}); }
    }
    if (p.ir.kind == KIND_OBJ) { pl.k_obj.push_back((int)k); if (p.o1step > 0) (p.tgrad ? pl.k_tgrad : p.gather1 ? pl.k_ggrad : pl.k_sgrad).push_back((int)k); }
    else { pl.k_cons.push_back((int)k); if (p.o1step > 0) pl.k_jac.push_back((int)k); }
    if (p.tgrad) pl.tgrad_halo = std::max(pl.tgrad_halo, (int)(p.g_cbmax - p.g_cbmin));
    if (p.ir.kind == KIND_AUG) pl.k_aug.push_back((int)k);
    pl.k_eval.push_back((int)k);   // every pattern has a value
  }
  auto kern = [&](const char* name, const char* body, const std::vector<int>& v, const char* targ) {
    if (v.empty()) return;
    o << "extern \"C\" __global__ void __launch_bounds__(EXB_BLOCK, EXB_MINB) " << name << "(const ExbGroup g, const ExbCall c) { "
      << body << "<" << targ << plist(v) << ">(g, c); }\n";
  };
  if (pl.k_hess_l.size() != pl.k_hess.size()) pl.hess_windowed = false;   // the (opt-in) persistent form knows no split entries
  if (!pl.k_hess_l.empty()) {
    std::string lst;
    for (size_t q = 0; q < pl.k_hess_l.size(); q++) {
      const int k = pl.k_hess_l[q];
      const bool whole = pl.k_hess_w[q].first == 0 && pl.k_hess_w[q].second == pl.pats[(size_t)k].o2step;
      if (q) lst += ", ";
      lst += whole ? "P" + std::to_string(k)
                   : "ExbSplit<P" + std::to_string(k) + ", " + std::to_string(pl.k_hess_w[q].first) + ", " + std::to_string(pl.k_hess_w[q].second) + ">";
    }
    o << "extern \"C\" __global__ void __launch_bounds__(EXB_BLOCK, EXB_MINB) exb_hess_g0(const ExbGroup g, const ExbCall c) { exb_hess_body<" << lst << ">(g, c); }\n";
  }
  if (pl.tile_ok) {
    o << "extern \"C\" __global__ void __launch_bounds__(EXB_BLOCK, EXB_MINB) exb_hessc_g0(const ExbGroup g, const ExbCall c, const ExbTile t) { "
      << "exb_tile_body<2, " << pl.hd.size() << ", " << pl.tile_ppt << ", " << plist(pl.k_hess) << ">(g, c, t); }\n";
  }
  if (pl.hess_windowed) kern("exb_hessp_g0", "exb_hessp_body", pl.k_hess, "");
  kern("exb_eval_g0", "exb_eval_body", pl.k_eval, "2, ");
  kern("exb_eval1_g0", "exb_eval_body", pl.k_eval, "1, ");
  kern("exb_eval0_g0", "exb_eval_body", pl.k_eval, "0, ");
  kern("exb_jac_g0", "exb_d1_body", pl.k_jac, "");
  kern("exb_sgrad_g0", "exb_d1_body", pl.k_sgrad, "");
  kern("exb_ggrad_g0", "exb_ggrad_body", pl.k_ggrad, "");
  if (!pl.k_tgrad.empty()) {
    pl.tgrad_ppt = pl.k_tgrad.size() > 4 ? 1 : 2;   // same reasoning as tile_ppt
    if (const char* e = getenv("EXB_TUNE_TGRAD_PPT")) { int v = atoi(e); if (v >= 1 && v <= 8) pl.tgrad_ppt = v; }
    o << "extern \"C\" __global__ void __launch_bounds__(EXB_BLOCK, EXB_MINB) exb_gradt_g0(const ExbGroup g, const ExbCall c, const ExbTile t) { "
      << "exb_tile_body<1, 1, " << pl.tgrad_ppt << ", " << plist(pl.k_tgrad) << ">(g, c, t); }\n";
  }
  kern("exb_cons_g0", "exb_cons_body", pl.k_cons, "");
  kern("exb_obj_g0", "exb_obj_body", pl.k_obj, "");
  kern("exb_jstruct64_g0", "exb_jstruct_body", pl.k_jac, "long long, ");
  kern("exb_jstruct32_g0", "exb_jstruct_body", pl.k_jac, "int, ");
  kern("exb_gstruct64_g0", "exb_jstruct_body", pl.k_sgrad, "long long, ");
  kern("exb_hstruct64_g0", "exb_hstruct_body", pl.k_hess, "long long, ");
  kern("exb_hstruct32_g0", "exb_hstruct_body", pl.k_hess, "int, ");
  kern("exb_augrow_g0", "exb_augrow_body", pl.k_aug, "");
  kern("exb_jprod_g0", "exb_jprod_body", pl.k_jac, "");
  kern("exb_jtprod_g0", "exb_jtprod_body", pl.k_jac, "");
  kern("exb_hprod_g0", "exb_hprod_body", pl.k_hess, "");
  pl.source = o.str();
  return true;
}

}  // namespace exb
