// =============================================================================
// oracle/exa_oracle.cpp — TEST INFRASTRUCTURE ONLY.
//
// A CPU restatement of the reference's (exanauts/ExaModels.jl v0.12.0) per-pattern
// evaluation path, interpreting the same pattern IR the product consumes.  Only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this; the product path (examodels.jl_b200/csrc) never links,
// calls or falls back to it.
//
// It is NOT the reference: the reference is pure Julia and there is no Julia in
// this image.  Parity of this restatement is pinned by the
// reference's own fixtures where they exist (tests/test_oracle_pins.py):
//   * per-pattern slot counts / compressor maps hand-derived from the passes
//      and the raw traversal counts 10 / 21 pinned by
//     test/JuMPTest/JuMPTest.jl:404-405,
//   * closed-form cons values of test/NLPTest/conaug_test.jl:86-213,
//   * every nnzj / nnzh hard-coded in test/JuMPTest/JuMPTest.jl (per-term slot counts),
//   * derivative tables against finite differences / sympy as in
//     test/ADTest/ADTest.jl:298-374,
//   * the Ipopt runs the reference's documentation build printed for the parametric LV N=10
//     model (docs/src/parameters.md): replaying them with these callbacks reproduces every
//     printed digit of every iteration (tests/golden/ipopt_logs.json) -- pins all five value
//     callbacks, both structures and the parameter updates against reference output,
//   * the Ipopt solution and
//     multipliers of LV N=10 printed in docs/src/develop.md:84-105, which must be a KKT
//     point of this restatement (cons = 0, grad f + J' lambda = 0 to the precision of the
//     solve: pins cons, grad!, jac_coord!, jac_structure! against reference output),
//   * the SpecialFunctions-extension values against mpmath (the reference's come from
//     SpecialFunctions.jl, not vendored).
// No stored numeric derivative VECTORS exist in the reference tree: second-order values
// are pinned by derivation, exact symbolic differentiation and finite differences.
//
// Every function cites the reference file:line it restates (paths relative to
// /root/reference/).  The recursion is kept literal (one C++ function per Julia
// method family); the only liberty is that the evaluated "adjoint tree" lives in
// per-thread arrays indexed by a pre-expanded tree instead of nested structs.
// =============================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <thread>
#include "ora_tables.hpp"

using namespace ora;
namespace {

// ---- build: IR parse, tree expansion, static kinds ---------------------------------------
int expand(Pattern& p, int ir) {
  const IRNode& n = p.ir[ir];
  TNode t; std::memset(&t, 0, sizeof t);
  t.tag = (int)n.tag; t.ir = ir; t.c1 = t.c2 = -1; t.ipay = n.payload; std::memcpy(&t.fpay, &n.payload, 8);
  t.op = (int)n.payload;
  if (n.tag == T_VAR || n.tag == T_PAR || n.tag == T_OP1) t.c1 = expand(p, (int)n.a);
  if (n.tag == T_OP2) { t.c1 = expand(p, (int)n.a); t.c2 = expand(p, (int)n.b); }
  if (n.tag == T_DATA_FIELD) t.op = (int)n.a;
  // static kind of the evaluated node under an (Second)AdjointNodeSource:
  switch (n.tag) {
    case T_VAR: t.kind = K_VAR; break;                           // graph.jl:397-400,491-494
    case T_NULL: t.kind = K_NULL; break;                         // graph.jl:499-502
    case T_OP1: t.kind = p.t[t.c1].kind == K_REAL ? K_REAL : K_N1; break;   // register.jl:65-71
    case T_OP2: {
      bool r1 = p.t[t.c1].kind == K_REAL, r2 = p.t[t.c2].kind == K_REAL;
      if (r1 && r2) t.kind = K_REAL;
      else if (r2) { t.kind = K_N1; t.fx = FX_SECOND; }          // register.jl:231-248
      else if (r1) { t.kind = K_N1; t.fx = FX_FIRST; }           // register.jl:249-266
      else t.kind = K_N2;                                        // register.jl:209-230
    } break;
    default: t.kind = K_REAL;                                    // constants, data, θ[...]
  }
  p.t.push_back(t);
  return (int)p.t.size() - 1;
}

bool ir_equal(const Pattern& p, int a, int b) {   // Julia `===` on immutable node structs
  if (a == b) return true;
  const IRNode &x = p.ir[a], &y = p.ir[b];
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

// ---- per-thread evaluation workspace ---------------------------------------------------
struct Work {
  std::vector<double> x, y1, y2, h11, h12, h22;   // per tree node (Node1: y=y1, h=h11)
  std::vector<Real> re;                           // Real values (kind K_REAL and index exprs)
  std::vector<i64> vi;                            // VAR: evaluated index
  void size(size_t n) { x.resize(n); y1.resize(n); y2.resize(n); h11.resize(n); h12.resize(n); h22.resize(n); re.resize(n); vi.resize(n); }
};

struct Ctx {
  const Pattern* p; const double* X; const double* TH; Work* w;
  i64 k;              // 0-based point number
  const unsigned char* elem;
};

Real data_value(const Ctx& c, const TNode& t) {
  Real r; r.is_int = false; r.i = 0; r.f = 0;
  if (t.tag == T_DATA_SELF) { r.is_int = true; r.i = c.p->range_start + c.k; return r; }
  const Field& f = c.p->fields[t.op];
  const unsigned char* q = c.elem + f.off;
  switch (f.type) {
    case FT_I64: { i64 v; std::memcpy(&v, q, 8); r.is_int = true; r.i = v; } break;
    case FT_F64: { double v; std::memcpy(&v, q, 8); r.f = v; } break;
    case FT_I32: { int32_t v; std::memcpy(&v, q, 4); r.is_int = true; r.i = v; } break;
    default: { float v; std::memcpy(&v, q, 4); r.f = v; }
  }
  return r;
}

// Evaluate a variable-free subtree: `node(i, x, θ)` on Reals (graph.jl:305-318, register.jl:70,268-273).
Real eval_real(const Ctx& c, int n) {
  const TNode& t = c.p->t[n];
  Real r; r.is_int = false; r.i = 0; r.f = 0;
  switch (t.tag) {
    case T_CONST_I: case T_VAL: r.is_int = true; r.i = t.ipay; return r;
    case T_CONST_F: r.f = t.fpay; return r;
    case T_DATA_SELF: case T_DATA_FIELD: return data_value(c, t);
    case T_PAR: { Real ix = eval_real(c, t.c1); r.f = c.TH ? c.TH[ix.i - 1] : NAN; return r; }  // graph.jl:310-311
    case T_OP1: return real_op1(t.op, eval_real(c, t.c1));
    case T_OP2: return real_op2(t.op, eval_real(c, t.c1), eval_real(c, t.c2));
  }
  r.f = NAN; return r;
}

// Primal evaluation `f(itr[k], x, θ)` (graph.jl:305-318; register.jl:70-71,268-273).
double eval0(const Ctx& c, int n) {
  const TNode& t = c.p->t[n];
  if (t.kind == K_REAL) return eval_real(c, n).val();
  switch (t.tag) {
    case T_VAR: { Real ix = eval_real(c, t.c1); return c.X[ix.i - 1]; }
    case T_NULL: return t.fpay;                                   // graph.jl:497-498
    case T_OP1: { double f, d, dd; uni(t.op, eval0(c, t.c1), f, d, dd, 0); return f; }
    case T_OP2: {
      const TNode &a = c.p->t[t.c1], &b = c.p->t[t.c2];
      Real e1, e2; e1.is_int = e2.is_int = false; e1.i = e2.i = 0; e1.f = e2.f = 0;
      double x1, x2;
      if (a.kind == K_REAL) { e1 = eval_real(c, t.c1); x1 = e1.val(); } else x1 = eval0(c, t.c1);
      if (b.kind == K_REAL) { e2 = eval_real(c, t.c2); x2 = e2.val(); } else { x2 = eval0(c, t.c2); e2.f = x2; }
      Bi r; bi(t.op, x1, x2, e1, e2, r, false, false); return r.f;
    }
  }
  return NAN;
}

// Forward sweep under (Second)AdjointNodeSource (graph.jl:397-400,491-494;
// register.jl:65-68,174-266).  order = 1: x,y1,y2 ; order = 2: also h11,h12,h22.
void fwd(const Ctx& c, int n, int order) {
  const TNode& t = c.p->t[n];
  Work& w = *c.w;
  if (t.kind == K_REAL) { w.re[n] = eval_real(c, n); w.x[n] = w.re[n].val(); return; }
  switch (t.tag) {
    case T_VAR: { Real ix = eval_real(c, t.c1); w.vi[n] = ix.i; w.x[n] = c.X ? c.X[ix.i - 1] : NAN; return; }
    case T_NULL: w.x[n] = t.fpay; return;
    case T_OP1: {
      fwd(c, t.c1, order);
      uni(t.op, w.x[t.c1], w.x[n], w.y1[n], w.h11[n], order);
      return;
    }
    case T_OP2: {
      fwd(c, t.c1, order); fwd(c, t.c2, order);
      Real e1, e2; e1.is_int = e2.is_int = false; e1.i = e2.i = 0; e1.f = w.x[t.c1]; e2.f = w.x[t.c2];
      if (c.p->t[t.c1].kind == K_REAL) e1 = w.re[t.c1];
      if (c.p->t[t.c2].kind == K_REAL) e2 = w.re[t.c2];
      Bi r;
      if (t.kind == K_N2) {
        bi(t.op, w.x[t.c1], w.x[t.c2], e1, e2, r, true, true);
        w.x[n] = r.f; w.y1[n] = r.y1; w.y2[n] = r.y2; w.h11[n] = r.h11; w.h12[n] = r.h12; w.h22[n] = r.h22;
      } else if (t.fx == FX_SECOND) {       // node OP Real: (f, df1, ddf11)   register.jl:239-247
        bi(t.op, w.x[t.c1], w.x[t.c2], e1, e2, r, true, false);
        w.x[n] = r.f; w.y1[n] = r.y1; w.h11[n] = r.h11;
      } else {                              // Real OP node: (f, df2, ddf22)   register.jl:257-265
        bi(t.op, w.x[t.c1], w.x[t.c2], e1, e2, r, false, true);
        w.x[n] = r.f; w.y1[n] = r.y2; w.h11[n] = r.h22;
      }
      return;
    }
  }
}
inline int inner_of(const TNode& t) { return (t.tag == T_OP2 && t.fx == FX_FIRST) ? t.c2 : t.c1; }

// ---- reverse passes ---------------------------------------------------------------
// Sink: where a leaf visit lands.  mode selects the reference method the visit dispatches to.
enum { M_VALUES, M_PROBE, M_STRUCT, M_DENSE, M_JPROD, M_JTPROD, M_HPROD };
struct Sink {
  int mode;
  double* y; i64 off; const int* comp;            // values:  y[off + comp(++cnt)] += v
  std::vector<int>* raw1; std::vector<std::pair<int, int>>* raw2;   // probe
  i64 *rows, *cols; i64 row;                      // structure
  const double* v; double* out;                   // products
};

// grpass / jrpass / drpass share one shape (gradient.jl:11-26,71-90; jacobian.jl:16-40,69-83)
void rpass1(const Ctx& c, int n, Sink& s, int& cnt, double adj) {
  const TNode& t = c.p->t[n];
  const Work& w = *c.w;
  switch (t.kind) {
    case K_REAL: case K_NULL: return;                                  // gradient.jl:59-69
    case K_N1: rpass1(c, inner_of(t), s, cnt, adj * w.y1[n]); return;  // gradient.jl:71-74
    case K_N2:                                                         // gradient.jl:75-79
      rpass1(c, t.c1, s, cnt, adj * w.y1[n]);
      rpass1(c, t.c2, s, cnt, adj * w.y2[n]);
      return;
    case K_VAR:
      switch (s.mode) {
        case M_VALUES: s.y[s.off + s.comp[cnt++] - 1] += adj; return;  // gradient.jl:80-83, jacobian.jl:36-39 (1-based comp)
        case M_DENSE: s.y[w.vi[n] - 1] += adj; return;                 // gradient.jl:23-26
        case M_PROBE: s.raw1->push_back(c.p->t[n].ir); return;         // gradient.jl:84-87
        case M_STRUCT: { i64 ind = s.off + s.comp[cnt++] - 1;          // jacobian.jl:69-83
          s.rows[ind] = s.row; s.cols[ind] = w.vi[n]; } return;
        case M_JPROD: s.out[s.row - 1] += adj * s.v[w.vi[n] - 1]; cnt++; return;   // jacobian.jl:41-54
        case M_JTPROD: s.out[w.vi[n] - 1] += adj * s.v[s.row - 1]; cnt++; return;  // jacobian.jl:55-68
      }
  }
}

void hleaf_pair(const Ctx& c, int a, int b, Sink& s, int& cnt, double adj) {   // hessian.jl:251-315,520-532,622-642
  const Work& w = *c.w;
  i64 i = w.vi[a], j = w.vi[b];
  switch (s.mode) {
    case M_VALUES: s.y[s.off + s.comp[cnt++] - 1] += (i == j ? 2.0 * adj : adj); return;
    case M_PROBE: s.raw2->push_back(std::make_pair(c.p->t[a].ir, c.p->t[b].ir)); return;
    case M_STRUCT: { i64 ind = s.off + s.comp[cnt++] - 1;
      if (i >= j) { s.rows[ind] = i; s.cols[ind] = j; } else { s.rows[ind] = j; s.cols[ind] = i; } } return;
    case M_HPROD:
      if (i == j) s.out[i - 1] += 2.0 * adj * s.v[i - 1];
      else { s.out[i - 1] += adj * s.v[j - 1]; s.out[j - 1] += adj * s.v[i - 1]; }
      cnt++; return;
  }
}

// hdrpass: cross terms (df1/dx)(df2/dx)'  (hessian.jl:16-320)
void hdrpass(const Ctx& c, int a, int b, Sink& s, int& cnt, double adj) {
  const TNode &t1 = c.p->t[a], &t2 = c.p->t[b];
  const Work& w = *c.w;
  if (t1.kind == K_NULL || t2.kind == K_NULL) return;                       // :318-320
  if (t1.kind == K_VAR && t2.kind == K_VAR) { hleaf_pair(c, a, b, s, cnt, adj); return; }   // :251-268
  if (t1.kind == K_N1 && t2.kind == K_N1) { hdrpass(c, inner_of(t1), inner_of(t2), s, cnt, adj * w.y1[a] * w.y1[b]); return; } // :16-28
  if (t1.kind == K_VAR && t2.kind == K_N1) { hdrpass(c, a, inner_of(t2), s, cnt, adj * w.y1[b]); return; }   // :44-56
  if (t1.kind == K_N1 && t2.kind == K_VAR) { hdrpass(c, inner_of(t1), b, s, cnt, adj * w.y1[a]); return; }   // :72-84
  if (t1.kind == K_N2 && t2.kind == K_N2) {                                 // :100-115
    hdrpass(c, t1.c1, t2.c1, s, cnt, adj * w.y1[a] * w.y1[b]);
    hdrpass(c, t1.c1, t2.c2, s, cnt, adj * w.y1[a] * w.y2[b]);
    hdrpass(c, t1.c2, t2.c1, s, cnt, adj * w.y2[a] * w.y1[b]);
    hdrpass(c, t1.c2, t2.c2, s, cnt, adj * w.y2[a] * w.y2[b]);
    return;
  }
  if (t1.kind == K_N1 && t2.kind == K_N2) {                                 // :134-147
    hdrpass(c, inner_of(t1), t2.c1, s, cnt, adj * w.y1[a] * w.y1[b]);
    hdrpass(c, inner_of(t1), t2.c2, s, cnt, adj * w.y1[a] * w.y2[b]);
    return;
  }
  if (t1.kind == K_N2 && t2.kind == K_N1) {                                 // :163-176
    hdrpass(c, t1.c1, inner_of(t2), s, cnt, adj * w.y1[a] * w.y1[b]);
    hdrpass(c, t1.c2, inner_of(t2), s, cnt, adj * w.y2[a] * w.y1[b]);
    return;
  }
  if (t1.kind == K_VAR && t2.kind == K_N2) {                                // :192-205
    hdrpass(c, a, t2.c1, s, cnt, adj * w.y1[b]);
    hdrpass(c, a, t2.c2, s, cnt, adj * w.y2[b]);
    return;
  }
  if (t1.kind == K_N2 && t2.kind == K_VAR) {                                // :221-234
    hdrpass(c, t1.c1, b, s, cnt, adj * w.y1[a]);
    hdrpass(c, t1.c2, b, s, cnt, adj * w.y2[a]);
    return;
  }
}

// hrpass: d²f/dx² portion (hessian.jl:337-380,533-536,566-621)
void hrpass(const Ctx& c, int n, Sink& s, int& cnt, double adj, double adj2) {
  const TNode& t = c.p->t[n];
  const Work& w = *c.w;
  switch (t.kind) {
    case K_REAL: case K_NULL: return;                                       // :337-348
    case K_N1:                                                              // :349-362
      hrpass(c, inner_of(t), s, cnt, adj * w.y1[n], adj2 * sq(w.y1[n]) + adj * w.h11[n]);
      return;
    case K_N2: {                                                            // :363-380
      double adj2y1y2 = adj2 * w.y1[n] * w.y2[n];
      double adjh12 = adj * w.h12[n];
      hrpass(c, t.c1, s, cnt, adj * w.y1[n], adj2 * sq(w.y1[n]) + adj * w.h11[n]);
      hrpass(c, t.c2, s, cnt, adj * w.y2[n], adj2 * sq(w.y2[n]) + adj * w.h22[n]);
      hdrpass(c, t.c1, t.c2, s, cnt, adj2y1y2 + adjh12);
      return;
    }
    case K_VAR:
      switch (s.mode) {
        case M_VALUES: s.y[s.off + s.comp[cnt++] - 1] += adj2; return;      // :580-592
        case M_PROBE: s.raw2->push_back(std::make_pair(t.ir, t.ir)); return;  // :533-536
        case M_STRUCT: { i64 ind = s.off + s.comp[cnt++] - 1;               // :593-607
          s.rows[ind] = w.vi[n]; s.cols[ind] = w.vi[n]; } return;
        case M_HPROD: s.out[w.vi[n] - 1] += adj2 * s.v[w.vi[n] - 1]; cnt++; return;   // :566-579
      }
  }
}

// hrpass0: top-level linear peeling (hessian.jl:382-517)
void hrpass0(const Ctx& c, int n, Sink& s, int& cnt, double adj, double adj2) {
  const TNode& t = c.p->t[n];
  const Work& w = *c.w;
  if (t.kind == K_VAR) return;                                              // :494-517
  if (t.kind == K_N1) {
    int in = inner_of(t);
    if (t.tag == T_OP2) {                                                   // FirstFixed / SecondFixed flavours
      if (t.op == B_MUL) { hrpass0(c, in, s, cnt, adj * w.y1[n], adj2 * sq(w.y1[n])); return; }   // :385-397
      if (t.op == B_ADD) { hrpass0(c, in, s, cnt, adj, adj2); return; }                           // :398-410
      if (t.op == B_SUB && t.fx == FX_FIRST) { hrpass0(c, in, s, cnt, -adj, adj2); return; }      // :411-423
      if (t.op == B_SUB && t.fx == FX_SECOND) { hrpass0(c, in, s, cnt, adj, adj2); return; }      // :424-436
    } else {
      if (t.op == U_PLUS) { hrpass0(c, in, s, cnt, adj, adj2); return; }    // :438-450
      if (t.op == U_MINUS) { hrpass0(c, in, s, cnt, -adj, adj2); return; }  // :451-463
    }
  }
  if (t.kind == K_N2 && t.op == B_ADD) {                                    // :465-478
    hrpass0(c, t.c1, s, cnt, adj, adj2); hrpass0(c, t.c2, s, cnt, adj, adj2); return;
  }
  if (t.kind == K_N2 && t.op == B_SUB) {                                    // :480-493
    hrpass0(c, t.c1, s, cnt, adj, adj2); hrpass0(c, t.c2, s, cnt, -adj, adj2); return;
  }
  hrpass(c, n, s, cnt, adj, adj2);                                          // :382
}

// ---- pattern build: probe + compressors (simdfunction.jl:66-100) ---------------------------
void probe(Pattern& p) {
  Work w; w.size(p.t.size());
  Ctx c; c.p = &p; c.X = nullptr; c.TH = nullptr; c.w = &w; c.k = 0; c.elem = nullptr;
  // The probe evaluates at `Identity()`: data leaves answer NaN and no index is ever
  // evaluated (graph.jl:307-308,317-318).  Kinds are static, so only the traversal matters.
  std::vector<int> raw1; std::vector<std::pair<int, int>> raw2;
  Sink s; std::memset(&s, 0, sizeof s); s.mode = M_PROBE; s.raw1 = &raw1; s.raw2 = &raw2;
  int cnt = 0;
  // traversal uses only kinds; feed NaN tapes
  std::fill(w.y1.begin(), w.y1.end(), NAN); std::fill(w.y2.begin(), w.y2.end(), NAN);
  std::fill(w.h11.begin(), w.h11.end(), NAN); std::fill(w.h12.begin(), w.h12.end(), NAN); std::fill(w.h22.begin(), w.h22.end(), NAN);
  rpass1(c, p.root, s, cnt, NAN);                                           // simdfunction.jl:81-83
  cnt = 0;
  hrpass0(c, p.root, s, cnt, NAN, NAN);                                     // simdfunction.jl:85-87
  std::vector<int> u1;                                                      // _ident_unique, :66-76
  for (int v : raw1) {
    int found = -1;
    for (size_t q = 0; q < u1.size(); q++) if (ir_equal(p, u1[q], v)) { found = (int)q; break; }
    if (found < 0) { u1.push_back(v); found = (int)u1.size() - 1; }
    p.comp1.push_back(found + 1);
  }
  p.o1step = (int)u1.size();
  std::vector<std::pair<int, int>> u2;
  for (auto& v : raw2) {
    int found = -1;
    for (size_t q = 0; q < u2.size(); q++)
      if (ir_equal(p, u2[q].first, v.first) && ir_equal(p, u2[q].second, v.second)) { found = (int)q; break; }
    if (found < 0) { u2.push_back(v); found = (int)u2.size() - 1; }
    p.comp2.push_back(found + 1);
  }
  p.o2step = (int)u2.size();
}

bool parse(Model& m, const i64* w, size_t nw, const void* const* bufs, int nbufs) {
  size_t q = 0;
  auto rd = [&](i64& out) { if (q >= nw) return false; out = w[q++]; return true; };
  i64 magic = 0, ver = 0, npat = 0, nb = 0;
  if (!rd(magic) || magic != 0x0031425845LL) { m.err = "bad IR magic"; return false; }
  if (!rd(ver) || ver != 1) { m.err = "bad IR version"; return false; }
  rd(m.nvar); rd(m.npar); rd(npat); rd(nb);
  if (nb > nbufs) { m.err = "IR references more data buffers than were passed"; return false; }
  m.pats.resize((size_t)npat);
  for (auto& p : m.pats) {
    i64 v = 0, nf = 0, nidx = 0, nn = 0;
    rd(v); p.kind = (int)v; rd(p.nitr); rd(v); p.itr_kind = (int)v; rd(p.range_start);
    rd(v); p.databuf = (int)v; rd(p.stride); rd(nf);
    p.fields.resize((size_t)nf);
    for (auto& f : p.fields) { rd(f.off); rd(f.type); }
    rd(p.o0); rd(p.o1); rd(p.o2); rd(v); p.base = (int)v; rd(nidx);
    p.idx_roots_ir.resize((size_t)nidx); p.dims.resize((size_t)nidx);
    for (auto& r : p.idx_roots_ir) { rd(v); r = (int)v; }
    for (auto& d : p.dims) rd(d);
    rd(nn); p.ir.resize((size_t)nn);
    for (auto& n : p.ir) { rd(n.tag); rd(n.a); rd(n.b); rd(n.payload); }
    rd(v); p.root_ir = (int)v;
    i64 nc1 = 0, nc2 = 0; rd(nc1); q += (size_t)nc1; rd(nc2); q += (size_t)nc2;   // supplied comps are ignored: recomputed
    if (q > nw) { m.err = "truncated IR"; return false; }
    p.data = p.databuf >= 0 ? (const unsigned char*)bufs[p.databuf] : nullptr;
    p.root = expand(p, p.root_ir);
    for (int r : p.idx_roots_ir) p.idx_roots.push_back(expand(p, r));
    probe(p);
  }
  // running counters in add order (nlp.jl:1474-1482, 1597-1611, 1730-1738)
  m.ncon = m.nobj = m.nconaug = m.nnzg = m.nnzj = m.nnzh = 0;
  for (auto& p : m.pats) {
    if (p.kind == KIND_OBJ) {
      p.o0 = m.nobj; p.o1 = m.nnzg; p.o2 = m.nnzh;                          // nlp.jl:1450
      m.nobj += p.nitr; m.nnzg += p.nitr * p.o1step; m.nnzh += p.nitr * p.o2step;
    } else if (p.kind == KIND_CON) {
      p.o0 = m.ncon; p.o1 = m.nnzj; p.o2 = m.nnzh;                          // nlp.jl:1587
      m.ncon += p.nitr; m.nnzj += p.nitr * p.o1step; m.nnzh += p.nitr * p.o2step;
    } else {
      p.o0 = m.pats[(size_t)p.base].o0; p.o1 = m.nnzj; p.o2 = m.nnzh;       // nlp.jl:1683 (offset0(c1, 0))
      m.nconaug += p.nitr; m.nnzj += p.nitr * p.o1step; m.nnzh += p.nitr * p.o2step;
    }
  }
  return true;
}

inline void point(Ctx& c, i64 k) {
  c.k = k;
  c.elem = c.p->data ? c.p->data + (size_t)k * (size_t)c.p->stride : nullptr;
}

// offset0 (nlp.jl:1980-2001): 1-based global row (constraints) / objbuffer slot (objectives)
i64 offset0(const Ctx& c) {
  const Pattern& p = *c.p;
  if (p.kind != KIND_AUG) return p.o0 + c.k + 1;                            // nlp.jl:1989
  if (p.idx_roots.size() == 1) return p.o0 + eval_real(c, p.idx_roots[0]).i;   // nlp.jl:1994-1997
  i64 a = 1, lin = 0;                                                       // idxx, nlp.jl:2012-2015
  for (size_t d = 0; d < p.idx_roots.size(); d++) {
    lin += a * (eval_real(c, p.idx_roots[d]).i - 1);
    a *= p.dims[d];
  }
  return p.o0 + lin + 1;
}

int nthreads_of(const Model& m) { return m.nthreads > 0 ? m.nthreads : 1; }

inline void shard_of(const Model& m, const Pattern& p, i64& lo, i64& hi) {
  lo = p.nitr * m.rank / m.world; hi = p.nitr * (m.rank + 1) / m.world;
}

template <class F>
void for_points(const Model& m, const Pattern& p, const double* X, F body) {
  int nt = nthreads_of(m);
  i64 s_lo, s_hi; shard_of(m, p, s_lo, s_hi);
  const i64 n = s_hi - s_lo;
  if (nt <= 1 || n < 1024) {
    Work w; w.size(p.t.size());
    Ctx c; c.p = &p; c.X = X; c.TH = m.theta.data(); c.w = &w;
    for (i64 k = s_lo; k < s_hi; k++) { point(c, k); body(c); }
    return;
  }
  // static partition of the data points over host threads: the shape of
  // KernelAbstractions.CPU() under `julia -t N` (docs/src/gpu.jl:2-5,48)
  std::vector<std::thread> th;
  for (int r = 0; r < nt; r++) {
    i64 lo = s_lo + n * r / nt, hi = s_lo + n * (r + 1) / nt;
    th.emplace_back([&, lo, hi]() {
      Work w; w.size(p.t.size());
      Ctx c; c.p = &p; c.X = X; c.TH = m.theta.data(); c.w = &w;
      for (i64 k = lo; k < hi; k++) { point(c, k); body(c); }
    });
  }
  for (auto& t : th) t.join();
}

// sequential visit of the (sharded) points of one pattern
template <class F>
void seq_points(const Model& m, const Pattern& p, const double* X, F body) {
  i64 lo, hi; shard_of(m, p, lo, hi);
  Work w; w.size(p.t.size());
  Ctx c; c.p = &p; c.X = X; c.TH = m.theta.data(); c.w = &w;
  for (i64 k = lo; k < hi; k++) { point(c, k); body(c); }
}

}  // namespace

// =============================================================================
// C interface (ctypes: tests/oracle_api.py)
// =============================================================================
extern "C" {

void* ora_create(const void* ir, size_t ir_bytes, const void* const* bufs, int nbufs) {
  Model* m = new Model(); m->nthreads = 1;
  parse(*m, (const i64*)ir, ir_bytes / 8, bufs, nbufs);
  m->theta.assign((size_t)m->npar, 0.0);
  return m;
}
const char* ora_error(void* h) { return ((Model*)h)->err.c_str(); }
void ora_destroy(void* h) { delete (Model*)h; }
void ora_set_threads(void* h, int n) { ((Model*)h)->nthreads = n; }
void ora_set_shard(void* h, int rank, int world) { ((Model*)h)->rank = rank; ((Model*)h)->world = world; }
int ora_max_threads() { unsigned n = std::thread::hardware_concurrency(); return n ? (int)n : 1; }
void ora_set_params(void* h, const double* th) { Model* m = (Model*)h; m->theta.assign(th, th + m->npar); }

// out[0..7] = nvar, ncon, nnzj, nnzh, nobj, nnzg, nconaug, npar
void ora_dims(void* h, i64* out) {
  Model* m = (Model*)h;
  out[0] = m->nvar; out[1] = m->ncon; out[2] = m->nnzj; out[3] = m->nnzh;
  out[4] = m->nobj; out[5] = m->nnzg; out[6] = m->nconaug; out[7] = m->npar;
}
int ora_npatterns(void* h) { return (int)((Model*)h)->pats.size(); }
// out[0..6] = kind, nitr, o0, o1, o2, o1step, o2step ; returns sizes of comp1/comp2 in out[7], out[8]
void ora_pattern_info(void* h, int k, i64* out) {
  const Pattern& p = ((Model*)h)->pats[(size_t)k];
  out[0] = p.kind; out[1] = p.nitr; out[2] = p.o0; out[3] = p.o1; out[4] = p.o2;
  out[5] = p.o1step; out[6] = p.o2step; out[7] = (i64)p.comp1.size(); out[8] = (i64)p.comp2.size();
}
void ora_pattern_comp(void* h, int k, int which, i64* out) {
  const Pattern& p = ((Model*)h)->pats[(size_t)k];
  const std::vector<int>& c = which == 1 ? p.comp1 : p.comp2;
  for (size_t q = 0; q < c.size(); q++) out[q] = c[q];
}

// obj (nlp.jl:1827-1839): sequential sum, patterns oldest first, points in order.
double ora_obj(void* h, const double* x) {
  Model* m = (Model*)h;
  double s = 0.0;
  for (auto& p : m->pats) {
    if (p.kind != KIND_OBJ) continue;
    if (nthreads_of(*m) <= 1) {
      seq_points(*m, p, x, [&](Ctx& c) { s += eval0(c, c.p->root); });
    } else {   // KA shape: objbuffer + sum (ext:253-271)
      std::vector<double> buf((size_t)p.nitr, 0.0);
      for_points(*m, p, x, [&](Ctx& c) { buf[(size_t)c.k] = eval0(c, c.p->root); });
      for (double v : buf) s += v;
    }
  }
  return s;
}

// cons_nln! (nlp.jl:1841-1854)
void ora_cons(void* h, const double* x, double* g) {
  Model* m = (Model*)h;
  std::fill(g, g + m->ncon, 0.0);
  for (auto& p : m->pats) {
    if (p.kind == KIND_OBJ) continue;
    if (p.kind == KIND_CON) for_points(*m, p, x, [&](Ctx& c) { g[offset0(c) - 1] += eval0(c, c.p->root); });
    else {   // augmentation rows collide across points: sequential (CPU path is sequential, nlp.jl:1849-1851)
      seq_points(*m, p, x, [&](Ctx& c) { g[offset0(c) - 1] += eval0(c, c.p->root); });
    }
  }
}

// grad! (nlp.jl:1858-1868 -> gradient.jl:39-49 -> drpass)
void ora_grad(void* h, const double* x, double* g) {
  Model* m = (Model*)h;
  std::fill(g, g + m->nvar, 0.0);
  for (auto& p : m->pats) {
    if (p.kind != KIND_OBJ) continue;
    Sink s; std::memset(&s, 0, sizeof s); s.mode = M_DENSE; s.y = g;
    seq_points(*m, p, x, [&](Ctx& c) { fwd(c, c.p->root, 1); int cnt = 0; rpass1(c, c.p->root, s, cnt, 1.0); });
  }
}

// sparse gradient slots (ext:310-336 without the final compress): gradbuffer[nnzg]
void ora_sgrad(void* h, const double* x, double* gb) {
  Model* m = (Model*)h;
  std::fill(gb, gb + m->nnzg, 0.0);
  for (auto& p : m->pats) {
    if (p.kind != KIND_OBJ) continue;
    for_points(*m, p, x, [&](Ctx& c) {
      const Pattern& q = *c.p;
      fwd(c, q.root, 1);
      Sink s; std::memset(&s, 0, sizeof s); s.mode = M_VALUES; s.y = gb; s.comp = q.comp1.data();
      s.off = q.o1 + q.o1step * c.k; int cnt = 0; rpass1(c, q.root, s, cnt, 1.0);
    });
  }
}

// jac_coord! (nlp.jl:1870-1880 -> jacobian.jl:112-132)
void ora_jac(void* h, const double* x, double* jac) {
  Model* m = (Model*)h;
  std::fill(jac, jac + m->nnzj, 0.0);
  for (auto& p : m->pats) {
    if (p.kind == KIND_OBJ) continue;
    for_points(*m, p, x, [&](Ctx& c) {
      const Pattern& q = *c.p;
      fwd(c, q.root, 1);
      Sink s; std::memset(&s, 0, sizeof s); s.mode = M_VALUES; s.y = jac; s.comp = q.comp1.data();
      s.off = q.o1 + q.o1step * c.k; int cnt = 0; rpass1(c, q.root, s, cnt, 1.0);
    });
  }
}

// jac_structure! (nlp.jl:1798-1807 -> jacobian.jl:69-83): 1-based rows/cols
void ora_jac_structure(void* h, i64* rows, i64* cols) {
  Model* m = (Model*)h;
  for (auto& p : m->pats) {
    if (p.kind == KIND_OBJ) continue;
    for_points(*m, p, nullptr, [&](Ctx& c) {
      const Pattern& q = *c.p;
      fwd(c, q.root, 1);
      Sink s; std::memset(&s, 0, sizeof s); s.mode = M_STRUCT; s.rows = rows; s.cols = cols; s.comp = q.comp1.data();
      s.off = q.o1 + q.o1step * c.k; s.row = offset0(c); int cnt = 0; rpass1(c, q.root, s, cnt, NAN);
    });
  }
}

// hess_coord! (nlp.jl:1906-1940 -> hessian.jl:681-717).  y == NULL: objective-only form (nlp.jl:1906-1915).
void ora_hess(void* h, const double* x, const double* y, double obj_weight, double* hess) {
  Model* m = (Model*)h;
  std::fill(hess, hess + m->nnzh, 0.0);
  for (auto& p : m->pats) {
    if (p.kind != KIND_OBJ && !y) continue;
    for_points(*m, p, x, [&](Ctx& c) {
      const Pattern& q = *c.p;
      fwd(c, q.root, 2);
      double adj1 = q.kind == KIND_OBJ ? obj_weight : y[offset0(c) - 1];     // hessian.jl:708
      Sink s; std::memset(&s, 0, sizeof s); s.mode = M_VALUES; s.y = hess; s.comp = q.comp2.data();
      s.off = q.o2 + q.o2step * c.k; int cnt = 0; hrpass0(c, q.root, s, cnt, adj1, 0.0);
    });
  }
}

// hess_structure! (nlp.jl:1809-1825): lower triangle, 1-based
void ora_hess_structure(void* h, i64* rows, i64* cols) {
  Model* m = (Model*)h;
  for (auto& p : m->pats) {
    for_points(*m, p, nullptr, [&](Ctx& c) {
      const Pattern& q = *c.p;
      fwd(c, q.root, 2);
      Sink s; std::memset(&s, 0, sizeof s); s.mode = M_STRUCT; s.rows = rows; s.cols = cols; s.comp = q.comp2.data();
      s.off = q.o2 + q.o2step * c.k; int cnt = 0; hrpass0(c, q.root, s, cnt, NAN, NAN);
    });
  }
}

// jprod_nln! / jtprod_nln! / hprod! (nlp.jl:1882-1978): sequential (rows/cols collide)
void ora_jprod(void* h, const double* x, const double* v, double* Jv) {
  Model* m = (Model*)h;
  std::fill(Jv, Jv + m->ncon, 0.0);
  for (auto& p : m->pats) {
    if (p.kind == KIND_OBJ) continue;
    Work w; w.size(p.t.size());
    Ctx c; c.p = &p; c.X = x; c.TH = m->theta.data(); c.w = &w;
    i64 lo, hi; shard_of(*m, p, lo, hi);   // shard mode (test aid): partial product of this shard's points
    for (i64 k = lo; k < hi; k++) {
      point(c, k); fwd(c, p.root, 1);
      Sink s; std::memset(&s, 0, sizeof s); s.mode = M_JPROD; s.v = v; s.out = Jv; s.row = offset0(c);
      int cnt = 0; rpass1(c, p.root, s, cnt, 1.0);
    }
  }
}
void ora_jtprod(void* h, const double* x, const double* v, double* Jtv) {
  Model* m = (Model*)h;
  std::fill(Jtv, Jtv + m->nvar, 0.0);
  for (auto& p : m->pats) {
    if (p.kind == KIND_OBJ) continue;
    Work w; w.size(p.t.size());
    Ctx c; c.p = &p; c.X = x; c.TH = m->theta.data(); c.w = &w;
    i64 lo, hi; shard_of(*m, p, lo, hi);   // shard mode (test aid): partial product of this shard's points
    for (i64 k = lo; k < hi; k++) {
      point(c, k); fwd(c, p.root, 1);
      Sink s; std::memset(&s, 0, sizeof s); s.mode = M_JTPROD; s.v = v; s.out = Jtv; s.row = offset0(c);
      int cnt = 0; rpass1(c, p.root, s, cnt, 1.0);
    }
  }
}
void ora_hprod(void* h, const double* x, const double* y, const double* v, double obj_weight, double* Hv) {
  Model* m = (Model*)h;
  std::fill(Hv, Hv + m->nvar, 0.0);
  for (auto& p : m->pats) {
    if (p.kind != KIND_OBJ && !y) continue;
    Work w; w.size(p.t.size());
    Ctx c; c.p = &p; c.X = x; c.TH = m->theta.data(); c.w = &w;
    i64 lo, hi; shard_of(*m, p, lo, hi);   // shard mode (test aid): partial product of this shard's points
    for (i64 k = lo; k < hi; k++) {
      point(c, k); fwd(c, p.root, 2);
      double adj1 = p.kind == KIND_OBJ ? obj_weight : y[offset0(c) - 1];
      Sink s; std::memset(&s, 0, sizeof s); s.mode = M_HPROD; s.v = v; s.out = Hv;
      int cnt = 0; hrpass0(c, p.root, s, cnt, adj1, 0.0);
    }
  }
}

// Scalar tables for the ADTest-style pins (test/ADTest/ADTest.jl:298-342)
void ora_uni(int op, double x, double* out) { uni(op, x, out[0], out[1], out[2], 1); }
void ora_bi(int op, double x1, double x2, int e2_is_int, double* out) {
  Real e1, e2; e1.is_int = false; e1.i = 0; e1.f = x1;
  e2.is_int = e2_is_int != 0; e2.i = (i64)x2; e2.f = x2;
  Bi r; bi(op, x1, x2, e1, e2, r, true, true);
  out[0] = r.f; out[1] = r.y1; out[2] = r.y2; out[3] = r.h11; out[4] = r.h12; out[5] = r.h22;
}

}  // extern "C"

// =============================================================================
// Compiled CPU baseline ("port-compiled").  The reference's CPU path is not an interpreter: Julia specialises the
// tape types on the expression tree and compiles `for k in itr; hrpass0(f(k, SecondAdjointNodeSource(x)), ...)` to
// native straight-line code (src/hessian.jl:681-712).  ora_emit_source writes the same thing for this restatement:
// for every pattern, the forward sweep and the reverse recursion above UNROLLED over the oracle's own expanded tree
// (`Pattern::t`), with the derivative tables of ora_tables.hpp inlined at a constant `op`.  Same expressions in the
// same order as fwd / hrpass0 / hrpass / hdrpass, compiled with the same flags (-ffp-contract=off), so the values
// are bit-identical to the interpreter's (tests/test_oracle_compiled.py).  Nothing of the product's generator
// (examodels.jl_b200/csrc/exb_plan.hpp) is used.
// =============================================================================
namespace {

struct Emit {
  const Pattern& p; int pi; std::string o; int ntmp = 0;
  Emit(const Pattern& pp, int idx) : p(pp), pi(idx) {}
  static std::string num(double v) { char b[64]; if (std::isnan(v)) return "NAN"; if (std::isinf(v)) return v > 0 ? "INFINITY" : "(-INFINITY)";
    snprintf(b, sizeof b, "%a", v); return std::string("(") + b + ")"; }
  static std::string S(long long v) { return std::to_string(v); }
  std::string tmp() { return "a" + S(ntmp++); }
  void line(const std::string& l) { o += "  " + l + "\n"; }
  // a variable-free subtree as a `Real` (eval_real)
  void real(int n, std::vector<char>& done) {
    if (done[(size_t)n]) return;
    done[(size_t)n] = 1;
    const TNode& t = p.t[(size_t)n];
    const std::string r = "r" + S(n);
    switch (t.tag) {
      case T_CONST_I: case T_VAL: line("const Real " + r + "{true, " + S(t.ipay) + "LL, 0.0};"); break;
      case T_CONST_F: line("const Real " + r + "{false, 0, " + num(t.fpay) + "};"); break;
      case T_DATA_SELF: line("const Real " + r + "{true, " + S(p.range_start) + "LL + k, 0.0};"); break;
      case T_DATA_FIELD: { const Field& f = p.fields[(size_t)t.op];
        line("const Real " + r + " = ld_field<" + S(f.type) + ">(elem + " + S(f.off) + ");"); } break;
      case T_PAR: real(t.c1, done); line("const Real " + r + "{false, 0, TH[r" + S(t.c1) + ".i - 1]};"); break;
      case T_OP1: real(t.c1, done); line("const Real " + r + " = real_op1(" + S(t.op) + ", r" + S(t.c1) + ");"); break;
      case T_OP2: real(t.c1, done); real(t.c2, done);
        line("const Real " + r + " = real_op2(" + S(t.op) + ", r" + S(t.c1) + ", r" + S(t.c2) + ");"); break;
      default: line("const Real " + r + "{false, 0, NAN};");
    }
  }
  // forward sweep (fwd): x / y1 / y2 / h11 / h12 / h22 per tree node as locals
  void fwd_(int n, int order, std::vector<char>& done) {
    const TNode& t = p.t[(size_t)n];
    const std::string N = S(n);
    if (t.kind == K_REAL) { real(n, done); line("const double x" + N + " = r" + N + ".val();"); return; }
    switch (t.tag) {
      case T_VAR: real(t.c1, done); line("const i64 v" + N + " = r" + S(t.c1) + ".i; const double x" + N + " = X[v" + N + " - 1];"); return;
      case T_NULL: line("const double x" + N + " = " + num(t.fpay) + ";"); return;
      case T_OP1:
        fwd_(t.c1, order, done);
        line("double x" + N + ", y1_" + N + ", h11_" + N + "; uni(" + S(t.op) + ", x" + S(t.c1) + ", x" + N + ", y1_" + N + ", h11_" + N + ", " + S(order) + ");");
        return;
      case T_OP2: {
        fwd_(t.c1, order, done); fwd_(t.c2, order, done);
        const bool r1 = p.t[(size_t)t.c1].kind == K_REAL, r2 = p.t[(size_t)t.c2].kind == K_REAL;
        line("const Real e1_" + N + (r1 ? " = r" + S(t.c1) + ";" : "{false, 0, x" + S(t.c1) + "};"));
        line("const Real e2_" + N + (r2 ? " = r" + S(t.c2) + ";" : "{false, 0, x" + S(t.c2) + "};"));
        const char* want = t.kind == K_N2 ? "true, true" : t.fx == FX_SECOND ? "true, false" : "false, true";
        line("Bi b" + N + "; bi(" + S(t.op) + ", x" + S(t.c1) + ", x" + S(t.c2) + ", e1_" + N + ", e2_" + N + ", b" + N + ", " + want + ");");
        if (t.kind == K_N2)
          line("const double x" + N + " = b" + N + ".f, y1_" + N + " = b" + N + ".y1, y2_" + N + " = b" + N + ".y2, h11_" + N + " = b" + N + ".h11, h12_" + N +
               " = b" + N + ".h12, h22_" + N + " = b" + N + ".h22;");
        else if (t.fx == FX_SECOND) line("const double x" + N + " = b" + N + ".f, y1_" + N + " = b" + N + ".y1, h11_" + N + " = b" + N + ".h11;");
        else line("const double x" + N + " = b" + N + ".f, y1_" + N + " = b" + N + ".y2, h11_" + N + " = b" + N + ".h22;");
        return;
      }
    }
  }
  int inner(int n) const { return inner_of(p.t[(size_t)n]); }
  std::string def(const std::string& expr) { std::string a = tmp(); line("const double " + a + " = " + expr + ";"); return a; }
  // hdrpass / hrpass / hrpass0 with symbolic adjoints: one local per product, in the interpreter's order
  void hd(int a, int b, int& cnt, const std::string& adj) {
    const TNode &t1 = p.t[(size_t)a], &t2 = p.t[(size_t)b];
    const std::string A = S(a), B = S(b);
    if (t1.kind == K_NULL || t2.kind == K_NULL || t1.kind == K_REAL || t2.kind == K_REAL) return;
    if (t1.kind == K_VAR && t2.kind == K_VAR) {
      line("s[" + S(p.comp2[(size_t)cnt++] - 1) + "] += (v" + A + " == v" + B + " ? 2.0 * " + adj + " : " + adj + ");"); return; }
    if (t1.kind == K_N1 && t2.kind == K_N1) { hd(inner(a), inner(b), cnt, def(adj + " * y1_" + A + " * y1_" + B)); return; }
    if (t1.kind == K_VAR && t2.kind == K_N1) { hd(a, inner(b), cnt, def(adj + " * y1_" + B)); return; }
    if (t1.kind == K_N1 && t2.kind == K_VAR) { hd(inner(a), b, cnt, def(adj + " * y1_" + A)); return; }
    if (t1.kind == K_N2 && t2.kind == K_N2) {
      hd(t1.c1, t2.c1, cnt, def(adj + " * y1_" + A + " * y1_" + B)); hd(t1.c1, t2.c2, cnt, def(adj + " * y1_" + A + " * y2_" + B));
      hd(t1.c2, t2.c1, cnt, def(adj + " * y2_" + A + " * y1_" + B)); hd(t1.c2, t2.c2, cnt, def(adj + " * y2_" + A + " * y2_" + B)); return; }
    if (t1.kind == K_N1 && t2.kind == K_N2) {
      hd(inner(a), t2.c1, cnt, def(adj + " * y1_" + A + " * y1_" + B)); hd(inner(a), t2.c2, cnt, def(adj + " * y1_" + A + " * y2_" + B)); return; }
    if (t1.kind == K_N2 && t2.kind == K_N1) {
      hd(t1.c1, inner(b), cnt, def(adj + " * y1_" + A + " * y1_" + B)); hd(t1.c2, inner(b), cnt, def(adj + " * y2_" + A + " * y1_" + B)); return; }
    if (t1.kind == K_VAR && t2.kind == K_N2) { hd(a, t2.c1, cnt, def(adj + " * y1_" + B)); hd(a, t2.c2, cnt, def(adj + " * y2_" + B)); return; }
    if (t1.kind == K_N2 && t2.kind == K_VAR) { hd(t1.c1, b, cnt, def(adj + " * y1_" + A)); hd(t1.c2, b, cnt, def(adj + " * y2_" + A)); return; }
  }
  void hr(int n, int& cnt, const std::string& adj, const std::string& adj2) {
    const TNode& t = p.t[(size_t)n];
    const std::string N = S(n);
    switch (t.kind) {
      case K_REAL: case K_NULL: return;
      case K_N1: { const std::string a = def(adj + " * y1_" + N), b = def(adj2 + " * sq(y1_" + N + ") + " + adj + " * h11_" + N);
        hr(inner(n), cnt, a, b); return; }
      case K_N2: {
        const std::string c1 = def(adj2 + " * y1_" + N + " * y2_" + N), c2 = def(adj + " * h12_" + N);
        const std::string a1 = def(adj + " * y1_" + N), b1 = def(adj2 + " * sq(y1_" + N + ") + " + adj + " * h11_" + N);
        hr(t.c1, cnt, a1, b1);
        const std::string a2 = def(adj + " * y2_" + N), b2 = def(adj2 + " * sq(y2_" + N + ") + " + adj + " * h22_" + N);
        hr(t.c2, cnt, a2, b2);
        hd(t.c1, t.c2, cnt, def(c1 + " + " + c2));
        return; }
      case K_VAR: line("s[" + S(p.comp2[(size_t)cnt++] - 1) + "] += " + adj2 + ";"); return;
    }
  }
  void hr0(int n, int& cnt, const std::string& adj, const std::string& adj2) {
    const TNode& t = p.t[(size_t)n];
    const std::string N = S(n);
    if (t.kind == K_VAR) return;
    if (t.kind == K_N1) {
      const int in = inner(n);
      if (t.tag == T_OP2) {
        if (t.op == B_MUL) { const std::string a = def(adj + " * y1_" + N), b = def(adj2 + " * sq(y1_" + N + ")"); hr0(in, cnt, a, b); return; }
        if (t.op == B_ADD) { hr0(in, cnt, adj, adj2); return; }
        if (t.op == B_SUB && t.fx == FX_FIRST) { hr0(in, cnt, def("-" + adj), adj2); return; }
        if (t.op == B_SUB && t.fx == FX_SECOND) { hr0(in, cnt, adj, adj2); return; }
      } else {
        if (t.op == U_PLUS) { hr0(in, cnt, adj, adj2); return; }
        if (t.op == U_MINUS) { hr0(in, cnt, def("-" + adj), adj2); return; }
      }
    }
    if (t.kind == K_N2 && t.op == B_ADD) { hr0(t.c1, cnt, adj, adj2); hr0(t.c2, cnt, adj, adj2); return; }
    if (t.kind == K_N2 && t.op == B_SUB) { hr0(t.c1, cnt, adj, adj2); hr0(t.c2, cnt, def("-" + adj), adj2); return; }
    hr(n, cnt, adj, adj2);
  }
  // one pattern: `static void hess_pK(k, X, Y, TH, elem, w, hess)` = the body of the for_points lambda of ora_hess
  std::string hess_fn() {
    o.clear(); ntmp = 0;
    o += "static inline void hess_p" + S(pi) + "(const i64 k, const double* __restrict__ X, const double* __restrict__ Y, const double* __restrict__ TH, "
         "const unsigned char* __restrict__ elem, const double w, double* __restrict__ hess) {\n";
    std::vector<char> done(p.t.size(), 0);
    if (p.o2step > 0) {
      fwd_(p.root, 2, done);
      // adj1 = obj_weight | y[offset0]  (hessian.jl:708; offset0: nlp.jl:1980-2001)
      if (p.kind == KIND_OBJ) line("const double adj1 = w;");
      else if (p.kind == KIND_CON) line("const double adj1 = Y[" + S(p.o0) + "LL + k];");
      else if (p.idx_roots.size() == 1) { real(p.idx_roots[0], done); line("const double adj1 = Y[" + S(p.o0) + "LL + r" + S(p.idx_roots[0]) + ".i - 1];"); }
      else {
        std::string lin = "0", a = "1";
        for (size_t d = 0; d < p.idx_roots.size(); d++) {
          real(p.idx_roots[d], done);
          lin = "(" + lin + " + " + a + " * (r" + S(p.idx_roots[d]) + ".i - 1))"; a = "(" + a + " * " + S(p.dims[d]) + "LL)";
        }
        line("const double adj1 = Y[" + S(p.o0) + "LL + " + lin + "];");
      }
      line("double s[" + S(p.o2step) + "] = {0.0};");
      line("const double zero = 0.0;");
      int cnt = 0;
      hr0(p.root, cnt, "adj1", "zero");
      line("double* __restrict__ out = hess + " + S(p.o2) + "LL + " + S(p.o2step) + "LL * k;");
      for (int j = 0; j < p.o2step; j++) line("out[" + S(j) + "] = s[" + S(j) + "];");
    }
    o += "}\n";
    return o;
  }
};

std::string emit_source(const Model& m) {
  std::string o;
  o += "// generated by oracle/exa_oracle.cpp (ora_emit_source): per-pattern straight-line port of the CPU restatement -- TEST INFRASTRUCTURE ONLY\n";
  o += "#include \"ora_tables.hpp\"\nusing namespace ora;\ntypedef int64_t i64;\nnamespace {\n";
  o += "template <int T> inline Real ld_field(const unsigned char* q) {\n  Real r{false, 0, 0.0};\n"
       "  if (T == FT_I64) { i64 v; std::memcpy(&v, q, 8); r.is_int = true; r.i = v; }\n  else if (T == FT_F64) { double v; std::memcpy(&v, q, 8); r.f = v; }\n"
       "  else if (T == FT_I32) { int32_t v; std::memcpy(&v, q, 4); r.is_int = true; r.i = v; }\n  else { float v; std::memcpy(&v, q, 4); r.f = v; }\n  return r;\n}\n";
  for (size_t k = 0; k < m.pats.size(); k++) { Emit e(m.pats[k], (int)k); o += e.hess_fn(); }
  o += "template <class F> void run(i64 lo, i64 hi, int nt, F body) {\n  const i64 n = hi - lo;\n  if (nt <= 1 || n < 1024) { for (i64 k = lo; k < hi; k++) body(k); return; }\n"
       "  std::vector<std::thread> th;\n  for (int r = 0; r < nt; r++) { const i64 a = lo + n * r / nt, b = lo + n * (r + 1) / nt; th.emplace_back([=]() { for (i64 k = a; k < b; k++) body(k); }); }\n"
       "  for (auto& t : th) t.join();\n}\n}  // namespace\n";
  // hess_coord! over all patterns (ora_hess): zero fill is implicit (every slot of every local point is assigned)
  o += "extern \"C\" void cmp_hess(const double* X, const double* Y, const double* TH, double w, double* hess, const unsigned char* const* data, int nthreads, int rank, int world) {\n";
  for (size_t k = 0; k < m.pats.size(); k++) {
    const Pattern& p = m.pats[k];
    if (p.o2step == 0) continue;
    const std::string K = std::to_string(k);
    o += "  {\n    const i64 lo = " + std::to_string(p.nitr) + "LL * rank / world, hi = " + std::to_string(p.nitr) + "LL * (rank + 1) / world;\n";
    o += "    const unsigned char* base = " + (p.databuf >= 0 ? "data[" + std::to_string(p.databuf) + "]" : std::string("nullptr")) + ";\n";
    if (p.kind != KIND_OBJ)
      o += "    if (!Y) { for (i64 q = " + std::to_string(p.o2) + "LL + " + std::to_string(p.o2step) + "LL * lo; q < " + std::to_string(p.o2) + "LL + " +
           std::to_string(p.o2step) + "LL * hi; q++) hess[q] = 0.0; } else\n";
    o += "    run(lo, hi, nthreads, [=](i64 k) { hess_p" + K + "(k, X, Y, TH, base ? base + (size_t)k * " + std::to_string(p.stride) + " : nullptr, w, hess); });\n  }\n";
  }
  o += "}\n";
  return o;
}

}  // namespace

extern "C" {
// returns a malloc'ed NUL-terminated C++ source (caller frees with ora_free)
char* ora_emit_source(void* h) {
  const std::string s = emit_source(*(Model*)h);
  char* out = (char*)malloc(s.size() + 1);
  std::memcpy(out, s.c_str(), s.size() + 1);
  return out;
}
void ora_free(void* p) { free(p); }
}
