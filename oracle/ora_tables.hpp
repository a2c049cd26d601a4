// =============================================================================
// oracle/ora_tables.hpp -- TEST INFRASTRUCTURE ONLY (see exa_oracle.cpp).
//
// The IR enums, Julia-semantics scalar helpers and the derivative tables of the CPU restatement
// (src/functionlist.jl:6-81, ext/functionlist.jl:6-126), shared by the interpreter (exa_oracle.cpp) and by the
// per-pattern straight-line C++ it emits for the compiled CPU baseline (ora_emit_source): with a constant `op`
// the inlined table collapses to the one formula, as Julia's type-specialised code does.
// =============================================================================
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <thread>

typedef int64_t i64;

namespace ora {
namespace {   // internal linkage: every translation unit (interpreter, generated ports) has its own copy

// ---- IR (include/exa_b200.h §IR; emitted by examodels.jl_b200/nlp.py) ---------
enum { T_CONST_I, T_CONST_F, T_DATA_SELF, T_DATA_FIELD, T_VAR, T_PAR, T_NULL, T_OP1, T_OP2, T_VAL };
enum { KIND_OBJ, KIND_CON, KIND_AUG };
enum { FT_I64, FT_F64, FT_I32, FT_F32 };

// univariate op codes: order of src/functionlist.jl:6-60
enum {
  U_PLUS, U_MINUS, U_INV, U_SQRT, U_CBRT, U_ABS, U_ABS2, U_SIGN, U_EXP, U_EXP2, U_EXP10,
  U_EXPM1, U_LOG, U_LOG2, U_LOG1P, U_LOG10, U_SIN, U_COS, U_TAN, U_ASIN, U_ACOS, U_ATAN,
  U_ACOT, U_CSC, U_SEC, U_COT, U_SINH, U_COSH, U_TANH, U_ASINH, U_ACOSH, U_CSCH, U_SECH,
  U_COTH, U_SIND, U_COSD, U_TAND, U_CSCD, U_SECD, U_COTD, U_ATAND, U_ACOTD, U_SINPI,
  U_COSPI, U_SINC, U_DEG2RAD, U_RAD2DEG, U_SIGNBIT, U_FLOOR, U_CEIL, U_ATANH, U_ACOTH,
  // SpecialFunctions extension: order of /root/reference/ext/functionlist.jl:6-104
  U_ERF, U_ERFC, U_ERFI, U_ERFCX, U_DIGAMMA, U_TRIGAMMA, U_INVDIGAMMA, U_GAMMA, U_AIRYAI, U_AIRYBI, U_AIRYAIPRIME,
  U_AIRYBIPRIME, U_BESSELJ0, U_BESSELY0, U_BESSELJ1, U_BESSELY1, U_DAWSON, U_ERFINV, U_ERFCINV, U_COUNT
};
// bivariate op codes: order of src/functionlist.jl:71-81
enum { B_ADD, B_SUB, B_MUL, B_DIV, B_POW, B_ATAN, B_HYPOT, B_MAX, B_MIN,
       B_BETA, B_LOGBETA,   // SpecialFunctions extension: ext/functionlist.jl:111-126
       B_COUNT };

struct IRNode { i64 tag, a, b, payload; };
struct Field { i64 off, type; };

// evaluated-node kinds: Real | (Second)AdjointNull | …NodeVar | …Node1 | …Node2  (graph.jl:106-461)
enum { K_REAL, K_NULL, K_VAR, K_N1, K_N2 };
// Node1 flavour: plain unary f, or a bivariate f with one Real operand (register.jl:231-266)
enum { FX_NONE, FX_FIRST, FX_SECOND };

struct TNode {          // one node of the expanded expression TREE
  int tag, op, c1, c2;  // c1/c2: tree children (value children; VAR/PAR: index-expression child)
  int ir;               // originating IR node (identity for the === probe)
  i64 ipay; double fpay;
  int kind, fx;         // static: evaluated kind and FirstFixed/SecondFixed flavour
};

struct Pattern {
  int kind; i64 nitr; int itr_kind; i64 range_start; int databuf; i64 stride;
  std::vector<Field> fields;
  i64 o0, o1, o2; int base; std::vector<int> idx_roots_ir; std::vector<i64> dims;
  std::vector<IRNode> ir; int root_ir;
  std::vector<TNode> t; int root; std::vector<int> idx_roots;
  std::vector<int> comp1, comp2; int o1step, o2step;
  const unsigned char* data;
};

struct Model {
  i64 nvar, npar, ncon, nobj, nconaug, nnzg, nnzj, nnzh;
  std::vector<Pattern> pats;
  std::vector<double> theta;
  // KA-extension style scratch (ext/ExaModelsKernelAbstractions.jl:21-31,39-53)
  std::vector<std::pair<i64, i64>> gsparsity; std::vector<i64> gptr;
  std::string err;
  int nthreads;
  int rank = 0, world = 1;   // test aid: evaluate only shard `rank` of `world` of every iterator
};

// ---- scalar helpers (Julia Base semantics used by src/functionlist.jl) --------------
inline double sq(double x) { return x * x; }              // literal x^2 == x*x
inline double cube(double x) { return x * x * x; }        // literal x^3 == x*x*x
const double PI = 3.14159265358979323846;
inline double jl_powi(double x, i64 n) {                  // Base.^(::Float64, ::Integer)
  if (n == 0) return 1.0;
  if (n == 1) return x;
  if (n == 2) return x * x;
  if (n == 3) return x * x * x;
  if (n == -1) return 1.0 / x;
  if (n == -2) { double r = 1.0 / x; return r * r; }
  return std::pow(x, (double)n);
}
inline double jl_sec(double x) { return 1.0 / std::cos(x); }
inline double jl_csc(double x) { return 1.0 / std::sin(x); }
inline double jl_cot(double x) { return 1.0 / std::tan(x); }
inline double jl_sech(double x) { return 1.0 / std::cosh(x); }
inline double jl_csch(double x) { return 1.0 / std::sinh(x); }
inline double jl_coth(double x) { return 1.0 / std::tanh(x); }
inline double jl_deg2rad(double x) { return x * (PI / 180.0); }
inline double jl_rad2deg(double x) { return x * (180.0 / PI); }
inline double jl_sinpi(double x) {
  double r = std::fmod(x, 2.0);                           // exact
  if (r > 1.0) r -= 2.0; else if (r < -1.0) r += 2.0;     // r in [-1,1]
  if (r > 0.5) r = 1.0 - r; else if (r < -0.5) r = -1.0 - r;
  return std::sin(PI * r);
}
inline double jl_cospi(double x) {
  double r = std::fabs(std::fmod(x, 2.0));                // [0,2)
  if (r > 1.0) r = 2.0 - r;                               // [0,1]
  if (r == 0.5) return 0.0;
  return r > 0.5 ? -std::cos(PI * (1.0 - r)) : std::cos(PI * r);
}
inline double jl_sind(double x) { return jl_sinpi(x / 180.0); }
inline double jl_cosd(double x) { return jl_cospi(x / 180.0); }
inline double jl_tand(double x) { return jl_sind(x) / jl_cosd(x); }
inline double jl_cscd(double x) { return 1.0 / jl_sind(x); }
inline double jl_secd(double x) { return 1.0 / jl_cosd(x); }
inline double jl_cotd(double x) { return 1.0 / jl_tand(x); }
inline double jl_sinc(double x) { return x == 0.0 ? 1.0 : jl_sinpi(x) / (PI * x); }
inline double jl_sign(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : x); }
const double LOG2 = 0.69314718055994530942, LOG10 = 2.30258509299404568402;
const double D2R = PI / 180.0;


// ---- SpecialFunctions extension (ext/functionlist.jl; values from SpecialFunctions.jl / openspecfun, a third-party
// dependency that is not vendored: published algorithms restated in extended precision, independently of the device
// code in examodels.jl_b200/csrc/exb_special.h, and pinned against scipy / mpmath in tests/test_special_functions.py) ----
typedef long double ld;
const ld PIl = 3.14159265358979323846264338327950288L;
inline ld sinpil_(ld x) { ld r = fmodl(x, 2.0L); return sinl(PIl * r); }
inline ld cospil_(ld x) { ld r = fmodl(x, 2.0L); return cosl(PIl * r); }
// psi^(n)(x), n = 0..3, x > 0: recurrence up to x >= 30, then the asymptotic series with Bernoulli numbers B_2..B_20
ld polygamma_pos(int n, ld x) {
  static const ld B[10] = {1.0L / 6, -1.0L / 30, 1.0L / 42, -1.0L / 30, 5.0L / 66, -691.0L / 2730, 7.0L / 6, -3617.0L / 510, 43867.0L / 798, -174611.0L / 330};
  ld r = 0.0L;
  const ld fact[4] = {1.0L, 1.0L, 2.0L, 6.0L};
  while (x < 30.0L) {
    if (n == 0) r -= 1.0L / x; else r += ((n & 1) ? 1.0L : -1.0L) * fact[n] / powl(x, n + 1);
    x += 1.0L;
  }
  ld s;
  if (n == 0) {
    s = logl(x) - 0.5L / x;
    for (int k = 1; k <= 10; k++) s -= B[k - 1] / (2.0L * k * powl(x, 2 * k));
  } else {
    // (-1)^(n+1) [ (n-1)!/x^n + n!/(2 x^(n+1)) + sum_k B_2k (2k+n-1)!/((2k)! x^(2k+n)) ]
    s = fact[n - 1] / powl(x, n) + fact[n] / (2.0L * powl(x, n + 1));
    for (int k = 1; k <= 10; k++) {
      ld c = 1.0L;                                   // (2k+n-1)! / (2k)!
      for (int j = 2 * k + 1; j <= 2 * k + n - 1; j++) c *= j;
      s += B[k - 1] * c / powl(x, 2 * k + n);
    }
    if (!(n & 1)) s = -s;
  }
  return r + s;
}
ld polygamma_l(int n, ld x) {
  if (x > 0.0L) return polygamma_pos(n, x);
  const ld sn = sinpil_(x), cs = cospil_(x), ct = cs / sn, c2 = 1.0L / (sn * sn);   // reflection about 1 - x
  switch (n) {
    case 0: return polygamma_pos(0, 1.0L - x) - PIl * ct;
    case 1: return -polygamma_pos(1, 1.0L - x) + PIl * PIl * c2;
    case 2: return polygamma_pos(2, 1.0L - x) - 2.0L * PIl * PIl * PIl * ct * c2;
    default: return -polygamma_pos(3, 1.0L - x) + 2.0L * PIl * PIl * PIl * PIl * c2 * (2.0L * ct * ct + c2);
  }
}
double sf_digamma(double x) { return (double)polygamma_l(0, x); }
double sf_trigamma(double x) { return (double)polygamma_l(1, x); }
double sf_polygamma(int n, double x) { return (double)polygamma_l(n, x); }
double sf_invdigamma(double y) {   // Minka's iteration (SpecialFunctions.jl invdigamma)
  ld xo = y >= -2.22 ? expl((ld)y) + 0.5L : -1.0L / ((ld)y + 0.57721566490153286060651209L);
  for (int it = 0; it < 40; it++) {
    ld xn = xo - (polygamma_l(0, xo) - (ld)y) / polygamma_l(1, xo);
    bool done = fabsl(xn - xo) <= 1e-17L * fabsl(xn);
    xo = xn;
    if (done) break;
  }
  return (double)xo;
}
// (sqrt(pi)/2) erfi(x) = sum x^(2k+1)/(k! (2k+1)); Dawson D(x) = e^{-x^2} * that; asymptotic series for |x| >= 8
ld erfi_series_l(ld x) {
  ld t = x, s = x;
  for (int k = 1; k < 400; k++) { t *= x * x / k; ld a = t / (2 * k + 1); s += a; if (fabsl(a) < 1e-22L * fabsl(s)) break; }
  return s;
}
ld dawson_asym_l(ld x) {
  ld q = 1.0L / (2.0L * x * x), t = 1.0L, s = 1.0L;
  for (int k = 1; k < 200; k++) { ld tn = t * (2 * k - 1) * q; if (fabsl(tn) >= fabsl(t) || fabsl(tn) < 1e-24L) break; t = tn; s += t; }
  return s / (2.0L * x);
}
double sf_dawson(double x) { return fabs(x) < 8.0 ? (double)(expl(-(ld)x * x) * erfi_series_l(x)) : (double)dawson_asym_l(x); }
double sf_erfi(double x) {
  const ld c = 2.0L / sqrtl(PIl);
  return fabs(x) < 8.0 ? (double)(c * erfi_series_l(x)) : (double)(c * expl((ld)x * x) * dawson_asym_l(x));
}
double sf_erfcx(double x) {
  if (x < 25.0) return (double)(expl((ld)x * x) * erfcl((ld)x));
  ld q = 1.0L / (2.0L * (ld)x * x), t = 1.0L, s = 1.0L;   // 1/(x sqrt(pi)) sum (-1)^k (2k-1)!!/(2x^2)^k
  for (int k = 1; k < 60; k++) { ld tn = -t * (2 * k - 1) * q; if (fabsl(tn) >= fabsl(t)) break; t = tn; s += t; }
  return (double)(s / ((ld)x * sqrtl(PIl)));
}
double sf_erfinv(double y) {       // Newton on erf in extended precision from a rational start (Winitzki)
  if (!(y > -1.0 && y < 1.0)) return y == 1.0 ? INFINITY : y == -1.0 ? -INFINITY : NAN;
  const ld a = 0.147L, l = logl(1.0L - (ld)y * y), t = 2.0L / (PIl * a) + l / 2.0L;
  ld x = sqrtl(sqrtl(t * t - l / a) - t); if (y < 0) x = -x;
  for (int it = 0; it < 60; it++) { ld dx = (erfl(x) - (ld)y) / (2.0L / sqrtl(PIl) * expl(-x * x)); x -= dx; if (fabsl(dx) <= 1e-19L * fabsl(x)) break; }
  return (double)x;
}
double sf_erfcinv(double y) {      // Newton on erfc (keeps relative accuracy for small y)
  if (!(y > 0.0 && y < 2.0)) return y == 0.0 ? INFINITY : y == 2.0 ? -INFINITY : NAN;
  ld x = y >= 0.25 && y <= 1.75 ? (ld)sf_erfinv(1.0 - y) : (y < 1 ? 1.0L : -1.0L) * sqrtl(-logl((y < 1 ? (ld)y : 2.0L - (ld)y)));
  for (int it = 0; it < 80; it++) { ld dx = (erfcl(x) - (ld)y) / (-2.0L / sqrtl(PIl) * expl(-x * x)); x -= dx; if (fabsl(dx) <= 1e-19L * fabsl(x)) break; }
  return (double)x;
}
// Airy: Maclaurin series in QUAD precision (__float128: the series cancels ~e^{2 zeta} for x > 0) for |x| <= 9 --
// Ai = c1 f - c2 g, Bi = sqrt(3)(c1 f + c2 g) -- else the Poincare asymptotic series (its truncation error e^{-2 zeta} is
// below 1e-15 there).  Independent of the device path's tabulated Taylor expansion.
void airy_l(ld x, ld& ai, ld& aip, ld& bi, ld& bip) {
  if (fabsl(x) <= 9.0L) {
    typedef __float128 qd;
    const qd c1 = 0.355028053887817239260063186004183176Q, c2 = 0.258819403792806798405183560189203963Q,
             s3 = 1.732050807568877293527446341505872367Q;
    // f = sum a_k, a_0 = 1, a_k = a_{k-1} x^3 / ((3k-1)(3k)); g = sum b_k, b_0 = x, b_k = b_{k-1} x^3 / ((3k)(3k+1));
    // f' = sum 3k a_k / x, g' = sum (3k+1) b_k / x  (accumulated as series in x^2 to stay finite at x = 0)
    const qd xq = x, x3 = xq * xq * xq;
    qd a = 1.0Q, b = xq, f = 1.0Q, g = xq, ap = 0.0Q, bp = 1.0Q, fp = 0.0Q, gp = 1.0Q;   // ap = a_k' , bp = b_k'
    for (int k = 1; k < 400; k++) {
      // a_k' = a_{k-1}' x^3/((3k-1)3k) * (3k)/(3k-3)  for k >= 2; a_1' = x^2/2
      ap = k == 1 ? xq * xq / 2.0Q : ap * x3 / ((3.0Q * k - 1) * (3.0Q * k - 3));
      bp = bp * x3 / ((3.0Q * k) * (3.0Q * k - 2));          // b_k' = x^(3k) / prod: (3k+1) b_k / x
      a *= x3 / ((3.0Q * k - 1) * (3.0Q * k)); b *= x3 / ((3.0Q * k) * (3.0Q * k + 1));
      f += a; g += b; fp += ap; gp += bp;
      const qd ta = a < 0 ? -a : a, tb = b < 0 ? -b : b, tf = f < 0 ? -f : f, tg = g < 0 ? -g : g;
      if (k > 3 && ta <= 1e-40Q * (tf + 1e-300Q) && tb <= 1e-40Q * (tg + 1e-300Q)) break;
    }
    ai = (ld)(c1 * f - c2 * g); aip = (ld)(c1 * fp - c2 * gp); bi = (ld)(s3 * (c1 * f + c2 * g)); bip = (ld)(s3 * (c1 * fp + c2 * gp));
    return;
  }
  const ld z = fabsl(x), z14 = sqrtl(sqrtl(z)), zeta = 2.0L / 3.0L * z * sqrtl(z), sp = sqrtl(PIl);
  ld u[16], v[16]; u[0] = v[0] = 1.0L;
  for (int k = 1; k < 16; k++) { u[k] = u[k - 1] * (6.0L * k - 5) * (6.0L * k - 3) * (6.0L * k - 1) / ((2.0L * k - 1) * 216.0L * k); v[k] = u[k] * (6.0L * k + 1) / (1.0L - 6.0L * k); }
  // sums truncated at the smallest term
  if (x > 0) {
    ld sa = 0, sap = 0, sb = 0, sbp = 0, p = 1.0L, sg = 1.0L, last = 1e300L;
    for (int k = 0; k < 16; k++) { ld t = u[k] * p; if (fabsl(t) > last) break; last = fabsl(t); sa += sg * t; sb += t; sap += sg * v[k] * p; sbp += v[k] * p; p /= zeta; sg = -sg; }
    ai = expl(-zeta) / (2.0L * sp * z14) * sa; aip = -z14 * expl(-zeta) / (2.0L * sp) * sap;
    bi = expl(zeta) / (sp * z14) * sb; bip = z14 * expl(zeta) / sp * sbp;
  } else {
    ld pe = 0, po = 0, qe = 0, qo = 0, p = 1.0L, last = 1e300L;
    for (int k = 0; k + 1 < 16; k += 2) {
      ld t = u[k] * p; if (fabsl(t) > last) break; last = fabsl(t);
      ld sg = (k & 2) ? -1.0L : 1.0L;
      pe += sg * u[k] * p; qe += sg * v[k] * p; p /= zeta;
      po += sg * u[k + 1] * p; qo += sg * v[k + 1] * p; p /= zeta;
    }
    ld th = zeta - PIl / 4.0L, c = cosl(th), sn = sinl(th);
    ai = (c * pe + sn * po) / (sp * z14); aip = z14 / sp * (sn * qe - c * qo);
    bi = (-sn * pe + c * po) / (sp * z14); bip = z14 / sp * (c * qe + sn * qo);
  }
}
double sf_airy(double x, int which) { ld a, ap, b, bp; airy_l(x, a, ap, b, bp); return (double)(which == 0 ? a : which == 1 ? ap : which == 2 ? b : bp); }
double sf_logbeta(double a, double b) { return (double)(lgammal(a) + lgammal(b) - lgammal((ld)a + b)); }
double sf_beta(double a, double b) {
  if (a > 0 && b > 0) return (double)expl(lgammal(a) + lgammal(b) - lgammal((ld)a + b));
  return (double)(tgammal(a) * tgammal(b) / tgammal((ld)a + b));
}
const double INVSQRTPI = 0.56418958354775628695, SQRTPIHALF = 0.88622692545275801365;   // _cinvsqrtpi, _csqrtpihalf (ext/ExaModelsSpecialFunctions.jl:6-8)

// ---- univariate table: f, f', f''  (src/functionlist.jl:6-60, formulas kept literally) -----
#define ORA_INLINE __attribute__((always_inline)) inline
ORA_INLINE void uni(int op, double x, double& f, double& d, double& dd, int order) {
  switch (op) {
    case U_PLUS: f = x; d = 1.0; dd = 0.0; break;                                   // :7
    case U_MINUS: f = -x; d = -1.0; dd = 0.0; break;                                // :8
    case U_INV: f = 1.0 / x; if (order) { d = -1.0 / sq(x); dd = 2.0 / cube(x); } break;  // :9
    case U_SQRT: { double s = std::sqrt(x); f = s;
      if (order) { d = 1.0 / (2.0 * s); dd = -1.0 / (4.0 * cube(s)); } } break;     // :10
    case U_CBRT: { double c = std::cbrt(x); f = c;
      if (order) { d = 1.0 / (3.0 * sq(c)); dd = -2.0 / (9.0 * std::pow(c, 5.0)); } } break; // :11
    case U_ABS: f = std::fabs(x); d = std::signbit(x) ? -1.0 : 1.0; dd = 0.0; break;  // :12
    case U_ABS2: f = x * x; d = 2.0 * x; dd = 2.0; break;                           // :13
    case U_SIGN: f = jl_sign(x); d = 0.0; dd = 0.0; break;                          // :14
    case U_EXP: f = std::exp(x); d = f; dd = f; break;                              // :15
    case U_EXP2: f = std::exp2(x); d = LOG2 * f; dd = sq(LOG2) * f; break;          // :16
    case U_EXP10: f = std::pow(10.0, x); d = LOG10 * f; dd = sq(LOG10) * f; break;  // :17
    case U_EXPM1: f = std::expm1(x); if (order) { d = std::exp(x); dd = d; } break; // :18
    case U_LOG: f = std::log(x); if (order) { d = 1.0 / x; dd = -1.0 / sq(x); } break; // :19
    case U_LOG2: f = std::log2(x);
      if (order) { d = 1.0 / (LOG2 * x); dd = -LOG2 / (sq(LOG2) * sq(x)); } break;  // :20
    case U_LOG1P: f = std::log1p(x);
      if (order) { d = 1.0 / (1.0 + x); dd = -1.0 / sq(1.0 + x); } break;           // :21
    case U_LOG10: f = std::log10(x);
      if (order) { d = 1.0 / (LOG10 * x); dd = -LOG10 / (sq(LOG10) * sq(x)); } break; // :22
    case U_SIN: f = std::sin(x); if (order) { d = std::cos(x); dd = -f; } break;    // :23
    case U_COS: f = std::cos(x); if (order) { d = -std::sin(x); dd = -f; } break;   // :24
    case U_TAN: f = std::tan(x);
      if (order) { double s2 = sq(jl_sec(x)); d = s2; dd = 2.0 * s2 * f; } break;   // :25
    case U_ASIN: f = std::asin(x);
      if (order) { double q = 1.0 - sq(x); d = 1.0 / std::sqrt(q); dd = x / (q * std::sqrt(q)); } break; // :26
    case U_ACOS: f = std::acos(x);
      if (order) { double q = 1.0 - sq(x); d = -1.0 / std::sqrt(q); dd = (-x) / (q * std::sqrt(q)); } break; // :27
    case U_ATAN: f = std::atan(x);
      if (order) { double q = 1.0 + sq(x); d = 1.0 / q; dd = (-2.0 * x) / sq(q); } break; // :28
    case U_ACOT: f = std::atan(1.0 / x);
      if (order) { double q = 1.0 + sq(x); d = -1.0 / q; dd = (2.0 * x) / sq(q); } break; // :29
    case U_CSC: { double c = jl_csc(x); f = c;
      if (order) { double ct = jl_cot(x); d = -ct * c; dd = -(-1.0 - sq(ct)) * c + sq(ct) * c; } } break; // :30
    case U_SEC: { double s = jl_sec(x); f = s;
      if (order) { double t = std::tan(x); d = s * t; dd = cube(s) + s * sq(t); } } break; // :31
    case U_COT: { double ct = jl_cot(x); f = ct;
      if (order) { d = -1.0 - sq(ct); dd = -2.0 * ct * (-1.0 - sq(ct)); } } break;  // :32
    case U_SINH: f = std::sinh(x); if (order) { d = std::cosh(x); dd = f; } break;  // :33
    case U_COSH: f = std::cosh(x); if (order) { d = std::sinh(x); dd = f; } break;  // :34
    case U_TANH: f = std::tanh(x);
      if (order) { d = 1.0 - sq(f); dd = -2.0 * f * (1.0 - sq(f)); } break;         // :35
    case U_ASINH: f = std::asinh(x);
      if (order) { double q = 1.0 + sq(x); d = 1.0 / std::sqrt(q); dd = (-x) / (q * std::sqrt(q)); } break; // :36
    case U_ACOSH: f = std::acosh(x);
      if (order) { double q = -1.0 + sq(x); d = 1.0 / std::sqrt(q); dd = (-x) / (q * std::sqrt(q)); } break; // :37
    case U_CSCH: { double c = jl_csch(x); f = c;
      if (order) { double ct = jl_coth(x); d = -c * ct; dd = cube(c) + c * sq(ct); } } break; // :38
    case U_SECH: { double s = jl_sech(x); f = s;
      if (order) { double t = std::tanh(x); d = -t * s; dd = -(1.0 - sq(t)) * s + sq(t) * s; } } break; // :39
    case U_COTH: { double ct = jl_coth(x); f = ct;
      if (order) { double c = jl_csch(x); d = -sq(c); dd = 2.0 * sq(c) * ct; } } break; // :40
    case U_SIND: f = jl_sind(x);
      if (order) { d = jl_deg2rad(jl_cosd(x)); dd = -D2R * jl_deg2rad(f); } break;  // :41
    case U_COSD: f = jl_cosd(x);
      if (order) { d = -jl_deg2rad(jl_sind(x)); dd = -D2R * jl_deg2rad(f); } break; // :42
    case U_TAND: f = jl_tand(x);
      if (order) { double q = jl_deg2rad(1.0 + sq(f)); d = q; dd = (2.0 * D2R) * f * q; } break; // :43
    case U_CSCD: { double c = jl_cscd(x); f = c;
      if (order) { double ct = jl_cotd(x); double a = -jl_deg2rad(c * ct); d = a;
        dd = -D2R * (a * ct - c * jl_deg2rad(1.0 + sq(ct))); } } break;             // :44
    case U_SECD: { double s = jl_secd(x); f = s;
      if (order) { double t = jl_tand(x); double a = jl_deg2rad(t * s); d = a;
        dd = D2R * (a * t + jl_deg2rad(1.0 + sq(t)) * s); } } break;                // :45
    case U_COTD: { double ct = jl_cotd(x); f = ct;
      if (order) { double q = jl_deg2rad(1.0 + sq(ct)); d = -q; dd = (2.0 * D2R) * ct * q; } } break; // :46
    case U_ATAND: f = jl_rad2deg(std::atan(x));
      if (order) { double q = jl_deg2rad(1.0 + sq(x)); d = 1.0 / q; dd = (-(2.0 * D2R) * x) / sq(q); } break; // :47
    case U_ACOTD: f = jl_rad2deg(std::atan(1.0 / x));
      if (order) { double q = jl_deg2rad(1.0 + sq(x)); d = -1.0 / q; dd = ((2.0 * D2R) * x) / sq(q); } break; // :48
    case U_SINPI: f = jl_sinpi(x);
      if (order) { d = PI * jl_cospi(x); dd = -sq(PI) * f; } break;                 // :49
    case U_COSPI: f = jl_cospi(x);
      if (order) { d = -PI * jl_sinpi(x); dd = -sq(PI) * f; } break;                // :50
    case U_SINC: f = jl_sinc(x);
      if (order) { double s = jl_sinpi(x), c = jl_cospi(x);
        d = (-s + PI * x * c) / (PI * sq(x));
        dd = ((2.0 * sq(PI)) * s - (2.0 * cube(PI)) * x * c - std::pow(PI, 4.0) * sq(x) * s) / (cube(PI) * cube(x)); } break; // :51
    case U_DEG2RAD: f = jl_deg2rad(x); d = D2R; dd = 0.0; break;                    // :52
    case U_RAD2DEG: f = jl_rad2deg(x); d = 180.0 / PI; dd = 0.0; break;             // :53
    case U_SIGNBIT: f = std::signbit(x) ? 1.0 : 0.0; d = 0.0; dd = 0.0; break;      // :54
    case U_FLOOR: f = std::floor(x); d = 0.0; dd = 0.0; break;                      // :55
    case U_CEIL: f = std::ceil(x); d = 0.0; dd = 0.0; break;                        // :56
    case U_ATANH: f = std::atanh(x);
      if (order) { if (std::fabs(x) > 1.0) { d = NAN; dd = NAN; }
        else { double iv = 1.0 / (1.0 - sq(x)); d = iv; dd = (-sq(iv)) * (-2.0 * x); } } break; // :58
    case U_ACOTH: f = std::atanh(1.0 / x);
      if (order) { if (std::fabs(x) < 1.0) { d = NAN; dd = NAN; }
        else { double iv = 1.0 / (1.0 - sq(x)); d = iv; dd = (-sq(iv)) * (-2.0 * x); } } break; // :59
    // ---- SpecialFunctions extension: ext/functionlist.jl:6-104, formulas kept literally ----
    case U_ERF: f = std::erf(x); if (order) { d = (2 * INVSQRTPI) * std::exp(-sq(x)); dd = -(4 * INVSQRTPI) * x * std::exp(-sq(x)); } break;       // ext:6-10
    case U_ERFC: f = std::erfc(x); if (order) { d = -(2 * INVSQRTPI) * std::exp(-sq(x)); dd = (4 * INVSQRTPI) * x * std::exp(-sq(x)); } break;    // ext:11-15
    case U_ERFI: f = sf_erfi(x); if (order) { d = (2 * INVSQRTPI) * std::exp(sq(x)); dd = (4 * INVSQRTPI) * x * std::exp(sq(x)); } break;          // ext:16-20
    case U_ERFCX: f = sf_erfcx(x); if (order) { d = 2 * (-INVSQRTPI + x * f); dd = 2 * (f + 2 * x * (-INVSQRTPI + x * f)); } break;                // ext:21-25
    case U_DIGAMMA: f = sf_digamma(x); if (order) { d = sf_trigamma(x); dd = sf_polygamma(2, x); } break;                                          // ext:26-30
    case U_TRIGAMMA: f = sf_trigamma(x); if (order) { d = sf_polygamma(2, x); dd = sf_polygamma(3, x); } break;                                    // ext:31-35
    case U_INVDIGAMMA: f = sf_invdigamma(x);
      if (order) { d = 1 / sf_trigamma(f); dd = (-sf_polygamma(2, f)) / cube(sf_trigamma(f)); } break;                                             // ext:36-40
    case U_GAMMA: f = std::tgamma(x);
      if (order) { d = f * sf_digamma(x); dd = f * (sf_trigamma(x) + sq(sf_digamma(x))); } break;                                                  // ext:41-45
    case U_AIRYAI: f = sf_airy(x, 0); if (order) { d = sf_airy(x, 1); dd = x * f; } break;                                                         // ext:46-50
    case U_AIRYBI: f = sf_airy(x, 2); if (order) { d = sf_airy(x, 3); dd = x * f; } break;                                                         // ext:51-55
    case U_AIRYAIPRIME: f = sf_airy(x, 1); if (order) { d = x * sf_airy(x, 0); dd = sf_airy(x, 0) + x * f; } break;                                // ext:56-60
    case U_AIRYBIPRIME: f = sf_airy(x, 3); if (order) { d = x * sf_airy(x, 2); dd = sf_airy(x, 2) + x * f; } break;                                // ext:61-65
    case U_BESSELJ0: f = ::j0(x); if (order) { d = -::j1(x); dd = (-f + ::jn(2, x)) / 2; } break;                                                  // ext:66-70
    case U_BESSELY0: f = ::y0(x); if (order) { d = -::y1(x); dd = (-f + ::yn(2, x)) / 2; } break;                                                  // ext:71-75
    case U_BESSELJ1: f = ::j1(x); if (order) { d = (::j0(x) - ::jn(2, x)) / 2; dd = ((-::jn(1, x) + ::jn(3, x)) / 2 - f) / 2; } break;             // ext:76-80
    case U_BESSELY1: f = ::y1(x); if (order) { d = (::y0(x) - ::yn(2, x)) / 2; dd = ((::yn(3, x) - ::yn(1, x)) / 2 - f) / 2; } break;              // ext:81-85
    case U_DAWSON: f = sf_dawson(x); if (order) { d = 1 - 2 * x * f; dd = -2 * f - 2 * x * (1 - 2 * x * f); } break;                               // ext:86-90
    case U_ERFINV: f = sf_erfinv(x);
      if (order) { d = SQRTPIHALF * std::exp(sq(f)); dd = SQRTPIHALF * std::exp(sq(f)) * 2 * f * SQRTPIHALF * std::exp(sq(f)); } break;            // ext:93-97
    case U_ERFCINV: f = sf_erfcinv(x);
      if (order) { d = -SQRTPIHALF * std::exp(sq(f)); dd = (PI / 2) * f * std::exp(2 * sq(f)); } break;                                            // ext:98-102
    default: f = d = dd = NAN;
  }
}

// A Real operand: Int or Float64 (Julia keeps Int arithmetic among Ints).
struct Real { bool is_int; i64 i; double f; double val() const { return is_int ? (double)i : f; } };

inline double bpow(double x1, const Real& e) {            // x1 ^ x2 with x2 Int or Float64
  return e.is_int ? jl_powi(x1, e.i) : std::pow(x1, e.f);
}
inline Real radd(const Real& e, i64 k) {                  // (k + x2) in the exponent formulas
  Real r = e; if (e.is_int) r.i = e.i + k; else r.f = e.f + (double)k; return r;
}

// ---- bivariate table (src/functionlist.jl:71-81) ---------------------------------------
// Computes f and whichever partials `need` asks for: bit0 first-order, bit1 second-order.
// `e2` carries the Int/Float nature of the second operand (matters for ^ only).
struct Bi { double f, y1, y2, h11, h12, h22; };
ORA_INLINE void bi(int op, double x1, double x2, const Real& e1, const Real& e2, Bi& r, bool want1, bool want2) {
  (void)e1;
  switch (op) {
    case B_ADD: r.f = x1 + x2; r.y1 = 1.0; r.y2 = 1.0; r.h11 = r.h12 = r.h22 = 0.0; break;   // :72
    case B_SUB: r.f = x1 - x2; r.y1 = 1.0; r.y2 = -1.0; r.h11 = r.h12 = r.h22 = 0.0; break;  // :73
    case B_MUL: r.f = x1 * x2; r.y1 = x2; r.y2 = x1; r.h11 = 0.0; r.h12 = 1.0; r.h22 = 0.0; break; // :74
    case B_DIV: r.f = x1 / x2; r.y1 = 1.0 / x2; r.y2 = (-x1) / sq(x2);
      r.h11 = 0.0; r.h12 = -1.0 / sq(x2); r.h22 = (2.0 * x1) / cube(x2); break;              // :75
    case B_POW:                                                                              // :76
      r.f = bpow(x1, e2);
      if (want1) {
        double pm1 = bpow(x1, radd(e2, -1));
        r.y1 = x2 * pm1;
        r.h11 = (-1.0 + x2) * x2 * bpow(x1, radd(e2, -2));
        if (e2.is_int) r.h11 = (double)((-1 + e2.i) * e2.i) * bpow(x1, radd(e2, -2));
      }
      if (want2) {
        double lg = std::log(x1), pm1 = bpow(x1, radd(e2, -1));
        r.y2 = lg * r.f;
        r.h12 = pm1 + x2 * pm1 * lg;
        r.h22 = sq(lg) * r.f;
      }
      break;
    case B_ATAN: { double q = sq(x1) + sq(x2); r.f = std::atan2(x1, x2);                     // :77
      r.y1 = x2 / q; r.y2 = (-x1) / q; r.h11 = (-2.0 * x1 * x2) / sq(q);
      r.h12 = (sq(x1) - sq(x2)) / (std::pow(x1, 4.0) + 2.0 * sq(x1) * sq(x2) + std::pow(x2, 4.0));
      r.h22 = (2.0 * x1 * x2) / sq(q); } break;
    case B_HYPOT: { double h = std::hypot(x1, x2); r.f = h;                                  // :78
      r.y1 = x1 / h; r.y2 = x2 / h; r.h11 = (-sq(x1) + sq(h)) / cube(h);
      r.h12 = (-x1 * x2) / cube(h); r.h22 = (-sq(x2) + sq(h)) / cube(h); } break;
    case B_MAX: r.f = (x1 < x2 || std::isnan(x2)) ? x2 : x1;                                 // :79
      r.y1 = x1 > x2 ? 1.0 : 0.0; r.y2 = x1 > x2 ? 0.0 : 1.0; r.h11 = r.h12 = r.h22 = 0.0; break;
    case B_MIN: r.f = (x2 < x1 || std::isnan(x2)) ? x2 : x1;                                 // :80
      r.y1 = x1 < x2 ? 1.0 : 0.0; r.y2 = x1 < x2 ? 0.0 : 1.0; r.h11 = r.h12 = r.h22 = 0.0; break;
    case B_BETA: { r.f = sf_beta(x1, x2);                                                          // ext:111-118
      double p1 = sf_digamma(x1), p2 = sf_digamma(x2), p12 = sf_digamma(x1 + x2), t1 = sf_trigamma(x1), t2 = sf_trigamma(x2), t12 = sf_trigamma(x1 + x2);
      r.y1 = r.f * (p1 - p12); r.y2 = r.f * (-p12 + p2);
      r.h11 = r.f * (t1 - t12 + sq(p1 - p12)); r.h12 = -r.f * t12 + r.f * (p1 - p12) * (-p12 + p2); r.h22 = r.f * (-t12 + t2 + sq(-p12 + p2)); } break;
    case B_LOGBETA: { r.f = sf_logbeta(x1, x2);                                                    // ext:119-126
      double p12 = sf_digamma(x1 + x2), t12 = sf_trigamma(x1 + x2);
      r.y1 = sf_digamma(x1) - p12; r.y2 = -p12 + sf_digamma(x2);
      r.h11 = sf_trigamma(x1) - t12; r.h12 = -t12; r.h22 = -t12 + sf_trigamma(x2); } break;
    default: r.f = NAN;
  }
}

// Real OP Real (both operands free of variables): plain Julia arithmetic.
ORA_INLINE Real real_op2(int op, const Real& a, const Real& b) {
  Real r; r.is_int = false; r.i = 0; r.f = 0;
  if (a.is_int && b.is_int && (op == B_ADD || op == B_SUB || op == B_MUL || op == B_MAX || op == B_MIN)) {
    r.is_int = true;
    switch (op) {
      case B_ADD: r.i = a.i + b.i; break;
      case B_SUB: r.i = a.i - b.i; break;
      case B_MUL: r.i = a.i * b.i; break;
      case B_MAX: r.i = std::max(a.i, b.i); break;
      default: r.i = std::min(a.i, b.i);
    }
    return r;
  }
  if (a.is_int && b.is_int && op == B_POW && b.i >= 0) {   // Int ^ Int (power_by_squaring)
    r.is_int = true; r.i = 1; for (i64 k = 0; k < b.i; k++) r.i *= a.i; return r;
  }
  Bi t; bi(op, a.val(), b.val(), a, b, t, false, false); r.f = t.f; return r;
}
ORA_INLINE Real real_op1(int op, const Real& a) {
  Real r; r.is_int = false; r.i = 0;
  if (a.is_int && (op == U_PLUS || op == U_MINUS || op == U_ABS || op == U_ABS2)) {
    r.is_int = true;
    r.i = op == U_PLUS ? a.i : op == U_MINUS ? -a.i : op == U_ABS ? (a.i < 0 ? -a.i : a.i) : a.i * a.i;
    return r;
  }
  double f, d, dd; uni(op, a.val(), f, d, dd, 0); r.f = f; return r;
}

}  // namespace
}  // namespace ora
