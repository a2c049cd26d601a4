// Pattern IR: parser for the int64 word stream documented in include/exa_b200.h §IR.
// Host-only C++; no CUDA.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace exb {

typedef int64_t i64;

enum { T_CONST_I, T_CONST_F, T_DATA_SELF, T_DATA_FIELD, T_VAR, T_PAR, T_NULL, T_OP1, T_OP2, T_VAL };
enum { KIND_OBJ, KIND_CON, KIND_AUG };
enum { ITR_RANGE, ITR_AOS };
enum { FT_I64, FT_F64, FT_I32, FT_F32 };

// univariate op codes: order of /root/reference/src/functionlist.jl:6-60
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
// bivariate op codes: order of /root/reference/src/functionlist.jl:71-81
enum { B_ADD, B_SUB, B_MUL, B_DIV, B_POW, B_ATAN, B_HYPOT, B_MAX, B_MIN,
       B_BETA, B_LOGBETA,   // SpecialFunctions extension: ext/functionlist.jl:111-126
       B_COUNT };

struct IRNode { i64 tag, a, b, payload; };
struct Field { i64 off, type; };

struct PatternIR {
  int kind = 0; i64 nitr = 0; int itr_kind = 0; i64 range_start = 0; int databuf = -1; i64 stride = 0;
  std::vector<Field> fields;
  i64 o0 = -1, o1 = -1, o2 = -1; int base = -1;
  std::vector<int> idx_roots; std::vector<i64> dims;
  std::vector<IRNode> nodes; int root = -1;
  std::vector<i64> comp1_given, comp2_given;
  // data hints (not part of the IR; filled from the iterator data by detect_iota when the host passes it): integer field f of an
  // AoS iterator holds iota0[f] + k at point k (k = 0, 1, ...) -- the `i` column of an array of NamedTuples built from 1:n
  std::vector<char> iota; std::vector<i64> iota0;
  bool has_iota() const { for (char c : iota) if (c) return true; return false; }
};

struct ModelIR {
  i64 nvar = 0, npar = 0; int ndatabufs = 0;
  std::vector<PatternIR> pats;
};

inline bool parse_ir(const void* ir, size_t bytes, ModelIR& m, std::string& err) {
  const i64* w = (const i64*)ir;
  size_t nw = bytes / 8, q = 0;
  bool ok = true;
  auto rd = [&]() -> i64 { if (q >= nw) { ok = false; return 0; } return w[q++]; };
  if (bytes % 8 != 0 || nw < 6) { err = "IR too short"; return false; }
  if (rd() != 0x0031425845LL) { err = "bad IR magic"; return false; }
  if (rd() != 1) { err = "unsupported IR version"; return false; }
  m.nvar = rd(); m.npar = rd();
  if (m.nvar < 0 || m.npar < 0) { err = "negative nvar / npar"; return false; }
  i64 npat = rd(); m.ndatabufs = (int)rd();
  if (npat < 0 || npat > (1 << 20)) { err = "bad pattern count"; return false; }
  m.pats.resize((size_t)npat);
  for (size_t pi = 0; pi < m.pats.size(); pi++) {
    PatternIR& p = m.pats[pi];
    p.kind = (int)rd(); p.nitr = rd(); p.itr_kind = (int)rd(); p.range_start = rd();
    p.databuf = (int)rd(); p.stride = rd();
    if (p.itr_kind == ITR_AOS) p.range_start = 0;   // point numbers of an AoS iterator start at 0 (see affine_index)
    i64 nf = rd();
    if (!ok || nf < 0 || nf > 4096) { err = "bad field count"; return false; }
    p.fields.resize((size_t)nf);
    for (auto& f : p.fields) { f.off = rd(); f.type = rd(); }
    p.o0 = rd(); p.o1 = rd(); p.o2 = rd(); p.base = (int)rd();
    i64 nidx = rd();
    if (!ok || nidx < 0 || nidx > 64) { err = "bad index-expression count"; return false; }
    p.idx_roots.resize((size_t)nidx); p.dims.resize((size_t)nidx);
    for (auto& r : p.idx_roots) r = (int)rd();
    for (auto& d : p.dims) d = rd();
    i64 nn = rd();
    if (!ok || nn <= 0 || (size_t)nn > nw) { err = "bad node count"; return false; }
    p.nodes.resize((size_t)nn);
    for (size_t k = 0; k < p.nodes.size(); k++) {
      IRNode& n = p.nodes[k];
      n.tag = rd(); n.a = rd(); n.b = rd(); n.payload = rd();
      bool child_ok = true;
      if (n.tag == T_VAR || n.tag == T_PAR || n.tag == T_OP1 || n.tag == T_OP2) child_ok = n.a >= 0 && (size_t)n.a < k;
      if (n.tag == T_OP2) child_ok = child_ok && n.b >= 0 && (size_t)n.b < k;
      if (n.tag == T_DATA_FIELD) child_ok = n.a >= 0 && n.a < nf;
      if (n.tag == T_OP1) child_ok = child_ok && n.payload >= 0 && n.payload < U_COUNT;
      if (n.tag == T_OP2) child_ok = child_ok && n.payload >= 0 && n.payload < B_COUNT;
      if (n.tag < 0 || n.tag > T_VAL || !child_ok) { err = "malformed IR node"; return false; }
    }
    p.root = (int)rd();
    i64 nc1 = rd(); if (!ok || nc1 < 0 || (size_t)nc1 > nw) { err = "bad comp1"; return false; }
    p.comp1_given.resize((size_t)nc1); for (auto& c : p.comp1_given) c = rd();
    i64 nc2 = rd(); if (!ok || nc2 < 0 || (size_t)nc2 > nw) { err = "bad comp2"; return false; }
    p.comp2_given.resize((size_t)nc2); for (auto& c : p.comp2_given) c = rd();
    if (!ok) { err = "truncated IR"; return false; }
    if (p.root < 0 || p.root >= nn) { err = "bad root"; return false; }
    for (int r : p.idx_roots) if (r < 0 || r >= nn) { err = "bad index root"; return false; }
    if (p.kind < KIND_OBJ || p.kind > KIND_AUG) { err = "bad pattern kind"; return false; }
    if (p.kind == KIND_AUG && (p.base < 0 || (size_t)p.base >= pi || m.pats[(size_t)p.base].kind != KIND_CON)) {
      err = "augmentation must reference an earlier Constraint pattern"; return false;
    }
    if (p.itr_kind == ITR_AOS && (p.databuf < 0 || p.databuf >= m.ndatabufs)) { err = "bad data buffer index"; return false; }
    if (p.nitr < 0) { err = "negative iterator length"; return false; }
    if (p.itr_kind != ITR_RANGE && p.itr_kind != ITR_AOS) { err = "unknown iterator kind"; return false; }
    for (auto& f : p.fields) {
      if (f.type < FT_I64 || f.type > FT_F32) { err = "unknown field type"; return false; }
      const i64 sz = (f.type == FT_I64 || f.type == FT_F64) ? 8 : 4;
      if (p.itr_kind == ITR_AOS && (f.off < 0 || f.off + sz > p.stride)) { err = "field offset outside the iterator element"; return false; }
    }
    for (auto& n : p.nodes) {
      if (n.tag == T_DATA_SELF && p.itr_kind != ITR_RANGE) { err = "DATA_SELF on a non-range iterator"; return false; }
      if (n.tag == T_DATA_FIELD && p.itr_kind != ITR_AOS) { err = "DATA_FIELD on a range iterator"; return false; }
    }
  }
  return true;
}

// An integer field whose values are v0, v0 + 1, v0 + 2, ... makes an AoS pattern as good as a range pattern: indices built
// from it are affine in the point number, so index relations, x windows, owner-computed gradients and the duplicate-free
// Hessian are decided at build time, and the column is never loaded (or even uploaded).
inline void detect_iota(ModelIR& m, const void* const* host_data, int n_data) {
  for (auto& p : m.pats) {
    p.iota.assign(p.fields.size(), 0); p.iota0.assign(p.fields.size(), 0);
    if (p.itr_kind != ITR_AOS || !host_data || p.databuf < 0 || p.databuf >= n_data || !host_data[p.databuf] || p.nitr < 1) continue;
    const unsigned char* base = (const unsigned char*)host_data[p.databuf];
    for (size_t f = 0; f < p.fields.size(); f++) {
      const Field& fd = p.fields[f];
      if (fd.type != FT_I64 && fd.type != FT_I32) continue;
      auto at = [&](i64 k) -> i64 {
        const unsigned char* q = base + (size_t)k * (size_t)p.stride + fd.off;
        if (fd.type == FT_I64) { int64_t t; std::memcpy(&t, q, 8); return t; }
        int32_t t; std::memcpy(&t, q, 4); return t;
      };
      const i64 v0 = at(0);
      bool ok = v0 > -(1LL << 40) && v0 < (1LL << 40);
      for (i64 k = 1; k < p.nitr && ok; k++) ok = at(k) == v0 + k;
      if (ok) { p.iota[f] = 1; p.iota0[f] = v0; }
    }
  }
}

}  // namespace exb
