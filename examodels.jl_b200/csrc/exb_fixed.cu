// Fixed (model-independent) device code of the runtime library: the deterministic reductions
// that follow the generated per-pattern kernels, and the build-time sort that prepares them.
//
//   exb_fx_compress   <- compress_to_dense   ext/ExaModelsKernelAbstractions.jl:691-697
//   exb_fx_sum        <- sum(objbuffer)      ext:259
//   exb_fx_sort_runs  <- sort! + getptr      ext:12-18,44-53,699-715   (build time; CUB radix sort)
//
// No atomics: every dense target is owned by one thread which adds its (pre-sorted) slots in a
// fixed order, so grad! and the augmentation part of cons! are bitwise reproducible run to run.
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include <vector>

#define EXB_TYPES_ONLY 1
#include "exb_device.cuh"   // ExbTile
#include "exb_fixed.h"

namespace {

// one thread per distinct target: y[target[t]] (+)= sum_{l in [ptr[t], ptr[t+1])} buf[slot[l]]
// IDX = int when every index fits 32 bits (halves the index traffic), else long long.
// target == nullptr means "every row/variable 1..nt is a target, in order" (dense case).
// Long runs (a variable at a fixed index shared by every point -- a step length --, a row that collects a whole pattern) would be
// summed serially by their one thread (the reference's compress_to_dense / kerspmv do exactly that, ext:482-511,691-697): a run of
// more than EXB_FX_COOP slots is summed by the thread's whole WARP (lane-strided partial sums + a fixed shuffle tree), and runs of
// at least `long_thr` slots are left to the chunked kernels below (k_long_partial / k_long_finish).  Every order is fixed.
#define EXB_FX_COOP 64
#define EXB_FX_LONG 8192       // runs at least this long are cut into chunks of EXB_FX_LONG slots, one block each
__device__ __forceinline__ double warp_tree(double p) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) p += __shfl_down_sync(0xffffffffu, p, o);
  return __shfl_sync(0xffffffffu, p, 0);
}
template <bool ACC, typename IDX>
__global__ void __launch_bounds__(256) k_compress(const double* __restrict__ buf, const IDX* __restrict__ ptr,
                                                  const IDX* __restrict__ slot, const IDX* __restrict__ target,
                                                  long long nt, double* __restrict__ y, long long long_thr) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  const bool live = t < nt;
  long long lo = 0, hi = 0;
  if (live) { lo = __ldg(ptr + t); hi = __ldg(ptr + t + 1); }
  const bool chunked = hi - lo >= long_thr, coop = !chunked && hi - lo > EXB_FX_COOP;
  double s = 0.0;
  if (!chunked && !coop)
    for (long long l = lo; l < hi; l++) s += __ldg(buf + __ldg(slot + l));
  unsigned todo = __ballot_sync(0xffffffffu, coop);
  const int lane = threadIdx.x & 31;
  while (todo) {                                   // warp-uniform
    const int src = __ffs(todo) - 1; todo &= todo - 1;
    const long long clo = __shfl_sync(0xffffffffu, lo, src), chi = __shfl_sync(0xffffffffu, hi, src);
    double p = 0.0;
    for (long long l = clo + lane; l < chi; l += 32) p += __ldg(buf + __ldg(slot + l));
    p = warp_tree(p);
    if (lane == src) s = p;
  }
  if (!live || chunked) return;
  const long long k = target ? (long long)__ldg(target + t) - 1 : t;
  if (ACC) y[k] += s; else y[k] = s;
}

__global__ void k_narrow(const long long* __restrict__ in, int* __restrict__ out, long long n) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t < n) out[t] = (int)in[t];
}
// flag[0] = 1 iff target[t] == t + 1 for all t
__global__ void k_is_iota(const long long* __restrict__ target, long long n, int* flag) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t < n && target[t] != t + 1) *flag = 0;
}

// SpMV over a pre-sorted COO structure (kerspmv / kerspmv2 / kersyspmv / kersyspmv2, ext:482-511):
// one thread per distinct target index; y[target] (+)= sum_l buf[slot[l]] * v[other[l]].
// SKIPDIAG drops entries whose other index equals the target (strict-triangle pass of the symmetric product).
// Long runs as in k_compress.
template <bool ACC, bool SKIPDIAG, typename IDX>
__global__ void __launch_bounds__(256) k_spmv(const double* __restrict__ buf, const IDX* __restrict__ ptr, const IDX* __restrict__ slot,
                                              const IDX* __restrict__ other, const IDX* __restrict__ target, long long nt,
                                              const double* __restrict__ v, double* __restrict__ y, long long long_thr) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  const bool live = t < nt;
  long long lo = 0, hi = 0, tg = 0;
  if (live) { lo = __ldg(ptr + t); hi = __ldg(ptr + t + 1); tg = target ? (long long)__ldg(target + t) : t + 1; }
  const bool chunked = hi - lo >= long_thr, coop = !chunked && hi - lo > EXB_FX_COOP;
  double s = 0.0;
  if (!chunked && !coop)
    for (long long l = lo; l < hi; l++) {
      const long long o = __ldg(other + l);
      if (SKIPDIAG && o == tg) continue;
      s += __ldg(buf + __ldg(slot + l)) * __ldg(v + (o - 1));
    }
  unsigned todo = __ballot_sync(0xffffffffu, coop);
  const int lane = threadIdx.x & 31;
  while (todo) {
    const int src = __ffs(todo) - 1; todo &= todo - 1;
    const long long clo = __shfl_sync(0xffffffffu, lo, src), chi = __shfl_sync(0xffffffffu, hi, src), ctg = __shfl_sync(0xffffffffu, tg, src);
    double p = 0.0;
    for (long long l = clo + lane; l < chi; l += 32) {
      const long long o = __ldg(other + l);
      if (SKIPDIAG && o == ctg) continue;
      p += __ldg(buf + __ldg(slot + l)) * __ldg(v + (o - 1));
    }
    p = warp_tree(p);
    if (lane == src) s = p;
  }
  if (!live || chunked) return;
  if (ACC) y[tg - 1] += s; else y[tg - 1] = s;
}

// ---- runs of >= EXB_FX_LONG slots: listed once at build time (exb_fx_long_runs), cut into chunks ----
struct LongRuns {
  int nruns = 0, nchunks = 0;
  long long *chunk_lo = nullptr, *chunk_hi = nullptr, *run_tg = nullptr;   // device
  int *chunk_run = nullptr, *run_c0 = nullptr;                              // device; run_c0[nruns + 1]
  double* partial = nullptr;                                                // device, one per chunk
};
template <typename IDX>
__global__ void k_find_long(const IDX* __restrict__ ptr, const IDX* __restrict__ target, long long nt, long long thr, int cap,
                            long long* __restrict__ out4, int* __restrict__ count) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= nt) return;
  const long long lo = ptr[t], hi = ptr[t + 1];
  if (hi - lo < thr) return;
  const int i = atomicAdd(count, 1);
  if (i < cap) { out4[4 * i] = t; out4[4 * i + 1] = lo; out4[4 * i + 2] = hi; out4[4 * i + 3] = target ? (long long)target[t] : t + 1; }
}
// one block per chunk: lane-strided partial sums, warp trees, then the 8 warp sums added in order
template <bool SPMV, bool SKIPDIAG, typename IDX>
__global__ void __launch_bounds__(256) k_long_partial(const double* __restrict__ buf, const IDX* __restrict__ slot, const IDX* __restrict__ other,
                                                      const double* __restrict__ v, const long long* __restrict__ chunk_lo,
                                                      const long long* __restrict__ chunk_hi, const int* __restrict__ chunk_run,
                                                      const long long* __restrict__ run_tg, double* __restrict__ partial) {
  __shared__ double sm[8];
  const int c = blockIdx.x;
  const long long lo = chunk_lo[c], hi = chunk_hi[c], tg = run_tg[chunk_run[c]];
  double p = 0.0;
  for (long long l = lo + threadIdx.x; l < hi; l += 256) {
    if (SPMV) {
      const long long o = __ldg(other + l);
      if (SKIPDIAG && o == tg) continue;
      p += __ldg(buf + __ldg(slot + l)) * __ldg(v + (o - 1));
    } else {
      p += __ldg(buf + __ldg(slot + l));
    }
  }
  p = warp_tree(p);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = p;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < 8; w++) s += sm[w];
    partial[c] = s;
  }
}
// one warp per long run: its chunk partials, lane-strided + fixed tree
template <bool ACC>
__global__ void __launch_bounds__(32) k_long_finish(const long long* __restrict__ run_tg, const int* __restrict__ run_c0,
                                                    const double* __restrict__ partial, double* __restrict__ y) {
  const int r = blockIdx.x;
  double p = 0.0;
  for (int c = run_c0[r] + (int)threadIdx.x; c < run_c0[r + 1]; c += 32) p += partial[c];
  p = warp_tree(p);
  if (threadIdx.x == 0) { const long long k = run_tg[r] - 1; if (ACC) y[k] += p; else y[k] = p; }
}
template <bool SPMV>
cudaError_t long_launch(const LongRuns* L, const double* buf, const void* slot, const void* other, int idx32, const double* v, double* y,
                        int accumulate, int skipdiag, cudaStream_t st) {
  if (!L || L->nchunks == 0) return cudaSuccess;
#define EXB_LP(S, T) k_long_partial<SPMV, S, T><<<(unsigned)L->nchunks, 256, 0, st>>>(buf, (const T*)slot, (const T*)other, v, L->chunk_lo, L->chunk_hi, \
                                                                                      L->chunk_run, L->run_tg, L->partial)
  if (idx32) { if (skipdiag) EXB_LP(true, int); else EXB_LP(false, int); }
  else { if (skipdiag) EXB_LP(true, long long); else EXB_LP(false, long long); }
#undef EXB_LP
  if (accumulate) k_long_finish<true><<<(unsigned)L->nruns, 32, 0, st>>>(L->run_tg, L->run_c0, L->partial, y);
  else k_long_finish<false><<<(unsigned)L->nruns, 32, 0, st>>>(L->run_tg, L->run_c0, L->partial, y);
  return cudaGetLastError();
}

template <typename IDX>
__global__ void k_gather(const long long* __restrict__ src, const IDX* __restrict__ slot, IDX* __restrict__ out, long long n) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t < n) out[t] = (IDX)src[slot[t]];
}
__global__ void k_make_keys(const long long* __restrict__ major, const long long* __restrict__ minor, long long mult, long long* keys, long long n) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t < n) keys[t] = major[t] * mult + minor[t];
}
__global__ void k_decode_keys(const long long* __restrict__ keys, long long mult, long long* major, long long* minor, long long n) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t < n) { major[t] = keys[t] / mult; minor[t] = keys[t] % mult; }
}

// deterministic sum of n partials into out[0]: fixed per-thread strided order + fixed tree
__global__ void __launch_bounds__(1024) k_sum(const double* __restrict__ part, long long n, double* __restrict__ out) {
  __shared__ double sm[32];
  double v = 0.0;
  for (long long k = threadIdx.x; k < n; k += 1024) v += part[k];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = sm[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) out[0] = v;
  }
}

// (row, col) of the duplicate-free Hessian of a shift-indexed model, in (col, row) order: one thread per column
// (see ExbTile; the values come from exb_tile_body in the generated module)
template <typename I>
__global__ void k_tile_structure(const ExbTile t, I* __restrict__ rows, I* __restrict__ cols) {
  const long long c = t.c_lo + (long long)blockIdx.x * 256 + threadIdx.x;
  if (c >= t.c_hi) return;
  long long p = 0;
  for (int r = 0; r < t.D; r++) { long long v = c - t.lo[r]; v = v < 0 ? 0 : v; p += v > t.len[r] ? t.len[r] : v; }
  for (int r = 0; r < t.D; r++)
    if (c >= t.lo[r] && c < t.lo[r] + t.len[r]) { rows[p] = (I)(c + t.dist[r]); cols[p] = (I)c; p++; }
}

__global__ void k_fill_ll(long long* p, long long n, long long v) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t < n) p[t] = v;
}
__global__ void k_iota_ll(long long* p, long long n) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t < n) p[t] = t;
}

}  // namespace

cudaError_t exb_fx_compress(const double* buf, const void* ptr, const void* slot, const void* target, int idx32,
                            long long nt, double* y, int accumulate, const void* long_runs, cudaStream_t st) {
  if (nt <= 0) return cudaSuccess;
  const unsigned grid = (unsigned)((nt + 255) / 256);
  const LongRuns* L = (const LongRuns*)long_runs;
  const long long thr = L ? EXB_FX_LONG : 0x7fffffffffffffffLL;
  if (idx32) {
    if (accumulate) k_compress<true, int><<<grid, 256, 0, st>>>(buf, (const int*)ptr, (const int*)slot, (const int*)target, nt, y, thr);
    else k_compress<false, int><<<grid, 256, 0, st>>>(buf, (const int*)ptr, (const int*)slot, (const int*)target, nt, y, thr);
  } else {
    if (accumulate) k_compress<true, long long><<<grid, 256, 0, st>>>(buf, (const long long*)ptr, (const long long*)slot, (const long long*)target, nt, y, thr);
    else k_compress<false, long long><<<grid, 256, 0, st>>>(buf, (const long long*)ptr, (const long long*)slot, (const long long*)target, nt, y, thr);
  }
  cudaError_t e = cudaGetLastError();
  return e != cudaSuccess ? e : long_launch<false>(L, buf, slot, nullptr, idx32, nullptr, y, accumulate, 0, st);
}

// Build-time list of the runs of >= EXB_FX_LONG slots of a packed (ptr, target) pair; *out = nullptr when there is none.
cudaError_t exb_fx_long_runs(const void* ptr, const void* target, int idx32, long long nruns, long long nslots, void** out, cudaStream_t st) {
  *out = nullptr;
  if (nruns <= 0 || nslots < EXB_FX_LONG) return cudaSuccess;
  const int cap = (int)(nslots / EXB_FX_LONG) + 1;
  long long* d4 = nullptr; int* dc = nullptr; cudaError_t e;
  if ((e = cudaMalloc(&d4, (size_t)cap * 32)) != cudaSuccess) return e;
  if ((e = cudaMalloc(&dc, 4)) != cudaSuccess) { cudaFree(d4); return e; }
  cudaMemsetAsync(dc, 0, 4, st);
  const unsigned grid = (unsigned)((nruns + 255) / 256);
  if (idx32) k_find_long<int><<<grid, 256, 0, st>>>((const int*)ptr, (const int*)target, nruns, EXB_FX_LONG, cap, d4, dc);
  else k_find_long<long long><<<grid, 256, 0, st>>>((const long long*)ptr, (const long long*)target, nruns, EXB_FX_LONG, cap, d4, dc);
  int n = 0;
  cudaMemcpyAsync(&n, dc, 4, cudaMemcpyDeviceToHost, st);
  e = cudaStreamSynchronize(st);
  std::vector<long long> h4((size_t)(n > 0 ? n : 0) * 4);
  if (e == cudaSuccess && n > 0) e = cudaMemcpy(h4.data(), d4, h4.size() * 8, cudaMemcpyDeviceToHost);
  cudaFree(d4); cudaFree(dc);
  if (e != cudaSuccess || n <= 0) return e;
  std::vector<int> order((size_t)n);
  for (int i = 0; i < n; i++) order[(size_t)i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return h4[4 * (size_t)a] < h4[4 * (size_t)b]; });   // by run number: a fixed layout
  std::vector<long long> clo, chi, tg; std::vector<int> crun, c0;
  for (int r = 0; r < n; r++) {
    const size_t i = (size_t)order[(size_t)r];
    c0.push_back((int)clo.size()); tg.push_back(h4[4 * i + 3]);
    for (long long l = h4[4 * i + 1]; l < h4[4 * i + 2]; l += EXB_FX_LONG) {
      clo.push_back(l); chi.push_back(std::min(l + (long long)EXB_FX_LONG, h4[4 * i + 2])); crun.push_back(r);
    }
  }
  c0.push_back((int)clo.size());
  LongRuns* L = new LongRuns();
  L->nruns = n; L->nchunks = (int)clo.size();
  auto up = [&](void** d, const void* h, size_t bytes) { cudaError_t q = cudaMalloc(d, bytes); return q != cudaSuccess ? q : cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice); };
  e = up((void**)&L->chunk_lo, clo.data(), clo.size() * 8);
  if (e == cudaSuccess) e = up((void**)&L->chunk_hi, chi.data(), chi.size() * 8);
  if (e == cudaSuccess) e = up((void**)&L->run_tg, tg.data(), tg.size() * 8);
  if (e == cudaSuccess) e = up((void**)&L->chunk_run, crun.data(), crun.size() * 4);
  if (e == cudaSuccess) e = up((void**)&L->run_c0, c0.data(), c0.size() * 4);
  if (e == cudaSuccess) e = cudaMalloc(&L->partial, clo.size() * 8);
  if (e != cudaSuccess) { exb_fx_long_free(L); return e; }
  *out = L;
  return cudaSuccess;
}
void exb_fx_long_free(void* long_runs) {
  LongRuns* L = (LongRuns*)long_runs;
  if (!L) return;
  cudaFree(L->chunk_lo); cudaFree(L->chunk_hi); cudaFree(L->run_tg); cudaFree(L->chunk_run); cudaFree(L->run_c0); cudaFree(L->partial);
  delete L;
}
int exb_fx_long_count(const void* long_runs) { return long_runs ? ((const LongRuns*)long_runs)->nruns : 0; }

// In-place post-processing of a sort_runs result: drop `target` when it is 1..nruns (returns *dense = 1),
// and narrow the three arrays to int32 when everything fits (returns *idx32 = 1; arrays are re-allocated).
cudaError_t exb_fx_pack_runs(void** slot, void** target, void** ptr, long long nslots, long long nruns, long long max_index,
                             int* idx32, int* dense, cudaStream_t st) {
  *idx32 = 0; *dense = 0;
  if (nruns <= 0) return cudaSuccess;
  cudaError_t e;
  if (!*target) { *dense = 1; }
  int* flag = nullptr; int h = 1;
  if (*target) {
  if ((e = cudaMalloc(&flag, 4)) != cudaSuccess) return e;
  cudaMemcpyAsync(flag, &h, 4, cudaMemcpyHostToDevice, st);
  k_is_iota<<<(unsigned)((nruns + 255) / 256), 256, 0, st>>>((const long long*)*target, nruns, flag);
  cudaMemcpyAsync(&h, flag, 4, cudaMemcpyDeviceToHost, st);
  e = cudaStreamSynchronize(st);
  cudaFree(flag);
  if (e != cudaSuccess) return e;
  if (h) { cudaFree(*target); *target = nullptr; *dense = 1; }
  }
  if (max_index < 2147483647LL && nslots < 2147483647LL) {
    void** arr[3] = {slot, target, ptr};
    long long len[3] = {nslots, nruns, nruns + 1};
    for (int k = 0; k < 3; k++) {
      if (!*arr[k]) continue;
      int* out = nullptr;
      if ((e = cudaMalloc(&out, (size_t)(len[k] ? len[k] : 1) * 4)) != cudaSuccess) return e;
      if (len[k] > 0) k_narrow<<<(unsigned)((len[k] + 255) / 256), 256, 0, st>>>((const long long*)*arr[k], out, len[k]);
      if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
      cudaFree(*arr[k]);
      *arr[k] = out;
    }
    *idx32 = 1;
  }
  return cudaGetLastError();
}

cudaError_t exb_fx_spmv(const double* buf, const void* ptr, const void* slot, const void* other, const void* target, int idx32,
                        long long nt, const double* v, double* y, int accumulate, int skipdiag, const void* long_runs, cudaStream_t st) {
  if (nt <= 0) return cudaSuccess;
  const unsigned grid = (unsigned)((nt + 255) / 256);
  const LongRuns* L = (const LongRuns*)long_runs;
  const long long thr = L ? EXB_FX_LONG : 0x7fffffffffffffffLL;
#define EXB_SPMV(A, S, T) k_spmv<A, S, T><<<grid, 256, 0, st>>>(buf, (const T*)ptr, (const T*)slot, (const T*)other, (const T*)target, nt, v, y, thr)
  if (idx32) {
    if (accumulate) { if (skipdiag) EXB_SPMV(true, true, int); else EXB_SPMV(true, false, int); }
    else { if (skipdiag) EXB_SPMV(false, true, int); else EXB_SPMV(false, false, int); }
  } else {
    if (accumulate) { if (skipdiag) EXB_SPMV(true, true, long long); else EXB_SPMV(true, false, long long); }
    else { if (skipdiag) EXB_SPMV(false, true, long long); else EXB_SPMV(false, false, long long); }
  }
#undef EXB_SPMV
  cudaError_t e = cudaGetLastError();
  return e != cudaSuccess ? e : long_launch<true>(L, buf, slot, other, idx32, v, y, accumulate, skipdiag, st);
}
// out[l] = src[slot[l]] for the sorted positions l (the "other" index of each sorted COO entry)
cudaError_t exb_fx_gather(const long long* src, const void* slot, int idx32, void* out, long long n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (idx32) k_gather<int><<<grid, 256, 0, st>>>(src, (const int*)slot, (int*)out, n);
  else k_gather<long long><<<grid, 256, 0, st>>>(src, (const long long*)slot, (long long*)out, n);
  return cudaGetLastError();
}
cudaError_t exb_fx_make_keys(const long long* major, const long long* minor, long long mult, long long* keys, long long n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_make_keys<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(major, minor, mult, keys, n);
  return cudaGetLastError();
}
cudaError_t exb_fx_decode_keys(const long long* keys, long long mult, long long* major, long long* minor, long long n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_decode_keys<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(keys, mult, major, minor, n);
  return cudaGetLastError();
}

cudaError_t exb_fx_tile_structure(const void* tile, void* rows, void* cols, int idx32, cudaStream_t st) {
  const ExbTile& t = *(const ExbTile*)tile;
  if (t.c_hi <= t.c_lo) return cudaSuccess;
  const unsigned grid = (unsigned)((t.c_hi - t.c_lo + 255) / 256);
  if (idx32) k_tile_structure<int><<<grid, 256, 0, st>>>(t, (int*)rows, (int*)cols);
  else k_tile_structure<long long><<<grid, 256, 0, st>>>(t, (long long*)rows, (long long*)cols);
  return cudaGetLastError();
}

cudaError_t exb_fx_narrow(const long long* in, int* out, long long n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_narrow<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, n);
  return cudaGetLastError();
}

cudaError_t exb_fx_sum(const double* part, long long n, double* out, cudaStream_t st) {
  k_sum<<<1, 1024, 0, st>>>(part, n, out);
  return cudaGetLastError();
}

cudaError_t exb_fx_fill(long long* p, long long n, long long v, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_fill_ll<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, n, v);
  return cudaGetLastError();
}

// keys[n] (1-based targets; EXB_FX_SENTINEL = not owned by this handle) -> stable sort by key with the
// slot number as payload, then run-length encode.  Outputs are freshly cudaMalloc'ed arrays sized to the
// number of owned runs: *slot_out[n_owned], *target_out[nruns], *ptr_out[nruns + 1].
cudaError_t exb_fx_sort_runs(const long long* keys, long long n, long long** slot_out, long long** target_out,
                             long long** ptr_out, long long* nruns_out, long long* nslots_out, cudaStream_t st) {
  *slot_out = *target_out = *ptr_out = nullptr; *nruns_out = 0; *nslots_out = 0;
  if (n <= 0) return cudaSuccess;
  cudaError_t e;
  long long *vals_in = nullptr, *keys_s = nullptr, *vals_s = nullptr, *uniq = nullptr, *cnt = nullptr, *d_nruns = nullptr;
  void* tmp = nullptr; size_t tb = 0, tb2 = 0, tb3 = 0;
#define EXB_FX_TRY(x) do { e = (x); if (e != cudaSuccess) goto fail; } while (0)
  EXB_FX_TRY(cudaMalloc(&vals_in, n * 8)); EXB_FX_TRY(cudaMalloc(&keys_s, n * 8)); EXB_FX_TRY(cudaMalloc(&vals_s, n * 8));
  EXB_FX_TRY(cudaMalloc(&uniq, n * 8)); EXB_FX_TRY(cudaMalloc(&cnt, (n + 1) * 8)); EXB_FX_TRY(cudaMalloc(&d_nruns, 8));
  k_iota_ll<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(vals_in, n);
  EXB_FX_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tb, keys, keys_s, vals_in, vals_s, n, 0, 64, st));
  EXB_FX_TRY(cub::DeviceRunLengthEncode::Encode(nullptr, tb2, keys_s, uniq, cnt, d_nruns, n, st));
  EXB_FX_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tb3, cnt, cnt, n + 1, st));
  tb = tb > tb2 ? tb : tb2; tb = tb > tb3 ? tb : tb3;
  EXB_FX_TRY(cudaMalloc(&tmp, tb ? tb : 8));
  EXB_FX_TRY(cub::DeviceRadixSort::SortPairs(tmp, tb, keys, keys_s, vals_in, vals_s, n, 0, 64, st));
  EXB_FX_TRY(cub::DeviceRunLengthEncode::Encode(tmp, tb, keys_s, uniq, cnt, d_nruns, n, st));
  {
    long long nruns = 0;
    EXB_FX_TRY(cudaMemcpyAsync(&nruns, d_nruns, 8, cudaMemcpyDeviceToHost, st));
    EXB_FX_TRY(cudaStreamSynchronize(st));
    // the sentinel run (if any) sorts last: drop it
    long long last_key = 0;
    if (nruns > 0) {
      EXB_FX_TRY(cudaMemcpy(&last_key, uniq + (nruns - 1), 8, cudaMemcpyDeviceToHost));
      if (last_key == EXB_FX_SENTINEL) nruns -= 1;
    }
    if (nruns > 0) {
      // counts -> ptr (exclusive scan over nruns + 1 entries; entry nruns is scratch)
      EXB_FX_TRY(cudaMemsetAsync(cnt + nruns, 0, 8, st));
      EXB_FX_TRY(cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, cnt, nruns + 1, st));
      long long nowned = 0;
      EXB_FX_TRY(cudaMemcpyAsync(&nowned, cnt + nruns, 8, cudaMemcpyDeviceToHost, st));
      EXB_FX_TRY(cudaStreamSynchronize(st));
      *nslots_out = nowned;
      EXB_FX_TRY(cudaMalloc(slot_out, (nowned ? nowned : 1) * 8));
      EXB_FX_TRY(cudaMalloc(target_out, nruns * 8));
      EXB_FX_TRY(cudaMalloc(ptr_out, (nruns + 1) * 8));
      EXB_FX_TRY(cudaMemcpyAsync(*slot_out, vals_s, nowned * 8, cudaMemcpyDeviceToDevice, st));
      EXB_FX_TRY(cudaMemcpyAsync(*target_out, uniq, nruns * 8, cudaMemcpyDeviceToDevice, st));
      EXB_FX_TRY(cudaMemcpyAsync(*ptr_out, cnt, (nruns + 1) * 8, cudaMemcpyDeviceToDevice, st));
      EXB_FX_TRY(cudaStreamSynchronize(st));
    }
    *nruns_out = nruns;
  }
  e = cudaSuccess;
fail:
#undef EXB_FX_TRY
  cudaFree(vals_in); cudaFree(keys_s); cudaFree(vals_s); cudaFree(uniq); cudaFree(cnt); cudaFree(d_nruns); cudaFree(tmp);
  return e;
}
