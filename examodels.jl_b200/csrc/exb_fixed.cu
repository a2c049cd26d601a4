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

#define EXB_TYPES_ONLY 1
#include "exb_device.cuh"   // ExbTile
#include "exb_fixed.h"

namespace {

// one thread per distinct target: y[target[t]] (+)= sum_{l in [ptr[t], ptr[t+1])} buf[slot[l]]
// IDX = int when every index fits 32 bits (halves the index traffic), else long long.
// target == nullptr means "every row/variable 1..nt is a target, in order" (dense case).
template <bool ACC, typename IDX>
__global__ void __launch_bounds__(256) k_compress(const double* __restrict__ buf, const IDX* __restrict__ ptr,
                                                  const IDX* __restrict__ slot, const IDX* __restrict__ target,
                                                  long long nt, double* __restrict__ y) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= nt) return;
  const long long lo = __ldg(ptr + t), hi = __ldg(ptr + t + 1);
  double s = 0.0;
  for (long long l = lo; l < hi; l++) s += __ldg(buf + __ldg(slot + l));
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
template <bool ACC, bool SKIPDIAG, typename IDX>
__global__ void __launch_bounds__(256) k_spmv(const double* __restrict__ buf, const IDX* __restrict__ ptr, const IDX* __restrict__ slot,
                                              const IDX* __restrict__ other, const IDX* __restrict__ target, long long nt,
                                              const double* __restrict__ v, double* __restrict__ y) {
  const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
  if (t >= nt) return;
  const long long lo = __ldg(ptr + t), hi = __ldg(ptr + t + 1);
  const long long tg = target ? (long long)__ldg(target + t) : t + 1;
  double s = 0.0;
  for (long long l = lo; l < hi; l++) {
    const long long o = __ldg(other + l);
    if (SKIPDIAG && o == tg) continue;
    s += __ldg(buf + __ldg(slot + l)) * __ldg(v + (o - 1));
  }
  if (ACC) y[tg - 1] += s; else y[tg - 1] = s;
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
                            long long nt, double* y, int accumulate, cudaStream_t st) {
  if (nt <= 0) return cudaSuccess;
  const unsigned grid = (unsigned)((nt + 255) / 256);
  if (idx32) {
    if (accumulate) k_compress<true, int><<<grid, 256, 0, st>>>(buf, (const int*)ptr, (const int*)slot, (const int*)target, nt, y);
    else k_compress<false, int><<<grid, 256, 0, st>>>(buf, (const int*)ptr, (const int*)slot, (const int*)target, nt, y);
  } else {
    if (accumulate) k_compress<true, long long><<<grid, 256, 0, st>>>(buf, (const long long*)ptr, (const long long*)slot, (const long long*)target, nt, y);
    else k_compress<false, long long><<<grid, 256, 0, st>>>(buf, (const long long*)ptr, (const long long*)slot, (const long long*)target, nt, y);
  }
  return cudaGetLastError();
}

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
                        long long nt, const double* v, double* y, int accumulate, int skipdiag, cudaStream_t st) {
  if (nt <= 0) return cudaSuccess;
  const unsigned grid = (unsigned)((nt + 255) / 256);
#define EXB_SPMV(A, S, T) k_spmv<A, S, T><<<grid, 256, 0, st>>>(buf, (const T*)ptr, (const T*)slot, (const T*)other, (const T*)target, nt, v, y)
  if (idx32) {
    if (accumulate) { if (skipdiag) EXB_SPMV(true, true, int); else EXB_SPMV(true, false, int); }
    else { if (skipdiag) EXB_SPMV(false, true, int); else EXB_SPMV(false, false, int); }
  } else {
    if (accumulate) { if (skipdiag) EXB_SPMV(true, true, long long); else EXB_SPMV(true, false, long long); }
    else { if (skipdiag) EXB_SPMV(false, true, long long); else EXB_SPMV(false, false, long long); }
  }
#undef EXB_SPMV
  return cudaGetLastError();
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
