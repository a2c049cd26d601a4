// Host-callable entry points of exb_fixed.cu (internal to the runtime library).
#pragma once
#include <cuda_runtime_api.h>

#define EXB_FX_SENTINEL 0x7fffffffffffffffLL

cudaError_t exb_fx_compress(const double* buf, const long long* ptr, const long long* slot, const long long* target,
                            long long nt, double* y, int accumulate, cudaStream_t st);
cudaError_t exb_fx_sum(const double* part, long long n, double* out, cudaStream_t st);
cudaError_t exb_fx_fill(long long* p, long long n, long long v, cudaStream_t st);
cudaError_t exb_fx_sort_runs(const long long* keys, long long n, long long** slot_out, long long** target_out,
                             long long** ptr_out, long long* nruns_out, cudaStream_t st);
