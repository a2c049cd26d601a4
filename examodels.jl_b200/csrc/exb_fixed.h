// Host-callable entry points of exb_fixed.cu (internal to the runtime library).
#pragma once
#include <cuda_runtime_api.h>

#define EXB_FX_SENTINEL 0x7fffffffffffffffLL

cudaError_t exb_fx_compress(const double* buf, const void* ptr, const void* slot, const void* target, int idx32,
                            long long nt, double* y, int accumulate, const void* long_runs, cudaStream_t st);
// long runs of a packed (ptr, target) pair (see k_compress): listed once at build time, freed with exb_fx_long_free
cudaError_t exb_fx_long_runs(const void* ptr, const void* target, int idx32, long long nruns, long long nslots, void** out, cudaStream_t st);
void exb_fx_long_free(void* long_runs);
int exb_fx_long_count(const void* long_runs);
cudaError_t exb_fx_pack_runs(void** slot, void** target, void** ptr, long long nslots, long long nruns, long long max_index,
                             int* idx32, int* dense, cudaStream_t st);
cudaError_t exb_fx_sum(const double* part, long long n, double* out, cudaStream_t st);
cudaError_t exb_fx_fill(long long* p, long long n, long long v, cudaStream_t st);
cudaError_t exb_fx_sort_runs(const long long* keys, long long n, long long** slot_out, long long** target_out,
                             long long** ptr_out, long long* nruns_out, long long* nslots_out, cudaStream_t st);
cudaError_t exb_fx_spmv(const double* buf, const void* ptr, const void* slot, const void* other, const void* target, int idx32,
                        long long nt, const double* v, double* y, int accumulate, int skipdiag, const void* long_runs, cudaStream_t st);
cudaError_t exb_fx_gather(const long long* src, const void* slot, int idx32, void* out, long long n, cudaStream_t st);
cudaError_t exb_fx_make_keys(const long long* major, const long long* minor, long long mult, long long* keys, long long n, cudaStream_t st);
cudaError_t exb_fx_decode_keys(const long long* keys, long long mult, long long* major, long long* minor, long long n, cudaStream_t st);
// tile: const ExbTile* (exb_device.cuh); rows / cols: int64 (idx32 = 0) or int32 device arrays
cudaError_t exb_fx_tile_structure(const void* tile, void* rows, void* cols, int idx32, cudaStream_t st);
cudaError_t exb_fx_narrow(const long long* in, int* out, long long n, cudaStream_t st);
