// Shared helpers for the fh_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define FH_OK 0
#define FH_ERR_ARG 1
#define FH_ERR_CUDA 2
#define FH_ERR_UNSUPPORTED 3

extern "C" void fh_set_error(const char* fmt, ...);

#define FH_CHECK_ARG(cond, ...)            \
	do {                                   \
		if (!(cond)) {                     \
			fh_set_error(__VA_ARGS__);     \
			return FH_ERR_ARG;             \
		}                                  \
	} while (0)

#define FH_CUDA(expr)                                                                       \
	do {                                                                                    \
		cudaError_t _e = (expr);                                                            \
		if (_e != cudaSuccess) {                                                            \
			fh_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
			return FH_ERR_CUDA;                                                             \
		}                                                                                   \
	} while (0)

extern "C" void fh_count_launch(int n);
#define FH_LAUNCH_CHECK()            \
	do {                             \
		fh_count_launch(1);          \
		FH_CUDA(cudaGetLastError()); \
	} while (0)

// optional CUDA-event timing of a kernel launch (fh_api.cu; ids = FH_TIME_* of include/fh_b200.h)
extern "C" int fh_time_begin(int id, void* stream);
extern "C" void fh_time_end(int idx, void* stream);

static inline int fh_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float fh_warp_sum(float v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}
__device__ __forceinline__ double fh_warp_sum(double v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}
__device__ __forceinline__ float fh_warp_max(float v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}

// block-wide sum; `red` must hold >= 32 elements; result valid in every thread
template <typename T>
__device__ __forceinline__ T fh_block_sum(T v, T* red) {
	v = fh_warp_sum(v);
	int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	__syncthreads();
	if (lane == 0) red[wid] = v;
	__syncthreads();
	int nw = (blockDim.x + 31) >> 5;
	T r = (threadIdx.x < nw) ? red[threadIdx.x] : T(0);
	if (wid == 0) {
		r = fh_warp_sum(r);
		if (lane == 0) red[0] = r;
	}
	__syncthreads();
	return red[0];
}
