// Probe of the tcgen05 kind::f16 operand layouts used by the 3xFP16 kernels (fh_rwr_chain16.cu):
// one CTA, D[128 x N] = A[128 x K] * B[K x N] on exact small-integer data, every case checked on the host.
//   case 0: A K-major SWIZZLE_64B (32 halves per row),  B K-major SWIZZLE_64B          (K = 32)
//   case 1: A K-major SWIZZLE_128B (64 halves per row), B K-major SWIZZLE_128B         (K = 64)
//   case 2: A K-major SWIZZLE_64B,                      B MN-major SWIZZLE_128B (LBO = n-group, SBO = k-group)
//   case 3: A from TMEM (two halves per 32-bit column, even k in the low half), B MN-major SWIZZLE_128B
//   case 4: as 3 with N = 64 (one 64-wide n group)
//   case 5: as 2 with the descriptor's LBO / SBO fields exchanged (only meaningful if 2 fails)
//   case 6: as 3 with the halves of a TMEM column exchanged (only meaningful if 3 fails)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I fast-higashi_b200/csrc -o gpurun_out/umma_f16_probe scripts/umma_f16_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>
#include "fh_tc.cuh"

using namespace fh_tc;

__device__ __forceinline__ uint32_t make_idesc_f16(bool a_mn, bool b_mn, int n, int m) {
	return (1u << 4) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
	asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
	asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__device__ __host__ inline int aval(int m, int k) { return ((m * 7 + k * 3) % 13) - 6; }
__device__ __host__ inline int bval(int k, int n) { return ((k * 5 + n * 11) % 17) - 8; }

__device__ int off_k_sw64(int m, int k) { return (m / 8) * 512 + (m % 8) * 64 + (((k / 8) ^ ((m % 8) >> 1)) * 16) + (k % 8) * 2; }
__device__ int off_k_sw128(int m, int k) { return (m / 8) * 1024 + (m % 8) * 128 + (((k / 8) ^ (m % 8)) * 16) + (k % 8) * 2; }
__device__ int off_mn_sw128(int k, int n, int lbo, int sbo) {
	return (n / 64) * lbo + (k / 8) * sbo + (k % 8) * 128 + ((((n % 64) / 8) ^ (k % 8)) * 16) + (n % 8) * 2;
}

__global__ void __launch_bounds__(128, 1) probe_kernel(int mode, float* out) {
	extern __shared__ uint8_t smem_raw[];
	uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
	uint8_t* sa = smem;            // 16 KB
	uint8_t* sb = smem + 16384;    // 16 KB
	uint64_t* bar = (uint64_t*)(smem + 32768);
	uint32_t* holder = (uint32_t*)(smem + 32768 + 64);
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int K = (mode == 1) ? 64 : 32;
	const int N = (mode == 4) ? 64 : 128;
	const bool a_tmem = (mode == 3 || mode == 4 || mode == 6);
	const bool b_mn = (mode >= 2);
	const int lbo = 4096, sbo = 1024;  // physical layout: 64-wide n groups 4096 B apart, 8-row k groups 1024 B apart
	if (threadIdx.x == 0) {
		mbar_init(bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder)), "r"(512) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	for (int i = threadIdx.x; i < 32768 / 4; i += 128) ((uint32_t*)smem)[i] = 0;
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = *holder;
	// fill the operands
	for (int i = threadIdx.x; i < 128 * K; i += 128) {
		const int m = i / K, k = i % K;
		const __half v = __float2half((float)aval(m, k));
		if (!a_tmem) *(__half*)(sa + (mode == 1 ? off_k_sw128(m, k) : off_k_sw64(m, k))) = v;
	}
	for (int i = threadIdx.x; i < K * N; i += 128) {
		const int k = i / N, n = i % N;
		const __half v = __float2half((float)bval(k, n));
		int off;
		if (b_mn) off = off_mn_sw128(k, n, lbo, sbo);
		else off = (mode == 1) ? off_k_sw128(n, k) : off_k_sw64(n, k);
		*(__half*)(sb + off) = v;
	}
	if (a_tmem) {  // thread = row = TMEM lane; K = 32 halves = 16 columns
		const int m = threadIdx.x;
		uint32_t v[32];
#pragma unroll
		for (int j = 0; j < 32; ++j) v[j] = 0;
#pragma unroll
		for (int j = 0; j < 16; ++j) {
			const __half lo = __float2half((float)aval(m, 2 * j)), hi = __float2half((float)aval(m, 2 * j + 1));
			const uint32_t l = __half_as_ushort(lo), h = __half_as_ushort(hi);
			v[j] = (mode == 6) ? (h | (l << 16)) : (l | (h << 16));
		}
		tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 0, v);
		tmem_st_wait();
	}
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	if (threadIdx.x == 0) {
		const uint32_t idesc = make_idesc_f16(false, b_mn, N, 128);
		const uint32_t acc = tmem + 256;
		for (int ks = 0; ks < K / 16; ++ks) {
			uint64_t da = 0, db;
			if (mode == 1) da = make_desc(smem_u32(sa) + ks * 32, 16, 1024, 2);
			else da = make_desc(smem_u32(sa) + ks * 32, 16, 512, 4);
			if (b_mn) db = (mode == 5) ? make_desc(smem_u32(sb) + ks * 2 * sbo, sbo, lbo, 2) : make_desc(smem_u32(sb) + ks * 2 * sbo, lbo, sbo, 2);
			else db = (mode == 1) ? make_desc(smem_u32(sb) + ks * 32, 16, 1024, 2) : make_desc(smem_u32(sb) + ks * 32, 16, 512, 4);
			if (a_tmem) umma_f16_ts(acc, tmem + 8 * ks, db, idesc, ks ? 1u : 0u);
			else umma_f16(acc, da, db, idesc, ks ? 1u : 0u);
		}
		umma_commit(bar);
	}
	mbar_wait(bar, 0);
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	for (int c = 0; c < N / 32; ++c) {
		uint32_t v[32];
		tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + 256 + c * 32, v);
		for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 128 + c * 32 + j] = __uint_as_float(v[j]);
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
	float* d;
	cudaMalloc(&d, 128 * 128 * 4);
	cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
	std::vector<float> h(128 * 128);
	int bad_total = 0;
	for (int mode = 0; mode <= 6; ++mode) {
		const int K = (mode == 1) ? 64 : 32, N = (mode == 4) ? 64 : 128;
		cudaMemset(d, 0, 128 * 128 * 4);
		probe_kernel<<<1, 128, 40000>>>(mode, d);
		cudaError_t e = cudaDeviceSynchronize();
		if (e != cudaSuccess) { printf("case %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
		cudaMemcpy(h.data(), d, 128 * 128 * 4, cudaMemcpyDeviceToHost);
		int bad = 0;
		for (int m = 0; m < 128; ++m)
			for (int n = 0; n < N; ++n) {
				float ref = 0;
				for (int k = 0; k < K; ++k) ref += (float)(aval(m, k) * bval(k, n));
				if (h[m * 128 + n] != ref) {
					if (bad < 4) printf("  case %d mismatch (%d,%d): got %g want %g\n", mode, m, n, h[m * 128 + n], ref);
					++bad;
				}
			}
		printf("case %d: %s (%d mismatches of %d)\n", mode, bad ? "FAIL" : "ok", bad, 128 * N);
		if (mode <= 4) bad_total += bad;
	}
	printf(bad_total ? "PROBE FAILED\n" : "PROBE OK\n");
	return 0;
}
