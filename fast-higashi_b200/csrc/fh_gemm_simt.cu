// Generic batched strided GEMM on the CUDA cores (fp32 / fp64 FFMA-DFMA), used for
//   * every fp64 contraction of the polar step (Gram, inverse-square-root apply) - fp64 is
//     required there for parity (DESIGN.md "polar"), tensor cores do not apply;
//   * the small per-bin batched products (temp_i, W_i, Y_i) and the CP-ALS MTTKRPs;
//   * the fp32 fallback/reference for the tcgen05 3xTF32 kernel (fh_gemm_tc.cu).
// C[b](m,n) = alpha * sum_k A[b](m,k) * kscale[b][k] * B[b](k,n)  (+ epilogue) (+ beta * C[b](m,n))
#include "fh_common.cuh"
#include "../../include/fh_b200.h"

namespace {

struct GemmP {
	int M, N, K, batch;
	long long sa_m, sa_k, sb_k, sb_n, ldc;
	long long batch_a, batch_b, batch_c;
	double alpha, beta;
	int epilogue;
	double diag;
	const float* kscale; long long kscale_batch;
	const float* cscale; long long cscale_batch; int cscale_recip;
	int splits, kchunk;  // split-K: grid.z = batch * splits, partial sums combined with atomicAdd
};

template <typename TA, typename TB, typename TC, typename TAcc, int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(GemmP p, const TA* __restrict__ A, const TB* __restrict__ B, TC* __restrict__ C) {
	constexpr int NT = 256;
	static_assert((BM / TM) * (BN / TN) == NT, "tile/thread mismatch");
	static_assert(BM == BN, "FH_EPI_SYMMETRIC mirrors square tiles");
	constexpr int PA = 4, PB = 4;
	__shared__ __align__(16) TAcc As[BK][BM + PA];
	__shared__ __align__(16) TAcc Bs[BK][BN + PB];
	const int b = blockIdx.z / p.splits;
	const int kbeg = (blockIdx.z % p.splits) * p.kchunk;
	const int kend = min(p.K, kbeg + p.kchunk);
	const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
	const bool sym = p.epilogue == FH_EPI_SYMMETRIC;  // square tiles (BM == BN): tiles below the diagonal are mirror images
	if (sym && blockIdx.x < blockIdx.y) return;
	const int tid = threadIdx.x;
	const int tx = tid % (BN / TN), ty = tid / (BN / TN);
	A += (long long)b * p.batch_a;
	B += (long long)b * p.batch_b;
	C += (long long)b * p.batch_c;
	const float* ks = p.kscale ? p.kscale + (long long)b * p.kscale_batch : nullptr;
	const bool a_kc = (p.sa_k == 1), b_nc = (p.sb_n == 1);
	TAcc acc[TM][TN];
#pragma unroll
	for (int i = 0; i < TM; ++i)
#pragma unroll
		for (int j = 0; j < TN; ++j) acc[i][j] = TAcc(0);

	for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
		for (int it = 0; it < (BM * BK + NT - 1) / NT; ++it) {
			int idx = tid + it * NT;
			if ((BM * BK) % NT != 0 && idx >= BM * BK) break;
			int kk, mm;
			if (a_kc) { kk = idx % BK; mm = idx / BK; } else { mm = idx % BM; kk = idx / BM; }
			int m = m0 + mm, k = k0 + kk;
			TAcc v = TAcc(0);
			if (m < p.M && k < kend) {
				v = (TAcc)A[(long long)m * p.sa_m + (long long)k * p.sa_k];
				if (ks) v *= (TAcc)ks[k];
			}
			As[kk][mm] = v;
		}
#pragma unroll
		for (int it = 0; it < (BN * BK + NT - 1) / NT; ++it) {
			int idx = tid + it * NT;
			if ((BN * BK) % NT != 0 && idx >= BN * BK) break;
			int kk, nn;
			if (b_nc) { nn = idx % BN; kk = idx / BN; } else { kk = idx % BK; nn = idx / BK; }
			int n = n0 + nn, k = k0 + kk;
			TAcc v = TAcc(0);
			if (n < p.N && k < kend) v = (TAcc)B[(long long)k * p.sb_k + (long long)n * p.sb_n];
			Bs[kk][nn] = v;
		}
		__syncthreads();
#pragma unroll
		for (int kk = 0; kk < BK; ++kk) {
			TAcc a[TM], bb[TN];
#pragma unroll
			for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
			for (int j = 0; j < TN; ++j) bb[j] = Bs[kk][tx * TN + j];
#pragma unroll
			for (int i = 0; i < TM; ++i)
#pragma unroll
				for (int j = 0; j < TN; ++j) acc[i][j] += a[i] * bb[j];
		}
		__syncthreads();
	}
	const float* cs = p.cscale ? p.cscale + (long long)b * p.cscale_batch : nullptr;
#pragma unroll
	for (int i = 0; i < TM; ++i) {
		int m = m0 + ty * TM + i;
		if (m >= p.M) continue;
#pragma unroll
		for (int j = 0; j < TN; ++j) {
			int n = n0 + tx * TN + j;
			if (n >= p.N) continue;
			TAcc v = (TAcc)p.alpha * acc[i][j];
			if (p.epilogue == FH_EPI_DIAG_ADD && m == n) v += (TAcc)p.diag;
			if (cs) {
				TAcc s = (TAcc)cs[n];
				v = p.cscale_recip ? v / s : v * s;
			}
			long long off = (long long)m * p.ldc + n;
			if (p.splits > 1) { atomicAdd(&C[off], (TC)v); continue; }
			if (p.beta != 0.0) v += (TAcc)p.beta * (TAcc)C[off];
			C[off] = (TC)v;
			if (sym && blockIdx.x > blockIdx.y) C[(long long)n * p.ldc + m] = (TC)v;
		}
	}
}

template <typename TA, typename TB, typename TC, typename TAcc>
int launch(GemmP p, const void* A, const void* B, void* C, cudaStream_t st) {
	if (p.M <= 0 || p.N <= 0 || p.batch <= 0) return FH_OK;
	constexpr bool dbl = sizeof(TAcc) == 8;
	long long t128 = (long long)fh_cdiv(p.M, 128) * fh_cdiv(p.N, 128) * p.batch;
	long long t64 = (long long)fh_cdiv(p.M, 64) * fh_cdiv(p.N, 64) * p.batch;
	p.splits = 1; p.kchunk = p.K;
	// long-K problems with few output tiles (mode-2 MTTKRP, Grams): split K over the grid
	if (t64 < 148 && p.K >= 2048 && p.beta == 0.0 && p.epilogue == FH_EPI_NONE && !p.cscale && p.batch <= 8) {
		int s = (int)(296 / t64);
		int smax = p.K / 512;
		s = s > smax ? smax : s;
		if (s > 1) {
			p.kchunk = (fh_cdiv(p.K, s) + 15) / 16 * 16;
			p.splits = fh_cdiv(p.K, p.kchunk);
			for (int b = 0; b < p.batch; ++b)
				FH_CUDA(cudaMemset2DAsync((TC*)C + (long long)b * p.batch_c, (size_t)p.ldc * sizeof(TC), 0,
				                          (size_t)p.N * sizeof(TC), (size_t)p.M, st));
		}
	}
	if (!dbl && t128 >= 296) {
		dim3 g(fh_cdiv(p.N, 128), fh_cdiv(p.M, 128), p.batch);
		gemm_simt_kernel<TA, TB, TC, TAcc, 128, 128, 8, 8, 8><<<g, 256, 0, st>>>(p, (const TA*)A, (const TB*)B, (TC*)C);
	} else {
		// 64 x 64 or 48 x 48 tiles, whichever pads the problem less (r = 137 -> 3 x 48 = 144 instead of 3 x 64 = 192)
		const long long pad64 = (long long)fh_cdiv(p.M, 64) * 64 * fh_cdiv(p.N, 64) * 64;
		const long long pad48 = (long long)fh_cdiv(p.M, 48) * 48 * fh_cdiv(p.N, 48) * 48;
		if (dbl && pad48 * 10 < pad64 * 9) {
			dim3 g(fh_cdiv(p.N, 48), fh_cdiv(p.M, 48), p.batch * p.splits);
			gemm_simt_kernel<TA, TB, TC, TAcc, 48, 48, 8, 3, 3><<<g, 256, 0, st>>>(p, (const TA*)A, (const TB*)B, (TC*)C);
		} else {
			dim3 g(fh_cdiv(p.N, 64), fh_cdiv(p.M, 64), p.batch * p.splits);
			if (dbl)
				gemm_simt_kernel<TA, TB, TC, TAcc, 64, 64, 8, 4, 4><<<g, 256, 0, st>>>(p, (const TA*)A, (const TB*)B, (TC*)C);
			else
				gemm_simt_kernel<TA, TB, TC, TAcc, 64, 64, 16, 4, 4><<<g, 256, 0, st>>>(p, (const TA*)A, (const TB*)B, (TC*)C);
		}
	}
	FH_LAUNCH_CHECK();
	return FH_OK;
}

}  // namespace

int fh_gemm_tc(const fh_gemm_desc* d, const float* A, const float* B, float* C, void* stream);
extern "C" void fh_count_tc_fallback(void);

extern "C" int fh_gemm_batched(const fh_gemm_desc* d, const void* A, const void* B, void* C, void* stream) {
	FH_CHECK_ARG(d != nullptr, "fh_gemm_batched: null descriptor");
	FH_CHECK_ARG(d->M >= 0 && d->N >= 0 && d->K >= 0 && d->batch >= 0, "fh_gemm_batched: negative size");
	FH_CHECK_ARG(d->batch <= 32768, "fh_gemm_batched: batch %d > 32768 (split the call)", d->batch);
	FH_CHECK_ARG(d->sa_m == 1 || d->sa_k == 1, "fh_gemm_batched: A needs a unit stride");
	FH_CHECK_ARG(d->sb_k == 1 || d->sb_n == 1, "fh_gemm_batched: B needs a unit stride");
	GemmP p;
	p.M = d->M; p.N = d->N; p.K = d->K; p.batch = d->batch;
	p.sa_m = d->sa_m; p.sa_k = d->sa_k; p.sb_k = d->sb_k; p.sb_n = d->sb_n; p.ldc = d->ldc;
	p.batch_a = d->batch_a; p.batch_b = d->batch_b; p.batch_c = d->batch_c;
	p.alpha = d->alpha; p.beta = d->beta; p.epilogue = d->epilogue; p.diag = d->diag;
	p.kscale = d->kscale; p.kscale_batch = d->kscale_batch;
	p.cscale = d->cscale; p.cscale_batch = d->cscale_batch; p.cscale_recip = d->cscale_recip;
	// FH_EPI_SYMMETRIC only where it is implemented and meaningful; otherwise the full product
	if (p.epilogue == FH_EPI_SYMMETRIC && (d->M != d->N || d->beta != 0.0 || d->cscale || d->dtype == FH_GEMM_TF32X3)) p.epilogue = FH_EPI_NONE;
	cudaStream_t st = (cudaStream_t)stream;
	switch (d->dtype) {
		case FH_GEMM_F32: return launch<float, float, float, float>(p, A, B, C, st);
		case FH_GEMM_F32_ACC64: return launch<float, float, double, double>(p, A, B, C, st);
		case FH_GEMM_F64: return launch<double, double, double, double>(p, A, B, C, st);
		case FH_GEMM_F32xF64_F32: return launch<float, double, float, double>(p, A, B, C, st);
		case FH_GEMM_F64xF32_F32: return launch<double, float, float, double>(p, A, B, C, st);
		case FH_GEMM_TF32X3: {
			// tensor-core path; operands TMA cannot describe (odd strides, k-scaling) run on the CUDA-core
			// fp32 kernel instead - still on the GPU, still exact fp32, counted in fh_tc_fallback_count()
			int rc = fh_gemm_tc(d, (const float*)A, (const float*)B, (float*)C, stream);
			if (rc != FH_ERR_UNSUPPORTED) return rc;
			fh_count_tc_fallback();
			return launch<float, float, float, float>(p, A, B, C, st);
		}
		default: break;
	}
	fh_set_error("fh_gemm_batched: unknown dtype %d", d->dtype);
	return FH_ERR_ARG;
}
