// Inner CP-ALS on the projected tensor of one chromosome (reference: parafac_integrative.py:12-112)
// and the core norm (parafac2_intergrative.py:623-632). The big contractions (two MTTKRP GEMMs per
// iteration) go through fh_gemm_batched; everything r x r (Hadamard of Grams, ridge, SPD inverse,
// column norms, balance) is fused into small single-CTA kernels in fp64.
#include "fh_common.cuh"
#include "../../include/fh_b200.h"
#include <math.h>

namespace {

// H = G1 * G2 (Hadamard) + ridge I, inverted in place in shared memory by the symmetric sweep
// operator (Gauss-Jordan without pivoting: H is SPD), fp64. One CTA.
__global__ void __launch_bounds__(1024)
hadamard_inverse_kernel(const double* __restrict__ G1, const double* __restrict__ G2, int r, double ridge,
                        double* __restrict__ Hinv) {
	extern __shared__ double H[];  // r x ld
	const int ld = r | 1;
	double* colk = H + (size_t)r * ld;  // r: column k before the step
	const int tid = threadIdx.x, nt = blockDim.x;
	for (int i = tid; i < r * r; i += nt) {
		int a = i / r, b = i - a * r;
		H[a * ld + b] = G1[i] * G2[i] + (a == b ? ridge : 0.0);
	}
	__syncthreads();
	for (int k = 0; k < r; ++k) {
		const double d = H[k * ld + k];
		for (int i = tid; i < r; i += nt) colk[i] = H[i * ld + k];
		__syncthreads();
		const double inv = 1.0 / d;
		// row k <- row k / d
		for (int j = tid; j < r; j += nt) H[k * ld + j] = (j == k) ? inv : H[k * ld + j] * inv;
		__syncthreads();
		// rows i != k: a_ij -= a_ik * a_kj (j != k); a_ik = -a_ik / d
		for (int i = tid; i < r * r; i += nt) {
			int a = i / r, b = i - a * r;
			if (a == k) continue;
			double f = colk[a];
			H[a * ld + b] = (b == k) ? -f * inv : H[a * ld + b] - f * H[k * ld + b];
		}
		__syncthreads();
	}
	for (int i = tid; i < r * r; i += nt) {
		int a = i / r, b = i - a * r;
		Hinv[i] = H[a * ld + b];
	}
}

// column 2-norms of A (n x r), B (r x r), D (R x r): grid 3, thread per column. fp32 result like
// torch.norm on fp32 tensors.
__global__ void __launch_bounds__(256)
col_norms_kernel(const float* __restrict__ A, int n, const float* __restrict__ B, const float* __restrict__ D,
                 int R, int r, float* __restrict__ norms /* 3 x r */) {
	const float* F = blockIdx.x == 0 ? A : (blockIdx.x == 1 ? B : D);
	const int rows = blockIdx.x == 0 ? n : (blockIdx.x == 1 ? r : R);
	for (int c = threadIdx.x; c < r; c += blockDim.x) {
		double s = 0.0;
		for (int i = 0; i < rows; ++i) {
			double v = F[(size_t)i * r + c];
			s += v * v;
		}
		norms[blockIdx.x * r + c] = (float)sqrt(s);
	}
}

// balance_norm (parafac_integrative.py:19-26): unit columns; product of the norms into D
__global__ void balance_kernel(float* __restrict__ A, int n, float* __restrict__ B, float* __restrict__ D,
                               int R, int r, const float* __restrict__ norms, const int* __restrict__ stopped) {
	if (stopped && *stopped) return;  // the reference leaves its loop before this step once the loss has stalled
	const long long total = (long long)(n + r + R) * r;
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
	     i += (long long)gridDim.x * blockDim.x) {
		long long row = i / r;
		int c = (int)(i - row * r);
		float na = norms[c], nb = norms[r + c], nd = norms[2 * r + c];
		if (row < n) A[i] = A[i] / (na + 1e-15f);
		else if (row < n + r) { long long j = i - (long long)n * r; B[j] = B[j] / (nb + 1e-15f); }
		else {
			long long j = i - (long long)(n + r) * r;
			float prod = (na * nb) * nd;
			D[j] = (D[j] / (nd + 1e-15f)) * (prod + 1e-15f);
		}
	}
}

// Early stop without a host round trip (parafac_integrative.py:98-110: break once the relative loss change is < 1e-5).
// Every factor update is computed into a scratch buffer and COMMITTED only while the device-side flag is clear; the
// flag is raised by cp_stop_kernel after the commits of the stopping iteration, so the factors end exactly where the
// reference's loop leaves them and the later (speculative) iterations change nothing. Round 1 read the loss back
// after every iteration: with n_iter_parafac > 1 that serialised the chromosomes' streams on the host (headline job:
// 173 ms of a 447 ms sweep at six inner iterations).
__global__ void cp_commit_kernel(float* __restrict__ dst, const float* __restrict__ src, long long count,
                                 const int* __restrict__ stopped) {
	if (*stopped) return;
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
}
// scal: [0] ||Y||^2, [1] <Y, Xhat>, [2] ||Xhat||^2, [3] previous loss, [6] != 0 once [3] is valid (and was >= 0, as the
// reference's `prev_loss >= 0` guard), [4] / [5] the pair reported to the caller
__global__ void cp_stop_kernel(double* __restrict__ scal, int* __restrict__ stopped) {
	if (*stopped) return;
	const double loss = scal[0] - 2.0 * scal[1] + scal[2];
	scal[4] = scal[2]; scal[5] = scal[1];
	if (scal[6] != 0.0 && (scal[3] - loss) / scal[3] < 1e-5) *stopped = 1;
	scal[3] = loss;
	scal[6] = loss >= 0.0 ? 1.0 : 0.0;
}

// M0[i,p] = sum_j Z[(i*r+j), p] * B[j,p]
__global__ void mode0_reduce_kernel(const float* __restrict__ Z, const float* __restrict__ B, int n, int r,
                                    float* __restrict__ M0) {
	const int i = blockIdx.x;
	for (int p = threadIdx.x; p < r; p += blockDim.x) {
		float s = 0.f;
		for (int j = 0; j < r; ++j) s += Z[((size_t)i * r + j) * r + p] * B[j * r + p];
		M0[(size_t)i * r + p] = s;
	}
}
// M1[j,p] = sum_i Z[(i*r+j), p] * A[i,p]
__global__ void mode1_reduce_kernel(const float* __restrict__ Z, const float* __restrict__ A, int n, int r,
                                    float* __restrict__ M1) {
	const int j = blockIdx.x;
	for (int p = threadIdx.x; p < r; p += blockDim.x) {
		double s = 0.0;
		for (int i = 0; i < n; ++i) s += (double)Z[((size_t)i * r + j) * r + p] * (double)A[(size_t)i * r + p];
		M1[(size_t)j * r + p] = (float)s;
	}
}
// KR[(i*r + j), p] = A[i,p] * B[j,p]; KR rows at pitch ldk (a multiple of 4 floats: the tcgen05 GEMM reads them through TMA)
__global__ void khatri_rao_kernel(const float* __restrict__ A, const float* __restrict__ B, int n, int r, int ldk,
                                  float* __restrict__ KR) {
	const long long total = (long long)n * r * r;
	for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
	     t += (long long)gridDim.x * blockDim.x) {
		int p = (int)(t % r);
		long long ij = t / r;
		int j = (int)(ij % r);
		long long i = ij / r;
		KR[ij * ldk + p] = A[i * r + p] * B[j * r + p];
	}
}
// dst (rows x ldd) = src (rows x r), pad columns zero
__global__ void pad_rows_kernel(const float* __restrict__ src, int rows, int r, int ldd, float* __restrict__ dst) {
	const long long total = (long long)rows * ldd;
	for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
		const int p = (int)(t % ldd);
		dst[t] = p < r ? src[(t / ldd) * r + p] : 0.f;
	}
}
// acc += sum(G1 * G2 * G3)
__global__ void __launch_bounds__(1024)
triple_hadamard_sum_kernel(const double* __restrict__ G1, const double* __restrict__ G2,
                           const double* __restrict__ G3, int nn, double* __restrict__ acc) {
	__shared__ double red[32];
	double s = 0.0;
	for (int i = threadIdx.x; i < nn; i += blockDim.x) s += G1[i] * G2[i] * G3[i];
	s = fh_block_sum(s, red);
	if (threadIdx.x == 0) *acc += s;
}

int gemm(int dtype, int M, int N, int K, const void* A, long long sa_m, long long sa_k, const void* B,
         long long sb_k, long long sb_n, void* C, long long ldc, void* stream) {
	fh_gemm_desc g;
	memset(&g, 0, sizeof(g));
	g.M = M; g.N = N; g.K = K; g.batch = 1;
	g.sa_m = sa_m; g.sa_k = sa_k; g.sb_k = sb_k; g.sb_n = sb_n; g.ldc = ldc;
	g.alpha = 1.0; g.dtype = dtype;
	return fh_gemm_batched(&g, A, B, C, stream);
}
// G = F^T F in fp64 (F rows x r, fp32)
int gram(const float* F, int rows, int r, double* G, void* stream) {
	return gemm(FH_GEMM_F32_ACC64, r, r, rows, F, 1, r, F, r, 1, G, r, stream);
}

// out[i][j][p] = F[j][p] * Arows[i][p]  (p < r; pad columns r..ldo-1 are 0): the per-bin scaled copies
// B diag(A_i), D diag(A_i) that turn temp_i = T1_i diag(A_i) B^T and W_i = (U_i B) diag(A_i) D^T into plain
// batched GEMMs for the tensor-core kernel.
__global__ void scale_cols_batched_kernel(const float* __restrict__ F, int rows, int r, long long ldf,
                                          const float* __restrict__ Arows, int nb, int ldo, float* __restrict__ out) {
	const long long total = (long long)nb * rows * ldo;
	for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
		const int p = (int)(t % ldo);
		const long long ij = t / ldo;
		const int j = (int)(ij % rows);
		const long long i = ij / rows;
		out[t] = (p < r) ? F[(long long)j * ldf + p] * Arows[i * r + p] : 0.f;
	}
}

size_t al(size_t x) { return (x + 255) / 256 * 256; }
struct CpWs {
	float *Z, *M, *Tm, *Dp, *norms;
	double *Ga, *Gb, *Gd, *Hinv, *scal;
	size_t bytes;
};
CpWs carve(int n, int r, int R, void* ws) {
	CpWs c;
	char* b = (char*)ws;
	const int rp = (r + 3) & ~3;
	c.Z = (float*)b; b += al((size_t)n * r * rp * 4);  // Y D (pitch r), then KhatriRao(A, B) (pitch rp)
	c.Dp = (float*)b; b += al((size_t)R * rp * 4);     // D at a TMA-describable pitch
	int mx = n > R ? n : R; mx = mx > r ? mx : r;
	c.M = (float*)b; b += al((size_t)mx * r * 4);
	c.Tm = (float*)b; b += al((size_t)mx * r * 4);  // factor update before its commit
	c.norms = (float*)b; b += al((size_t)3 * r * 4);
	c.Ga = (double*)b; b += al((size_t)r * r * 8);
	c.Gb = (double*)b; b += al((size_t)r * r * 8);
	c.Gd = (double*)b; b += al((size_t)r * r * 8);
	c.Hinv = (double*)b; b += al((size_t)r * r * 8);
	c.scal = (double*)b; b += 256;
	c.bytes = (size_t)(b - (char*)ws);
	return c;
}

int balance(float* A, int n, float* B, float* D, int R, int r, float* norms, const int* stopped, cudaStream_t st) {
	col_norms_kernel<<<3, 256, 0, st>>>(A, n, B, D, R, r, norms);
	FH_LAUNCH_CHECK();
	long long total = (long long)(n + r + R) * r;
	balance_kernel<<<fh_cdiv(total, 256), 256, 0, st>>>(A, n, B, D, R, r, norms, stopped);
	FH_LAUNCH_CHECK();
	return FH_OK;
}

// F = M * (G1 * G2 + ridge I)^{-1}
int solve_update(const float* M, int rows, const double* G1, const double* G2, int r, double* Hinv, float* tmp, float* out,
                 const int* stopped, cudaStream_t st) {
	size_t smem = ((size_t)r * (r | 1) + r) * 8;
	FH_CHECK_ARG(smem <= 227 * 1024, "fh_cp_als: r=%d too large for the shared-memory SPD inverse", r);
	FH_CUDA(cudaFuncSetAttribute(hadamard_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	int threads = r * r >= 1024 ? 1024 : ((r * r + 31) / 32 * 32);
	hadamard_inverse_kernel<<<1, threads, smem, st>>>(G1, G2, r, 1e-10, Hinv);
	FH_LAUNCH_CHECK();
	int rc = gemm(FH_GEMM_F32xF64_F32, rows, r, r, M, r, 1, Hinv, r, 1, tmp, r, (void*)st);
	if (rc) return rc;
	const long long count = (long long)rows * r;
	cp_commit_kernel<<<fh_cdiv(count, 256) > 592 ? 592 : fh_cdiv(count, 256), 256, 0, st>>>(out, tmp, count, stopped);
	FH_LAUNCH_CHECK();
	return FH_OK;
}

}  // namespace

extern "C" size_t fh_cp_als_workspace_bytes(int n, int r, int R) { return carve(n, r, R, nullptr).bytes; }

extern "C" int fh_cp_als(const float* Y, int n, int r, int R, float* A, float* B, float* D, int n_iter_max,
                         void* workspace, size_t workspace_bytes, double* host_out, void* stream) {
	FH_CHECK_ARG(n > 0 && r > 0 && R > 0, "fh_cp_als: bad shape");
	FH_CHECK_ARG(workspace && workspace_bytes >= fh_cp_als_workspace_bytes(n, r, R), "fh_cp_als: workspace too small");
	cudaStream_t st = (cudaStream_t)stream;
	CpWs w = carve(n, r, R, workspace);
	const int rp = (r + 3) & ~3;
	int rc;
	if (host_out) host_out[0] = host_out[1] = 0.0;
	rc = balance(A, n, B, D, R, r, w.norms, nullptr, st);
	if (rc) return rc;
	const bool need_loss = n_iter_max > 1;
	int* stopped = (int*)(w.scal + 8);
	FH_CUDA(cudaMemsetAsync(w.scal, 0, 128, st));
	if (need_loss) {
		rc = fh_sqnorm_accum(Y, 1, (long long)n * r * R, (long long)n * r * R, w.scal, stream);
		if (rc) return rc;
	}
	for (int it = 0; it < n_iter_max; ++it) {
		// Z = Y_(n r x R) D : shared by the A and B updates (D is unchanged between them). Both MTTKRP products run on the
		// tcgen05 3xTF32 kernel (round 1: CUDA-core fp32, 18 of the CP-ALS stage's launch time): D and the Khatri-Rao
		// matrix are read at the padded pitch rp; operands TMA cannot describe (R not a multiple of 4) fall back by themselves.
		pad_rows_kernel<<<fh_cdiv((long long)R * rp, 256), 256, 0, st>>>(D, R, r, rp, w.Dp);
		FH_LAUNCH_CHECK();
		rc = gemm(FH_GEMM_TF32X3, n * r, r, R, Y, R, 1, w.Dp, rp, 1, w.Z, r, stream);
		if (rc) return rc;
		// mode 0: A = M0 ((B^T B) * (D^T D))^{-1}
		if ((rc = gram(B, r, r, w.Gb, stream))) return rc;
		if ((rc = gram(D, R, r, w.Gd, stream))) return rc;
		mode0_reduce_kernel<<<n, 128, 0, st>>>(w.Z, B, n, r, w.M);
		FH_LAUNCH_CHECK();
		if ((rc = solve_update(w.M, n, w.Gb, w.Gd, r, w.Hinv, w.Tm, A, stopped, st))) return rc;
		// mode 1: B = M1 ((A^T A) * (D^T D))^{-1}
		if ((rc = gram(A, n, r, w.Ga, stream))) return rc;
		mode1_reduce_kernel<<<r, 128, 0, st>>>(w.Z, A, n, r, w.M);
		FH_LAUNCH_CHECK();
		if ((rc = solve_update(w.M, r, w.Ga, w.Gd, r, w.Hinv, w.Tm, B, stopped, st))) return rc;
		// mode 2: D = M2 ((A^T A) * (B^T B))^{-1},  M2 = Y_(R x n r) KhatriRao(A, B)
		if ((rc = gram(B, r, r, w.Gb, stream))) return rc;
		khatri_rao_kernel<<<fh_cdiv((long long)n * r * r, 256) > 4096 ? 4096 : fh_cdiv((long long)n * r * r, 256), 256, 0, st>>>(A, B, n, r, rp, w.Z);
		FH_LAUNCH_CHECK();
		rc = gemm(FH_GEMM_TF32X3, R, r, n * r, Y, 1, R, w.Z, rp, 1, w.M, r, stream);
		if (rc) return rc;
		if ((rc = solve_update(w.M, R, w.Ga, w.Gb, r, w.Hinv, w.Tm, D, stopped, st))) return rc;
		if (need_loss) {
			// loss = ||Y||^2 - 2 <Y, Xhat> + ||Xhat||^2 ; <Y, Xhat> = <M2, D>; decided on the device (cp_stop_kernel)
			FH_CUDA(cudaMemsetAsync(w.scal + 1, 0, 16, st));
			if ((rc = fh_dot_accum(w.M, D, R, r, r, r, w.scal + 1, stream))) return rc;
			if ((rc = gram(D, R, r, w.Gd, stream))) return rc;
			triple_hadamard_sum_kernel<<<1, 1024, 0, st>>>(w.Ga, w.Gb, w.Gd, r * r, w.scal + 2);
			FH_LAUNCH_CHECK();
			cp_stop_kernel<<<1, 1, 0, st>>>(w.scal, stopped);
			FH_LAUNCH_CHECK();
		}
		rc = balance(A, n, B, D, R, r, w.norms, stopped, st);
		if (rc) return rc;
	}
	if (host_out && need_loss) {  // only the API-compatible `parafac` entry asks for these: one read-back at the end
		double h[2];
		FH_CUDA(cudaMemcpyAsync(h, w.scal + 4, 16, cudaMemcpyDeviceToHost, st));
		FH_CUDA(cudaStreamSynchronize(st));
		host_out[0] = h[0]; host_out[1] = h[1];
	}
	return FH_OK;
}

extern "C" int fh_cp_core_sqnorm(const float* A, int n, const float* B, const float* D, int R, int r,
                                 double* ws, double* acc, void* stream) {
	FH_CHECK_ARG(ws && acc, "fh_cp_core_sqnorm: null workspace");
	int rc;
	double* Ga = ws; double* Gb = ws + (size_t)r * r; double* Gd = Gb + (size_t)r * r;
	if ((rc = gram(A, n, r, Ga, stream))) return rc;
	if ((rc = gram(B, r, r, Gb, stream))) return rc;
	if ((rc = gram(D, R, r, Gd, stream))) return rc;
	triple_hadamard_sum_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(Ga, Gb, Gd, r * r, acc);
	FH_LAUNCH_CHECK();
	return FH_OK;
}

extern "C" int fh_scale_cols_batched(const float* F, int rows, int r, long long ldf, const float* Arows, int nb, int ldo,
                                     float* out, void* stream) {
	FH_CHECK_ARG(rows > 0 && r > 0 && nb > 0 && ldo >= r && ldf >= r, "fh_scale_cols_batched: bad shape");
	long long total = (long long)nb * rows * ldo;
	int grid = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
	scale_cols_batched_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(F, rows, r, ldf, Arows, nb, ldo, out);
	FH_LAUNCH_CHECK();
	return FH_OK;
}
