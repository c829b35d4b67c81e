// Batched polar factor (reference: project2orthogonal.py:6-55, called at
// parafac2_intergrative.py:396 with one (w x r) matrix per bin).
//
// U = T (T^T T)^{-1/2}. The per-bin matrices have condition numbers 1e4..2e7, so the
// Gram route only reaches parity in fp64 (measured: fp32 Gram -> 15 % loss error, fp64 Gram -> 3e-6;
// DESIGN.md "polar"). Pipeline, batched over the bins of one block, all fp64:
//   G = T^T T                          fp64 accumulate      (fh_gemm_batched, FH_GEMM_F32_ACC64)
//   G = P L L^T P^T                    diagonally pivoted Cholesky, G resident in shared memory
//   L V = W, columns of W orthogonal   one-sided (Hestenes) Jacobi on the columns of L, register-blocked
//                                      (fh_polar_rb.cuh: a warp per pair of 4-row blocks, odd-even block ordering).
//                                      (Veselic-Hari: L^T L is far closer to diagonal than L L^T, so
//                                      the strongly graded spectra converge in <= 9 sweeps where Jacobi
//                                      on G itself needed 14-24, measured.) lambda_j = |w_j|^2.
//   M = sum_j w_j w_j^T lambda_j^{-3/2}  = G^{-1/2}          (FH_GEMM_F64)
//   U = T M                            fp64 accumulate, fp32 out
#include "fh_common.cuh"
#include "../../include/fh_b200.h"
#include <math.h>

namespace {

constexpr int kMaxGram = 160;   // shared memory: 160 x 161 fp64 = 206 KB

#include "fh_polar_rb.cuh"     // register-blocked one-sided Jacobi


// ---------------------------------------------------------------------------------------------
// One CTA per matrix (problem b: size prob_n[b], data at prob_off[b] doubles into Gall / WTall). Shared memory holds R
// (rb_rows(n) x ld, row-major, pad rows / columns zero): first the symmetric G, then its pivoted Cholesky factor as the
// UPPER triangle R = L^T (row j of R = column j of L), then the rows are orthogonalised in place (fh_polar_rb.cuh).
// Output WT (n x n): row j = w_j * lambda_j^{-3/4} in the ORIGINAL index order; sigma_all[b n + j] = sqrt(lambda_j);
// sigma_sum[prob_slot[b]] = sum_j sqrt(lambda_j). One instantiation per row length class PL = ceil(n / 32) so that each
// gets its own register budget and CTA count per SM.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pivoted_cholesky_upper(double* __restrict__ R, const int n, const int ld, int* __restrict__ perm) {
	__shared__ int s_piv;
	__shared__ double s_val;
	const int JT = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	double dmax0 = 0.0;
	for (int k = 0; k < n; ++k) {
		if (warp == 0) {  // pivot = largest remaining diagonal
			double best = -1.0; int bi = k;
			for (int i = k + lane; i < n; i += 32) {
				double v = R[i * ld + i];
				if (v > best) { best = v; bi = i; }
			}
			for (int o = 16; o > 0; o >>= 1) {
				double ov = __shfl_xor_sync(0xffffffffu, best, o);
				int oi = __shfl_xor_sync(0xffffffffu, bi, o);
				if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
			}
			if (lane == 0) { s_piv = bi; s_val = best; }
		}
		__syncthreads();
		const int pv = s_piv;
		const double pval = s_val;
		if (k == 0) dmax0 = pval;
		// rank-revealing stop: once the largest remaining diagonal is below the rounding level of the fp64 Gram
		// (~n eps lambda_max) the trailing block is noise. It is replaced by thr * I, i.e. singular values
		// ~1e-7 sigma_max in directions T has no resolvable energy in.
		if (pval <= dmax0 * 1e-14 || !(pval > 0.0)) {
			const double rt = sqrt(fmax(dmax0 * 1e-14, 1e-300));
			const int rem0 = n - k;
			__syncthreads();
			for (int t = tid; t < rem0 * rem0; t += JT) {
				int i = k + t / rem0, j = k + t % rem0;
				R[i * ld + j] = (i == j) ? rt : 0.0;
			}
			__syncthreads();
			break;
		}
		if (pv != k) {  // symmetric swap k <-> pv: rows, then columns (earlier factor rows included)
			for (int j = tid; j < n; j += JT) { double t = R[k * ld + j]; R[k * ld + j] = R[pv * ld + j]; R[pv * ld + j] = t; }
			__syncthreads();
			for (int i = tid; i < n; i += JT) { double t = R[i * ld + k]; R[i * ld + k] = R[i * ld + pv]; R[i * ld + pv] = t; }
			if (tid == 0) { int t = perm[k]; perm[k] = perm[pv]; perm[pv] = t; }
			__syncthreads();
		}
		const double rkk = sqrt(R[k * ld + k]);
		const double rinv = 1.0 / rkk;
		__syncthreads();
		for (int j = k + tid; j < n; j += JT) R[k * ld + j] = (j == k) ? rkk : R[k * ld + j] * rinv;
		__syncthreads();
		// trailing update, a warp per row (full square keeps the swaps simple): G[i][j] -= R[k][i] R[k][j]
		const int nwp = JT >> 5;
		for (int i = k + 1 + warp; i < n; i += nwp) {
			const double rki = R[k * ld + i];
			for (int j = k + 1 + lane; j < n; j += 32) R[i * ld + j] = fma(-rki, R[k * ld + j], R[i * ld + j]);
		}
		__syncthreads();
	}
	for (int t = tid; t < n * n; t += JT) {  // strict lower triangle of R is not part of the factor
		int i = t / n, j = t % n;
		if (j < i) R[i * ld + j] = 0.0;
	}
	__syncthreads();
}

template <int PL>
__global__ void __launch_bounds__(PL == 5 ? 640 : PL * 128, PL >= 3 ? 1 : (PL == 2 ? 3 : 6))
chol_jacobi_rb_kernel(const double* __restrict__ Gall, const int* __restrict__ prob_n, const long long* __restrict__ prob_off,
                      const int* __restrict__ prob_slot, int uniform_n, int max_sweeps, double skip_tol,
                      double* __restrict__ WTall, double* __restrict__ sigma_all, double* __restrict__ sigma_sum,
                      int* __restrict__ nsweep_out) {
	extern __shared__ __align__(16) double sm[];
	const int b = blockIdx.x;
	const int n = prob_n ? prob_n[b] : uniform_n;
	const long long off = prob_off ? prob_off[b] : (long long)b * n * n;
	const int slot = prob_slot ? prob_slot[b] : b;
	const int ld = PL * 32 + 1, nrow = rb_rows(n);
	double* R = sm;                          // nrow x ld
	double* red = R + (size_t)nrow * ld;     // 64 doubles scratch
	double* nrm = red + 64;                  // nrow squared norms
	int* perm = (int*)(nrm + nrow);          // n
	const int JT = blockDim.x;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = JT >> 5;
	const double* Gg = Gall + off;
	for (int r = warp; r < nrow; r += nw)
		for (int c = lane; c < ld; c += 32) R[r * ld + c] = (c < n && r < n) ? Gg[r * n + c] : 0.0;
	for (int i = tid; i < n; i += JT) perm[i] = i;
	__syncthreads();
	pivoted_cholesky_upper(R, n, ld, perm);
	const int sweep = (PL > 4) ? jacobi_sweeps_rb_half<PL>(R, n, ld, nrm, red, max_sweeps, skip_tol)
	                           : jacobi_sweeps_rb<PL>(R, n, ld, nrm, red, max_sweeps, skip_tol);
	if (tid == 0 && nsweep_out) nsweep_out[slot] = sweep;
	// ---------------- lambda_j = |w_j|^2, outputs ----------------
	__shared__ double s_lmax, s_ssum;
	if (tid == 0) { s_lmax = 0.0; s_ssum = 0.0; }
	__syncthreads();
	double* lamv = sigma_all ? sigma_all + (long long)b * n : nullptr;
	for (int j = warp; j < n; j += nw) {
		double s2 = 0.0;
		for (int i = lane; i < n; i += 32) { double v = R[j * ld + i]; s2 += v * v; }
		s2 = fh_warp_sum(s2);
		if (lane == 0) {
			nrm[j] = s2;
			atomicMax((unsigned long long*)&s_lmax, (unsigned long long)__double_as_longlong(s2));  // s2 >= 0: order preserved
			if (lamv) lamv[j] = sqrt(s2);
		}
	}
	__syncthreads();
	if (tid == 0 && sigma_sum) {  // fixed summation order: the value feeds the loss
		double t = 0.0;
		for (int j = 0; j < n; ++j) t += sqrt(nrm[j]);
		sigma_sum[slot] = t;
	}
	const double floor_l = fmax(s_lmax * 1e-17, 1e-300);
	double* WT = WTall + off;
	for (int j = warp; j < n; j += nw) {
		const double l = fmax(nrm[j], floor_l);
		const double f = rsqrt(l) * rsqrt(sqrt(l));  // lambda^{-3/4}
		for (int i = lane; i < n; i += 32) WT[(size_t)j * n + perm[i]] = R[j * ld + i] * f;
	}
}

int gemm(int dtype, int M, int N, int K, int batch, const void* A, long long sa_m, long long sa_k, long long ba,
         const void* B, long long sb_k, long long sb_n, long long bb, void* C, long long ldc, long long bc,
         void* stream) {
	fh_gemm_desc g;
	memset(&g, 0, sizeof(g));
	g.M = M; g.N = N; g.K = K; g.batch = batch;
	g.sa_m = sa_m; g.sa_k = sa_k; g.sb_k = sb_k; g.sb_n = sb_n; g.ldc = ldc;
	g.batch_a = ba; g.batch_b = bb; g.batch_c = bc;
	g.alpha = 1.0; g.beta = 0.0; g.dtype = dtype;
	return fh_gemm_batched(&g, A, B, C, stream);
}

size_t al(size_t x) { return (x + 255) / 256 * 256; }

struct PolarWs {
	double *G, *WT;
	int* nsweep;
	size_t bytes;
};
PolarWs carve(int batch, int n, void* ws) {
	PolarWs p;
	char* b = (char*)ws;
	size_t nn = al((size_t)batch * n * n * 8);
	p.G = (double*)b; b += nn;
	p.WT = (double*)b; b += nn;
	p.nsweep = (int*)b; b += al((size_t)batch * 4);
	p.bytes = (size_t)(b - (char*)ws);
	return p;
}

constexpr int kMaxSweeps = 30;

template <int PL>
int launch_rb(int grid, int nmax, size_t smem, cudaStream_t st, const double* G, const int* pn, const long long* po, const int* ps,
              int uniform_n, int max_sweeps, double skip, double* WT, double* sigma, double* sigma_sum, int* nsweep) {
	FH_CUDA(cudaFuncSetAttribute(chol_jacobi_rb_kernel<PL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	const int tmr = pn ? -1 : fh_time_begin(FH_TIME_POLAR_JACOBI, st);  // table-driven launches overlap on side streams: timed as one group by the caller
	chol_jacobi_rb_kernel<PL><<<grid, rb_threads(nmax), smem, st>>>(G, pn, po, ps, uniform_n, max_sweeps, skip, WT, sigma, sigma_sum, nsweep);
	fh_time_end(tmr, st);
	FH_LAUNCH_CHECK();
	return FH_OK;
}

int launch_jacobi(int grid, int nmax, size_t smem, cudaStream_t st, const double* G, const int* pn, const long long* po,
                  const int* ps, int uniform_n, int max_sweeps, double* WT, double* sigma, double* sigma_sum, int* nsweep) {
	static double skip = -1.0;
	if (skip < 0.0) { const char* e = getenv("FH_JACOBI_SKIP"); skip = e ? atof(e) : 1e-17; }
	switch (rb_pl(nmax)) {
		case 1: return launch_rb<1>(grid, nmax, smem, st, G, pn, po, ps, uniform_n, max_sweeps, skip, WT, sigma, sigma_sum, nsweep);
		case 2: return launch_rb<2>(grid, nmax, smem, st, G, pn, po, ps, uniform_n, max_sweeps, skip, WT, sigma, sigma_sum, nsweep);
		case 3: return launch_rb<3>(grid, nmax, smem, st, G, pn, po, ps, uniform_n, max_sweeps, skip, WT, sigma, sigma_sum, nsweep);
		case 4: return launch_rb<4>(grid, nmax, smem, st, G, pn, po, ps, uniform_n, max_sweeps, skip, WT, sigma, sigma_sum, nsweep);
		default: return launch_rb<5>(grid, nmax, smem, st, G, pn, po, ps, uniform_n, max_sweeps, skip, WT, sigma, sigma_sum, nsweep);
	}
}

// A launch covers problems of side <= n with one shared-memory size: the layout of the largest.
size_t jacobi_smem(int n) { return ((size_t)rb_rows(n) * rb_ld(n) + 64 + rb_rows(n) + (n + 1) / 2) * 8 + 16; }

}  // namespace

extern "C" size_t fh_polar_workspace_bytes(int batch, int rows, int cols) {
	int n = rows < cols ? rows : cols;
	if (batch <= 0 || n <= 0) return 0;
	return carve(batch, n, nullptr).bytes;
}

extern "C" int fh_polar_batched(const float* T, float* U, int batch, int rows, int cols, long long ld,
                                long long batch_stride, double* sigma_sum, double* sigma, int max_sweeps,
                                void* workspace, size_t workspace_bytes, int* host_max_sweeps, void* stream) {
	FH_CHECK_ARG(batch >= 0 && rows > 0 && cols > 0 && ld >= cols, "fh_polar_batched: bad shape");
	if (host_max_sweeps) *host_max_sweeps = 0;
	if (batch == 0) return FH_OK;
	FH_CHECK_ARG(batch <= 32768, "fh_polar_batched: batch > 32768");
	if (max_sweeps <= 0 || max_sweeps > kMaxSweeps) max_sweeps = kMaxSweeps;
	const bool tall = rows >= cols;
	const int n = tall ? cols : rows;
	const size_t smem = jacobi_smem(n);
	FH_CHECK_ARG(smem <= 227 * 1024 && n <= kMaxGram, "fh_polar_batched: Gram side %d does not fit shared memory (max 160)", n);
	FH_CHECK_ARG(workspace && workspace_bytes >= fh_polar_workspace_bytes(batch, rows, cols),
	             "fh_polar_batched: workspace too small");
	PolarWs ws = carve(batch, n, workspace);
	cudaStream_t st = (cudaStream_t)stream;
	const long long nn = (long long)n * n;
	int rc;
	// Gram (fp64 accumulate)
	if (tall) rc = gemm(FH_GEMM_F32_ACC64, n, n, rows, batch, T, 1, ld, batch_stride, T, ld, 1, batch_stride, ws.G, n, nn, stream);
	else rc = gemm(FH_GEMM_F32_ACC64, n, n, cols, batch, T, ld, 1, batch_stride, T, 1, ld, batch_stride, ws.G, n, nn, stream);
	if (rc) return rc;
	rc = launch_jacobi(batch, n, smem, st, ws.G, nullptr, nullptr, nullptr, n, max_sweeps, ws.WT, sigma, sigma_sum, ws.nsweep);
	if (rc) return rc;
	if (host_max_sweeps) {  // diagnostics only: synchronises
		int* h = (int*)malloc(sizeof(int) * batch);
		if (h) {
			FH_CUDA(cudaMemcpyAsync(h, ws.nsweep, sizeof(int) * batch, cudaMemcpyDeviceToHost, st));
			FH_CUDA(cudaStreamSynchronize(st));
			int mx = 0;
			for (int i = 0; i < batch; ++i) mx = h[i] > mx ? h[i] : mx;
			*host_max_sweeps = mx;
			free(h);
		}
	}
	// M = WT^T WT (into G):  M[a][b] = sum_j WT[j][a] WT[j][b]
	rc = gemm(FH_GEMM_F64, n, n, n, batch, ws.WT, 1, n, nn, ws.WT, n, 1, nn, ws.G, n, nn, stream);
	if (rc) return rc;
	if (tall) rc = gemm(FH_GEMM_F32xF64_F32, rows, n, n, batch, T, ld, 1, batch_stride, ws.G, n, 1, nn, U, ld, batch_stride, stream);
	else rc = gemm(FH_GEMM_F64xF32_F32, n, cols, n, batch, ws.G, n, 1, nn, T, ld, 1, batch_stride, U, ld, batch_stride, stream);
	return rc;
}

// Many Gram matrices of different sizes in one go (all bins of all chromosomes of a sweep): the
// device tables are sorted by decreasing n (longest problems first); one launch per size class so
// that small problems share an SM (8 / 4 / 2 / 1 CTAs per SM).
extern "C" int fh_polar_isqrt_multi(const double* G_all, double* WT_all, const int* dev_prob_n,
                                    const long long* dev_prob_off, const int* dev_prob_slot, const int* host_prob_n,
                                    int count, double* sigma_sum, int max_sweeps, int* dev_nsweep, void* stream) {
	FH_CHECK_ARG(count >= 0 && G_all && WT_all && dev_prob_n && dev_prob_off && dev_prob_slot && host_prob_n,
	             "fh_polar_isqrt_multi: null argument");
	if (max_sweeps <= 0 || max_sweeps > kMaxSweeps) max_sweeps = kMaxSweeps;
	cudaStream_t st = (cudaStream_t)stream;
	// class lower bounds (exclusive): one class per row-length template (32 columns per lane element). The classes are
	// independent, so every launch after the first goes to its own side stream: the last, partly filled wave of a class
	// (938 problems of the largest class on 148 SMs = 6.3 waves) shares the GPU with the next class instead of idling it.
	const int bounds[5] = {128, 96, 64, 32, 0};
	static cudaStream_t side[4] = {nullptr, nullptr, nullptr, nullptr};
	static cudaEvent_t fork_ev = nullptr, join_ev[4];
	if (!fork_ev) {
		FH_CUDA(cudaEventCreateWithFlags(&fork_ev, cudaEventDisableTiming));
		for (int c = 0; c < 4; ++c) {
			FH_CUDA(cudaStreamCreateWithFlags(&side[c], cudaStreamNonBlocking));
			FH_CUDA(cudaEventCreateWithFlags(&join_ev[c], cudaEventDisableTiming));
		}
	}
	const int tmr = fh_time_begin(FH_TIME_POLAR_JACOBI, st);  // one record for the whole group of overlapping launches
	FH_CUDA(cudaEventRecord(fork_ev, st));
	int i = 0, launch = 0;
	while (i < count) {
		const int nmax = host_prob_n[i];
		FH_CHECK_ARG(nmax > 0 && jacobi_smem(nmax) <= 227 * 1024 && nmax <= kMaxGram,
		             "fh_polar_isqrt_multi: Gram side %d does not fit shared memory (max 160)", nmax);
		int lb = 0;
		for (int c = 0; c < 5; ++c)
			if (nmax > bounds[c]) { lb = bounds[c]; break; }
		int j = i;
		while (j < count && host_prob_n[j] > lb) {
			FH_CHECK_ARG(host_prob_n[j] <= nmax, "fh_polar_isqrt_multi: table not sorted by decreasing n");
			++j;
		}
		const size_t smem = jacobi_smem(nmax);
		cudaStream_t ls = st;
		if (launch > 0 && launch <= 4) {
			ls = side[launch - 1];
			FH_CUDA(cudaStreamWaitEvent(ls, fork_ev, 0));
		}
		int rc = launch_jacobi(j - i, nmax, smem, ls, G_all, dev_prob_n + i, dev_prob_off + i, dev_prob_slot + i, 0, max_sweeps,
		                       WT_all, nullptr, sigma_sum, dev_nsweep);
		if (rc) return rc;
		if (ls != st) {
			FH_CUDA(cudaEventRecord(join_ev[launch - 1], ls));
			FH_CUDA(cudaStreamWaitEvent(st, join_ev[launch - 1], 0));
		}
		++launch;
		i = j;
	}
	fh_time_end(tmr, st);
	return FH_OK;
}

// ---------------------------------------------------------------------------------------------
// G^{-1/2} of one SPD matrix by the coupled Newton-Schulz iteration (cells x R polar):
//   Y_0 = G / tr G, Z_0 = I;  T = Z Y;  Y <- Y (3 I - T) / 2,  Z <- (3 I - T) Z / 2;  Z -> (G / tr G)^{-1/2}.
// ONE cooperative kernel runs the whole iteration (grid-wide barriers between the product phases, 32 x 32 fp64 tiles
// spread over the SMs, the residual ||I - Z Y||_F^2 summed from per-CTA partials in a fixed order so that every CTA takes
// the same stopping decision). Round 1 drove it from the host: three 256^3 GEMM launches (16 CTAs each), a residual
// kernel and two copies per step, a stream synchronisation every fourth step, and a stopping rule that never fired on
// the sweep's Gram matrices (all 200 steps ran): 6.4 ms per sweep for 2.5 GFLOP.
// ---------------------------------------------------------------------------------------------
#include <cooperative_groups.h>
namespace {
namespace cg = cooperative_groups;

constexpr int NS_T = 32, NS_K = 16;  // output tile, k-block

// c[i][j] = sum_k A[m0 + 2 ty + i][k] B[k][n0 + 2 tx + j] for row-major n x n matrices; 256 threads as 16 x 16
__device__ __forceinline__ void ns_tile_mm(const double* __restrict__ A, const double* __restrict__ B, const int n, const int m0,
                                           const int n0, double (&c)[2][2], double (*As)[NS_T + 1], double (*Bs)[NS_T + 1]) {
	const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
	c[0][0] = c[0][1] = c[1][0] = c[1][1] = 0.0;
	for (int k0 = 0; k0 < n; k0 += NS_K) {
#pragma unroll
		for (int it = 0; it < 2; ++it) {
			const int idx = tid + 256 * it;
			{
				const int mm = idx >> 4, kk = idx & 15;
				const int m = m0 + mm, k = k0 + kk;
				As[kk][mm] = (m < n && k < n) ? A[(size_t)m * n + k] : 0.0;
			}
			{
				const int kk = idx >> 5, nn = idx & 31;
				const int k = k0 + kk, nc = n0 + nn;
				Bs[kk][nn] = (k < n && nc < n) ? B[(size_t)k * n + nc] : 0.0;
			}
		}
		__syncthreads();
#pragma unroll
		for (int kk = 0; kk < NS_K; ++kk) {
			const double a0 = As[kk][2 * ty], a1 = As[kk][2 * ty + 1], b0 = Bs[kk][2 * tx], b1 = Bs[kk][2 * tx + 1];
			c[0][0] = fma(a0, b0, c[0][0]); c[0][1] = fma(a0, b1, c[0][1]);
			c[1][0] = fma(a1, b0, c[1][0]); c[1][1] = fma(a1, b1, c[1][1]);
		}
		__syncthreads();
	}
}

// scal: [0] trace, [1] last residual, [2] steps taken, [3] 1 = converged / 0 = not / -1 = NaN
__global__ void __launch_bounds__(256)
ns_fused_kernel(const double* __restrict__ G, const int n, double* Ya, double* Za, double* Yb, double* Zb, double* T,
                double* part, double* scal, double* __restrict__ out, const int max_it) {
	cg::grid_group grid = cg::this_grid();
	__shared__ double As[NS_K][NS_T + 1], Bs[NS_K][NS_T + 1];
	__shared__ double red[32];
	const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
	const int nt = (n + NS_T - 1) / NS_T, ntile = nt * nt;
	const size_t nn = (size_t)n * n;
	double tr = 0.0;  // every CTA sums the trace itself (same order: same value everywhere)
	for (int i = tid; i < n; i += 256) tr += G[(size_t)i * n + i];
	tr = fh_block_sum(tr, red);
	for (size_t i = (size_t)blockIdx.x * 256 + tid; i < nn; i += (size_t)gridDim.x * 256) {
		Ya[i] = G[i] / tr;
		Za[i] = (i / n == i % n) ? 1.0 : 0.0;
	}
	grid.sync();
	// converged below `tol`, or once the residual has stopped shrinking at the fp64 rounding level of the product
	const double tol = 1e-26 * (double)n * (double)n, stall = 1e-18 * (double)n * (double)n;
	double *Y = Ya, *Z = Za, *Yn = Yb, *Zn = Zb;
	double hres = 1.0, prev = 1e300;
	int it = 0, status = 0;
	for (; it < max_it; ++it) {
		// T = Z Y and this CTA's share of ||I - Z Y||_F^2
		double r = 0.0;
		for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
			const int m0 = (tile / nt) * NS_T, n0 = (tile % nt) * NS_T;
			double c[2][2];
			ns_tile_mm(Z, Y, n, m0, n0, c, As, Bs);
#pragma unroll
			for (int i = 0; i < 2; ++i)
#pragma unroll
				for (int j = 0; j < 2; ++j) {
					const int m = m0 + 2 * ty + i, nc = n0 + 2 * tx + j;
					if (m < n && nc < n) {
						T[(size_t)m * n + nc] = c[i][j];
						const double d = (m == nc ? 1.0 : 0.0) - c[i][j];
						r = fma(d, d, r);
					}
				}
		}
		r = fh_block_sum(r, red);
		if (tid == 0) part[blockIdx.x] = r;
		grid.sync();
		hres = 0.0;
		for (int i = 0; i < (int)gridDim.x; ++i) hres += part[i];  // fixed order: identical in every thread of the grid
		if (!(hres == hres)) { status = -1; break; }
		if (hres < tol || (hres < stall && hres > 0.25 * prev)) { status = 1; break; }
		prev = hres;
		// Yn = (3 Y - Y T) / 2,  Zn = (3 Z - T Z) / 2
		for (int w = blockIdx.x; w < 2 * ntile; w += gridDim.x) {
			const bool zpart = w >= ntile;
			const int tile = zpart ? w - ntile : w;
			const int m0 = (tile / nt) * NS_T, n0 = (tile % nt) * NS_T;
			double c[2][2];
			if (zpart) ns_tile_mm(T, Z, n, m0, n0, c, As, Bs); else ns_tile_mm(Y, T, n, m0, n0, c, As, Bs);
			const double* src = zpart ? Z : Y;
			double* dst = zpart ? Zn : Yn;
#pragma unroll
			for (int i = 0; i < 2; ++i)
#pragma unroll
				for (int j = 0; j < 2; ++j) {
					const int m = m0 + 2 * ty + i, nc = n0 + 2 * tx + j;
					if (m < n && nc < n) dst[(size_t)m * n + nc] = 1.5 * src[(size_t)m * n + nc] - 0.5 * c[i][j];
				}
		}
		grid.sync();
		double* t_ = Y; Y = Yn; Yn = t_;
		t_ = Z; Z = Zn; Zn = t_;
	}
	const double f = rsqrt(tr);
	for (size_t i = (size_t)blockIdx.x * 256 + tid; i < nn; i += (size_t)gridDim.x * 256) out[i] = Z[i] * f;
	if (blockIdx.x == 0 && tid == 0) { scal[0] = tr; scal[1] = hres; scal[2] = (double)it; scal[3] = (double)status; }
}
}  // namespace

extern "C" int fh_inv_sqrt_spd(const double* G, double* out, int n, void* ws, size_t ws_bytes, int* host_iters,
                               void* stream) {
	FH_CHECK_ARG(n > 0, "fh_inv_sqrt_spd: n <= 0");
	const size_t nn = (size_t)n * n;
	FH_CHECK_ARG(ws && ws_bytes >= (5 * nn + 512) * 8, "fh_inv_sqrt_spd: workspace too small (need (5 n^2 + 512) doubles)");
	cudaStream_t st = (cudaStream_t)stream;
	double* Ya = (double*)ws;
	double *Za = Ya + nn, *Yb = Za + nn, *Zb = Yb + nn, *T = Zb + nn, *part = T + nn, *scal = part + 504;
	static int num_sms = 0, per_sm = 0;
	if (!num_sms) {
		int dev = 0;
		FH_CUDA(cudaGetDevice(&dev));
		FH_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
		FH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ns_fused_kernel, 256, 0));
		FH_CHECK_ARG(per_sm >= 1, "fh_inv_sqrt_spd: the cooperative kernel does not fit an SM");
	}
	const int nt = (n + NS_T - 1) / NS_T;
	int grid = 2 * nt * nt;                    // the two-product phase has 2 nt^2 tiles
	if (grid > num_sms) grid = num_sms;        // one CTA per SM at most: all CTAs must be co-resident (grid barrier)
	if (grid > 500) grid = 500;
	int max_it = 200;
	int n_ = n;
	void* args[] = {(void*)&G, (void*)&n_, (void*)&Ya, (void*)&Za, (void*)&Yb, (void*)&Zb, (void*)&T, (void*)&part, (void*)&scal, (void*)&out, (void*)&max_it};
	FH_CUDA(cudaLaunchCooperativeKernel((void*)ns_fused_kernel, dim3(grid), dim3(256), args, 0, st));
	fh_count_launch(1);
	double h[4] = {0, 0, 0, 0};  // one read-back per call: the error status must reach the caller
	FH_CUDA(cudaMemcpyAsync(h, scal, sizeof(h), cudaMemcpyDeviceToHost, st));
	FH_CUDA(cudaStreamSynchronize(st));
	if (host_iters) *host_iters = (int)h[2];
	static int ns_dbg = -1;
	if (ns_dbg < 0) { const char* e = getenv("FH_NS_DEBUG"); ns_dbg = (e && e[0] == '1') ? 1 : 0; }
	if (ns_dbg) fprintf(stderr, "fh_inv_sqrt_spd: n %d iterations %d residual^2 %.3e status %d grid %d\n", n, (int)h[2], h[1], (int)h[3], grid);
	if (h[3] < 0.0) { fh_set_error("fh_inv_sqrt_spd: NaN (matrix not SPD?)"); return FH_ERR_ARG; }
	// not converged and the residual still large: singular / numerically rank-deficient Gram (the null directions of Z grow
	// ~1.5x per step) - never hand that back
	if (h[3] < 1.0 && !(h[1] < 1e-12 * (double)n * (double)n)) {
		fh_set_error("fh_inv_sqrt_spd: Newton-Schulz did not converge in 200 iterations, ||I - ZY||_F^2 = %.3g (Gram matrix singular: fewer rows than columns, or a rank-deficient input)", h[1]);
		return FH_ERR_ARG;
	}
	return FH_OK;
}
