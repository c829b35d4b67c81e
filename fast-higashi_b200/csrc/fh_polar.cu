// Batched polar factor (reference: project2orthogonal.py:6-55, called at
// parafac2_intergrative.py:396 with one (w x r) matrix per bin).
//
// U = T (T^T T)^{-1/2}. The per-bin matrices have condition numbers 1e4..2e7, so the
// Gram route only reaches parity in fp64 (measured: fp32 Gram -> 15 % loss error, fp64 Gram -> 3e-6;
// DESIGN.md "polar"). Pipeline, all batched over the bins of one block:
//   G = T^T T                     fp64 accumulate           (fh_gemm_batched, FH_GEMM_F32_ACC64)
//   [warm start: G <- V0^T G V0]  fp64                      (FH_GEMM_F64)
//   G -> diag(lambda), rotation log   two-sided cyclic Jacobi, G resident in shared memory,
//                                     parallel (round-robin) ordering, fused 2x2-block updates
//   V = V0 * rotations            row slabs of V in shared memory, embarrassingly parallel
//   M = V lambda^{-1/2} V^T       fp64
//   U = T M                       fp64 accumulate, fp32 out
#include "fh_common.cuh"
#include "../../include/fh_b200.h"
#include <math.h>

namespace {

__host__ __device__ inline int even_up(int n) { return (n + 1) & ~1; }

// round-robin tournament on m (even) players: pair t of step s
__device__ __forceinline__ void rr_pair(int m, int s, int t, int& p, int& q) {
	const int mm = m - 1;
	if (t == 0) {
		p = mm;
		q = s % mm;
	} else {
		p = (s + t) % mm;
		q = (s - t + mm) % mm;
	}
}

// One CTA per matrix. G (n x n, fp64, symmetric) is copied to shared memory (m x ldg, zero
// padded to even m), swept until no rotation exceeds the threshold, and written back
// diagonalised. Every rotation (c, s) is logged: rot[((sweep*(m-1) + step)*(m/2) + t)].
__global__ void __launch_bounds__(512)
jacobi_kernel(double* __restrict__ Gall, int n, int max_sweeps, double2* __restrict__ rot_all,
              int* __restrict__ nsweep_out, double* __restrict__ lam_all) {
	extern __shared__ double sm[];
	const int m = even_up(n), half = m >> 1;
	const int ldg = m | 1;
	double* G = sm;                       // m x ldg
	double* cs_c = G + (size_t)m * ldg;   // half
	double* cs_s = cs_c + half;           // half
	int* pq = (int*)(cs_s + half);        // 2*half
	__shared__ int s_rotated;
	const int b = blockIdx.x;
	double* Gg = Gall + (size_t)b * n * n;
	double2* rot = rot_all + (size_t)b * max_sweeps * (m - 1) * half;
	const int tid = threadIdx.x, nt = blockDim.x;
	for (int i = tid; i < m * ldg; i += nt) {
		int r = i / ldg, c = i - r * ldg;
		G[i] = (r < n && c < n) ? Gg[(size_t)r * n + c] : 0.0;
	}
	__syncthreads();
	int sweep = 0;
	for (; sweep < max_sweeps; ++sweep) {
		if (tid == 0) s_rotated = 0;
		__syncthreads();
		for (int step = 0; step < m - 1; ++step) {
			if (tid < half) {
				int p, q;
				rr_pair(m, step, tid, p, q);
				double c = 1.0, s = 0.0;
				if (p < n && q < n) {
					double a = G[p * ldg + p], bb = G[q * ldg + q], g = G[p * ldg + q];
					if (fabs(g) > 1e-15 * sqrt(fabs(a * bb)) && g != 0.0) {
						double theta = (bb - a) / (2.0 * g);
						double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
						c = 1.0 / sqrt(t * t + 1.0);
						s = t * c;
						s_rotated = 1;
					}
				}
				cs_c[tid] = c; cs_s[tid] = s;
				pq[2 * tid] = p; pq[2 * tid + 1] = q;
				rot[((size_t)sweep * (m - 1) + step) * half + tid] = make_double2(c, s);
			}
			__syncthreads();
			// G <- J^T G J, one thread per 2x2 block (row pair a, column pair bcol)
			for (int i = tid; i < half * half; i += nt) {
				int a = i / half, bc = i - a * half;
				int pa = pq[2 * a], qa = pq[2 * a + 1], pb = pq[2 * bc], qb = pq[2 * bc + 1];
				double ca = cs_c[a], sa = cs_s[a], cb = cs_c[bc], sb = cs_s[bc];
				double g00 = G[pa * ldg + pb], g01 = G[pa * ldg + qb];
				double g10 = G[qa * ldg + pb], g11 = G[qa * ldg + qb];
				// columns: col_p' = c col_p - s col_q ; col_q' = s col_p + c col_q
				double h00 = cb * g00 - sb * g01, h01 = sb * g00 + cb * g01;
				double h10 = cb * g10 - sb * g11, h11 = sb * g10 + cb * g11;
				// rows, same form
				G[pa * ldg + pb] = ca * h00 - sa * h10;
				G[pa * ldg + qb] = ca * h01 - sa * h11;
				G[qa * ldg + pb] = sa * h00 + ca * h10;
				G[qa * ldg + qb] = sa * h01 + ca * h11;
			}
			__syncthreads();
		}
		if (!s_rotated) break;  // block-uniform: read after the barrier above
		__syncthreads();
	}
	const int done = sweep;  // sweeps [0, done) contain rotations; a rotation-free sweep ends the loop
	if (tid == 0) nsweep_out[b] = done;
	for (int i = tid; i < n * n; i += nt) {
		int r = i / n, c = i - r * n;
		Gg[i] = G[r * ldg + c];
	}
	for (int i = tid; i < n; i += nt) lam_all[(size_t)b * n + i] = G[i * ldg + i];
}

// V <- V0 * (logged rotations). grid (slabs, batch); RS rows of V per CTA in shared memory.
constexpr int RS = 32;
__global__ void __launch_bounds__(256)
vapply_kernel(const double* __restrict__ V0, double* __restrict__ Vout, int n, int max_sweeps,
              const double2* __restrict__ rot_all, const int* __restrict__ nsweep) {
	extern __shared__ double sm[];
	const int m = even_up(n), half = m >> 1;
	const int ldv = m | 1;
	double* V = sm;  // RS x ldv
	double2* cs = (double2*)(V + (size_t)RS * ldv);
	const int b = blockIdx.y, r0 = blockIdx.x * RS;
	const int rows = min(RS, n - r0);
	const int tid = threadIdx.x, nt = blockDim.x;
	for (int i = tid; i < RS * ldv; i += nt) {
		int r = i / ldv, c = i - r * ldv;
		double v = 0.0;
		if (r < rows && c < n) v = V0 ? V0[((size_t)b * n + r0 + r) * n + c] : ((r0 + r) == c ? 1.0 : 0.0);
		V[i] = v;
	}
	const double2* rot = rot_all + (size_t)b * max_sweeps * (m - 1) * half;
	const int ns = nsweep[b];
	__syncthreads();
	for (int sw = 0; sw < ns; ++sw) {
		for (int step = 0; step < m - 1; ++step) {
			for (int i = tid; i < half; i += nt) cs[i] = rot[((size_t)sw * (m - 1) + step) * half + i];
			__syncthreads();
			for (int i = tid; i < rows * half; i += nt) {
				int r = i / half, t = i - r * half;
				double2 c_s = cs[t];
				if (c_s.y == 0.0) continue;
				int p, q;
				rr_pair(m, step, t, p, q);
				double vp = V[r * ldv + p], vq = V[r * ldv + q];
				V[r * ldv + p] = c_s.x * vp - c_s.y * vq;
				V[r * ldv + q] = c_s.y * vp + c_s.x * vq;
			}
			__syncthreads();
		}
	}
	for (int i = tid; i < rows * n; i += nt) {
		int r = i / n, c = i - r * n;
		Vout[((size_t)b * n + r0 + r) * n + c] = V[r * ldv + c];
	}
}

// W = V * diag(lambda_clamped^{-1/4});  sigma_sum[b] = sum sqrt(max(lambda, 0))
__global__ void __launch_bounds__(256)
scale_cols_kernel(const double* __restrict__ V, const double* __restrict__ lam, int n,
                  double* __restrict__ W, double* __restrict__ sigma_sum, double* __restrict__ sigma) {
	__shared__ double red[32];
	__shared__ double s_max;
	const int b = blockIdx.x;
	const double* l = lam + (size_t)b * n;
	double mx = 0.0, ss = 0.0;
	for (int i = threadIdx.x; i < n; i += blockDim.x) {
		mx = fmax(mx, l[i]);
		double sv = sqrt(fmax(l[i], 0.0));
		ss += sv;
		if (sigma) sigma[(size_t)b * n + i] = sv;
	}
	ss = fh_block_sum(ss, red);
	// block max via the same scratch
	for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
	__syncthreads();
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
	__syncthreads();
	if (threadIdx.x == 0) {
		double v = 0.0;
		for (int i = 0; i < (blockDim.x + 31) / 32; ++i) v = fmax(v, red[i]);
		s_max = v;
		if (sigma_sum) sigma_sum[b] = ss;
	}
	__syncthreads();
	const double floor_l = fmax(s_max * 1e-17, 1e-300);
	const double* v = V + (size_t)b * n * n;
	double* w = W + (size_t)b * n * n;
	for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
		int c = i % n;
		double lc = fmax(l[c], floor_l);
		w[i] = v[i] * rsqrt(sqrt(lc));
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
	double *G, *V, *W, *lam;
	double2* rot;
	int* nsweep;
	size_t bytes;
};
PolarWs carve(int batch, int n, int max_sweeps, void* ws) {
	PolarWs p;
	const int m = even_up(n);
	char* b = (char*)ws;
	size_t nn = al((size_t)batch * n * n * 8);
	p.G = (double*)b; b += nn;
	p.V = (double*)b; b += nn;
	p.W = (double*)b; b += nn;
	p.lam = (double*)b; b += al((size_t)batch * n * 8);
	p.rot = (double2*)b; b += al((size_t)batch * max_sweeps * (m - 1) * (m / 2) * 16);
	p.nsweep = (int*)b; b += al((size_t)batch * 4);
	p.bytes = (size_t)(b - (char*)ws);
	return p;
}

constexpr int kMaxSweeps = 32;

}  // namespace

extern "C" size_t fh_polar_workspace_bytes(int batch, int rows, int cols) {
	int n = rows < cols ? rows : cols;
	if (batch <= 0 || n <= 0) return 0;
	return carve(batch, n, kMaxSweeps, nullptr).bytes;
}

extern "C" int fh_polar_batched(const float* T, float* U, int batch, int rows, int cols, long long ld,
                                long long batch_stride, double* sigma_sum, double* sigma, double* eigvec_state,
                                int warm, int max_sweeps, void* workspace, size_t workspace_bytes, void* stream) {
	FH_CHECK_ARG(batch >= 0 && rows > 0 && cols > 0 && ld >= cols, "fh_polar_batched: bad shape");
	if (batch == 0) return FH_OK;
	FH_CHECK_ARG(batch <= 65535, "fh_polar_batched: batch > 65535");
	if (max_sweeps <= 0 || max_sweeps > kMaxSweeps) max_sweeps = kMaxSweeps;
	const bool tall = rows >= cols;
	const int n = tall ? cols : rows;
	const int m = even_up(n), half = m / 2, ldg = m | 1;
	size_t smem = ((size_t)m * ldg + 2 * half) * 8 + (size_t)2 * half * 4;
	FH_CHECK_ARG(smem <= 227 * 1024, "fh_polar_batched: Gram side %d does not fit shared memory (max ~166)", n);
	FH_CHECK_ARG(workspace && workspace_bytes >= fh_polar_workspace_bytes(batch, rows, cols),
	             "fh_polar_batched: workspace too small");
	PolarWs ws = carve(batch, n, kMaxSweeps, workspace);
	cudaStream_t st = (cudaStream_t)stream;
	const long long nn = (long long)n * n;
	int rc;
	// Gram (fp64 accumulate)
	if (tall) rc = gemm(FH_GEMM_F32_ACC64, n, n, rows, batch, T, 1, ld, batch_stride, T, ld, 1, batch_stride, ws.G, n, nn, stream);
	else rc = gemm(FH_GEMM_F32_ACC64, n, n, cols, batch, T, ld, 1, batch_stride, T, 1, ld, batch_stride, ws.G, n, nn, stream);
	if (rc) return rc;
	const double* V0 = nullptr;
	if (warm && eigvec_state) {
		// G <- V0^T G V0 : nearly diagonal when the factors moved little since the last sweep
		rc = gemm(FH_GEMM_F64, n, n, n, batch, ws.G, n, 1, nn, eigvec_state, n, 1, nn, ws.W, n, nn, stream);
		if (rc) return rc;
		rc = gemm(FH_GEMM_F64, n, n, n, batch, eigvec_state, 1, n, nn, ws.W, n, 1, nn, ws.G, n, nn, stream);
		if (rc) return rc;
		V0 = eigvec_state;
	}
	int threads = half * half;
	threads = threads < 64 ? 64 : (threads > 512 ? 512 : (threads + 31) / 32 * 32);
	FH_CUDA(cudaFuncSetAttribute(jacobi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
	jacobi_kernel<<<batch, threads, smem, st>>>(ws.G, n, max_sweeps, ws.rot, ws.nsweep, ws.lam);
	FH_LAUNCH_CHECK();
	size_t smem_v = ((size_t)RS * (m | 1)) * 8 + (size_t)half * 16;
	FH_CUDA(cudaFuncSetAttribute(vapply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_v));
	dim3 gv(fh_cdiv(n, RS), batch);
	vapply_kernel<<<gv, 256, smem_v, st>>>(V0, ws.V, n, max_sweeps, ws.rot, ws.nsweep);
	FH_LAUNCH_CHECK();
	if (eigvec_state) FH_CUDA(cudaMemcpyAsync(eigvec_state, ws.V, (size_t)batch * nn * 8, cudaMemcpyDeviceToDevice, st));
	scale_cols_kernel<<<batch, 256, 0, st>>>(ws.V, ws.lam, n, ws.W, sigma_sum, sigma);
	FH_LAUNCH_CHECK();
	// M = W W^T (into G), U = T M or M T
	rc = gemm(FH_GEMM_F64, n, n, n, batch, ws.W, n, 1, nn, ws.W, 1, n, nn, ws.G, n, nn, stream);
	if (rc) return rc;
	if (tall) rc = gemm(FH_GEMM_F32xF64_F32, rows, n, n, batch, T, ld, 1, batch_stride, ws.G, n, 1, nn, U, ld, batch_stride, stream);
	else rc = gemm(FH_GEMM_F64xF32_F32, n, cols, n, batch, ws.G, n, 1, nn, T, ld, 1, batch_stride, U, ld, batch_stride, stream);
	return rc;
}

// ---------------------------------------------------------------------------------------------
// G^{-1/2} of one SPD matrix by the coupled Newton-Schulz iteration (cells x R polar).
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void ns_init_kernel(const double* __restrict__ G, int n, double* __restrict__ Y, double* __restrict__ Z,
                               double* __restrict__ scal) {
	__shared__ double red[32];
	double tr = 0.0;
	for (int i = threadIdx.x; i < n; i += blockDim.x) tr += G[(size_t)i * n + i];
	tr = fh_block_sum(tr, red);
	if (threadIdx.x == 0) scal[0] = tr;
	for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
		Y[i] = G[i] / tr;
		Z[i] = (i / n == i % n) ? 1.0 : 0.0;
	}
}
// T = 0.5 * (3 I - T);  scal[1] = ||I - ZY||_F^2 (T holds ZY on entry)
__global__ void ns_mid_kernel(double* __restrict__ T, int n, double* __restrict__ scal) {
	__shared__ double red[32];
	double r = 0.0;
	for (int i = threadIdx.x; i < n * n; i += blockDim.x) {
		double eye = (i / n == i % n) ? 1.0 : 0.0;
		double zy = T[i];
		double d = eye - zy;
		r += d * d;
		T[i] = 0.5 * (3.0 * eye - zy);
	}
	r = fh_block_sum(r, red);
	if (threadIdx.x == 0) scal[1] = r;
}
__global__ void ns_final_kernel(const double* __restrict__ Z, int n, const double* __restrict__ scal,
                                double* __restrict__ out) {
	double f = rsqrt(scal[0]);
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * n; i += gridDim.x * blockDim.x) out[i] = Z[i] * f;
}
}  // namespace

extern "C" int fh_inv_sqrt_spd(const double* G, double* out, int n, void* ws, size_t ws_bytes, int* host_iters,
                               void* stream) {
	FH_CHECK_ARG(n > 0, "fh_inv_sqrt_spd: n <= 0");
	const size_t nn = (size_t)n * n;
	FH_CHECK_ARG(ws && ws_bytes >= (4 * nn + 8) * 8, "fh_inv_sqrt_spd: workspace too small");
	cudaStream_t st = (cudaStream_t)stream;
	double* Y = (double*)ws;
	double* Z = Y + nn;
	double* T = Z + nn;
	double* tmp = T + nn;
	double* scal = tmp + nn;
	ns_init_kernel<<<1, 1024, 0, st>>>(G, n, Y, Z, scal);
	FH_LAUNCH_CHECK();
	int it = 0, rc;
	double hres = 1.0;
	const double tol = 1e-26 * (double)n * (double)n;  // ||I - ZY||_F^2
	for (; it < 200; ++it) {
		rc = gemm(FH_GEMM_F64, n, n, n, 1, Z, n, 1, 0, Y, n, 1, 0, T, n, 0, stream);  // T = Z Y
		if (rc) return rc;
		ns_mid_kernel<<<1, 1024, 0, st>>>(T, n, scal);
		FH_LAUNCH_CHECK();
		if ((it & 3) == 3 || it > 24) {
			FH_CUDA(cudaMemcpyAsync(&hres, scal + 1, 8, cudaMemcpyDeviceToHost, st));
			FH_CUDA(cudaStreamSynchronize(st));
			if (!(hres == hres)) { fh_set_error("fh_inv_sqrt_spd: NaN (matrix not SPD?)"); return FH_ERR_ARG; }
			if (hres < tol) break;
		}
		rc = gemm(FH_GEMM_F64, n, n, n, 1, Y, n, 1, 0, T, n, 1, 0, tmp, n, 0, stream);  // Y = Y T
		if (rc) return rc;
		FH_CUDA(cudaMemcpyAsync(Y, tmp, nn * 8, cudaMemcpyDeviceToDevice, st));
		rc = gemm(FH_GEMM_F64, n, n, n, 1, T, n, 1, 0, Z, n, 1, 0, tmp, n, 0, stream);  // Z = T Z
		if (rc) return rc;
		FH_CUDA(cudaMemcpyAsync(Z, tmp, nn * 8, cudaMemcpyDeviceToDevice, st));
	}
	if (host_iters) *host_iters = it;
	ns_final_kernel<<<fh_cdiv(nn, 256), 256, 0, st>>>(Z, n, scal, out);
	FH_LAUNCH_CHECK();
	return FH_OK;
}
