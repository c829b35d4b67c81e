/* fh_b200 - C ABI of the B200-native Fast-Higashi hot path (libfh_b200.so).
 *
 * The reference (ma-compbio/Fast-Higashi) is pure Python and has no FFI layer; its seam for this
 * path is Python-level. Each entry point below therefore cites the
 * reference *call site* it replaces (file:line under fasthigashi/). INTEGRATION.md shows the
 * ctypes binding a reference maintainer would add.
 *
 * Conventions: plain C types only; every pointer is a DEVICE pointer unless named host_*;
 * `stream` is a cudaStream_t passed as void*; the caller (PyTorch) owns all memory, outputs and
 * workspaces included; return value 0 = ok, otherwise fh_last_error() describes the failure.
 * One host thread per device; calls are asynchronous on `stream` unless stated otherwise.
 */
#ifndef FH_B200_H
#define FH_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* fh_last_error(void);
int fh_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches claim) */
long long fh_launch_count(void);
/* FH_GEMM_TF32X3 calls that ran on the CUDA-core fp32 kernel because TMA could not describe them */
long long fh_tc_fallback_count(void);
/* Optional per-kernel timing for the bench's roofline: while enabled, every launch of the kernels below is bracketed by
 * CUDA events on its own stream; fh_timing_read synchronises on them and returns the summed duration and the launch count.
 * fh_timing_enable(0 or 1) also discards what was recorded before. */
enum { FH_TIME_DENSIFY = 0, FH_TIME_RWR_CHAIN = 1, FH_TIME_GEMM_TC = 2, FH_TIME_POLAR_JACOBI = 3 };
void fh_timing_enable(int on);
int fh_timing_read(int id, double* host_total_ms, long long* host_launches);

/* ---------------------------------------------------------------------------------------------
 * Generic batched strided GEMM.  C[b](m,n) = alpha * sum_k A[b](m,k)*kscale[b][k]*B[b](k,n)
 *                                            (+diag on m==n) (*|/ cscale[b][n]) + beta*C[b](m,n)
 * Replaces the torch.bmm / torch.matmul / opt_einsum.contract call sites of
 * partial_rwr.py:85,116,138 and parafac2_intergrative.py:374-383,422-430,522-524,
 * parafac_integrative.py:43-62.
 * ------------------------------------------------------------------------------------------- */
enum { FH_GEMM_F32 = 0,         /* fp32 in/out, fp32 accumulate (CUDA cores)            */
       FH_GEMM_F32_ACC64 = 1,   /* fp32 in, fp64 accumulate, fp64 out                     */
       FH_GEMM_F64 = 2,         /* fp64 in/out                                            */
       FH_GEMM_F32xF64_F32 = 3, /* A fp32, B fp64, fp64 accumulate, fp32 out              */
       FH_GEMM_TF32X3 = 4,      /* fp32 in/out on tcgen05 tensor cores, 3xTF32 split      */
       FH_GEMM_F64xF32_F32 = 5 };/* A fp64, B fp32, fp64 accumulate, fp32 out              */
enum { FH_EPI_NONE = 0, FH_EPI_DIAG_ADD = 1,
       FH_EPI_SYMMETRIC = 2 };  /* the caller guarantees C = C^T (B = A^T, M = N, beta = 0): only the tiles on and above the
                                  diagonal are computed, the others are their mirror images (CUDA-core kernels; elsewhere = NONE) */

typedef struct fh_gemm_desc {
	int M, N, K, batch;
	long long sa_m, sa_k;          /* A(m,k) = A[m*sa_m + k*sa_k]; one of them must be 1 */
	long long sb_k, sb_n;          /* B(k,n) = B[k*sb_k + n*sb_n]; one of them must be 1 */
	long long ldc;                 /* C(m,n) = C[m*ldc + n]                              */
	long long batch_a, batch_b, batch_c; /* element strides between batch items          */
	double alpha, beta;
	int dtype;                     /* FH_GEMM_*                                          */
	int epilogue;                  /* FH_EPI_*                                           */
	double diag;                   /* FH_EPI_DIAG_ADD: value added where m == n          */
	const float* kscale;           /* optional [batch][K] scale on the reduction index   */
	long long kscale_batch;
	const float* cscale;           /* optional [batch][N] scale on output columns        */
	long long cscale_batch;
	int cscale_recip;              /* 1: divide by cscale instead of multiplying         */
} fh_gemm_desc;

int fh_gemm_batched(const fh_gemm_desc* d, const void* A, const void* B, void* C, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Partial RWR imputation of one bin-block for a range of cells, from the block-CSR.
 * Replaces Chrom_Dataset.fetch -> densify_jit (sparse_for_schic.py:279-320,588-613) followed by
 * partial_rwr (partial_rwr.py:45-175) as called at parafac2_intergrative.py:131,216,357,451,508,
 * 789,813.
 * ------------------------------------------------------------------------------------------- */
typedef struct fh_rwr_desc {
	int nb;        /* rows of the block                                                       */
	int w;         /* window width (columns)                                                  */
	int ldw;       /* leading dimension (floats) of every nb x w panel, >= w, multiple of 4   */
	int s;         /* first window column of the diagonal nb x nb block (local_bin_slice)     */
	int k;         /* forced number of RWR steps; < 0 = auto-stop (partial_rwr.py:99-126)     */
	int do_conv, do_rwr, do_col;
	int cell0;     /* first cell: CSR row of (cell c, row r) is rowptr[(cell0+c)*nb + r]       */
	int ncell;     /* cells imputed by this call                                              */
	int use_tensor_cores; /* 1: tcgen05 (fused 3xFP16 / 3xTF32 kernels, 3xTF32 GEMMs), 0: CUDA-core fp32 GEMMs */
	long long nnz; /* total entries of col/val (bounds the 128-bit vector loads)               */
} fh_rwr_desc;

size_t fh_rwr_workspace_bytes(const fh_rwr_desc* d);
/* out: (ncell, nb, ldw) fp32, cell stride out_cell_stride floats; pad columns [w, ldw) are 0.
 * Forced step counts on the tensor cores run as ONE fused kernel per call when nb <= 128 (3xFP16 operand pairs, do_col included
 * from two steps on; FH_RWR_F16=0 / FH_RWR_FUSED select the 3xTF32 kernel or the per-step kernels): INTEGRATION.md.
 * bin_cov: [ncell][>=w] per-cell coverage of the window columns (row stride bin_cov_ld), only
 * read when do_col. host_n_iter (may be NULL): receives the RWR step count the reference would
 * return (auto mode: applied steps - 1; forced mode: k). Auto mode synchronises the stream. */
int fh_rwr_batched(const fh_rwr_desc* d, const int32_t* rowptr, const int16_t* col, const float* val,
                   const float* bin_cov, long long bin_cov_ld, float* out, long long out_cell_stride,
                   void* workspace, size_t workspace_bytes, int* host_n_iter, void* stream);

/* densify only (sparse_for_schic.py:279-320): out (ncell, nb, ldw), floor 1e-8, pad columns 0 */
int fh_densify(const fh_rwr_desc* d, const int32_t* rowptr, const int16_t* col, const float* val,
               float* out, long long out_cell_stride, void* stream);
/* partial_rwr on an already dense (ncell, nb, ldw) block, in place (API-compat path) */
int fh_rwr_dense(const fh_rwr_desc* d, float* x, long long cell_stride, const float* bin_cov,
                 long long bin_cov_ld, void* workspace, size_t workspace_bytes, int* host_n_iter,
                 void* stream);

/* init_params helpers (parafac2_intergrative.py:141-142,147,229):
 * cov[c][col] += sum_rows x[c][row][col];  pooled = avg_pool2d(x, ll, ll) (floor mode) */
int fh_colsum_accum(const float* x, int ncell, int nb, int w, int ldw, long long cell_stride,
                    float* cov, long long cov_ld, void* stream);
int fh_avgpool(const float* x, int ncell, int nb, int w, int ldw, long long cell_stride, int ll,
               float* out, long long out_cell_stride, void* stream);

/* sum of squares of a strided 2-D view, accumulated (+=) into a device double
 * (parafac2_intergrative.py:368-369, C2) */
int fh_sqnorm_accum(const float* x, long long rows, long long cols, long long ld, double* acc, void* stream);
/* acc += <x, y> over a strided 2-D view (parafac2_intergrative.py:486) */
int fh_dot_accum(const float* x, const float* y, long long rows, long long cols, long long ldx,
                 long long ldy, double* acc, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Batched polar factor U = T (T^T T)^{-1/2}  (= U_svd Vh_svd) and sum of singular values.
 * Replaces project2orthogonal (project2orthogonal.py:6-55) as called at
 * parafac2_intergrative.py:396. fp64 Gram + pivoted Cholesky + one-sided Jacobi (DESIGN.md "polar").
 * ------------------------------------------------------------------------------------------- */
size_t fh_polar_workspace_bytes(int batch, int rows, int cols);
/* T, U: (batch, rows, ld) fp32; sigma_sum: NULL or [batch] doubles (sum of singular values);
 * sigma: NULL or [batch][n] doubles, the singular values (unsorted), n = min(rows, cols);
 * max_sweeps <= 0: default cap; host_max_sweeps: NULL, or receives the largest Jacobi sweep count
 * of the batch (diagnostics; synchronises the stream). */
int fh_polar_batched(const float* T, float* U, int batch, int rows, int cols, long long ld,
                     long long batch_stride, double* sigma_sum, double* sigma, int max_sweeps,
                     void* workspace, size_t workspace_bytes, int* host_max_sweeps, void* stream);

/* The eigen step of the polar factor for MANY Gram matrices of different sizes at once (all bins of
 * all chromosomes of one sweep). G_all / WT_all: concatenated n_i x n_i fp64 matrices; problem i
 * (device tables, sorted by DEcreasing n; host_prob_n is the same n table on the host) reads G at
 * dev_prob_off[i] doubles and writes WT there: row j = w_j lambda_j^{-3/4}, so that
 * G^{-1/2} = WT^T WT; sigma_sum[dev_prob_slot[i]] = sum of sqrt(eigenvalues). dev_nsweep: NULL or
 * [count] ints (by slot). */
int fh_polar_isqrt_multi(const double* G_all, double* WT_all, const int* dev_prob_n,
                         const long long* dev_prob_off, const int* dev_prob_slot, const int* host_prob_n,
                         int count, double* sigma_sum, int max_sweeps, int* dev_nsweep, void* stream);

/* Inverse square root of ONE symmetric positive definite n x n fp64 matrix by the coupled
 * Newton-Schulz iteration (tall cells x R polar, parafac2_intergrative.py:483,831: V = M G^{-1/2}
 * with G = M^T M all-reduced across ranks). One cooperative kernel runs the whole iteration; the stream is synchronised
 * once at the end to return the iteration count and the error status (singular input = error).
 * ws: >= 5*n*n + 512 doubles. */
int fh_inv_sqrt_spd(const double* G, double* out, int n, void* ws, size_t ws_bytes, int* host_iters,
                    void* stream);

/* ---------------------------------------------------------------------------------------------
 * Inner CP-ALS on the projected tensor of one chromosome (parafac_integrative.py:28-112, called
 * at parafac2_intergrative.py:680-686) and the core norm (:623-632,697-706).
 * Y (n, r, R) fp32 contiguous; factors A (n x r), B (r x r), D (R x r) fp32 row-major, updated in
 * place. Per mode ONE small-kernel epilogue does Hadamard-of-Grams + 1e-10 ridge + SPD inverse
 * (fp64) and the factor is the MTTKRP times that inverse; balance_norm is fused in the same call.
 * host_out[0] = ||Xhat||^2, host_out[1] = <Xhat, Y> of the last iteration (both only computed when
 * n_iter_max > 1, where the reference's early-stop test needs them; otherwise 0).
 * The early stop is decided on the device (no read-back per iteration); the stream is synchronised only when
 * host_out is given AND n_iter_max > 1 (one read of the two scalars at the end).
 * ------------------------------------------------------------------------------------------- */
size_t fh_cp_als_workspace_bytes(int n, int r, int R);
int fh_cp_als(const float* Y, int n, int r, int R, float* A, float* B, float* D, int n_iter_max,
              void* workspace, size_t workspace_bytes, double* host_out, void* stream);
/* acc += ||[[A,B,D]]||^2 = sum((A^T A)*(B^T B)*(D^T D)); ws >= 3*r*r doubles */
int fh_cp_core_sqnorm(const float* A, int n, const float* B, const float* D, int R, int r,
                      double* ws, double* acc, void* stream);

/* out[i][j][p] = F[j][p] * Arows[i][p] (p < r), zero in the pad columns r..ldo-1: the per-bin scaled
 * factor copies (B diag(A_i), D diag(A_i)) of the 'ir,jr,kr->...' contractions at
 * parafac2_intergrative.py:374,422. F: rows x r (row pitch ldf); Arows: nb x r; out: (nb, rows, ldo). */
int fh_scale_cols_batched(const float* F, int rows, int r, long long ldf, const float* Arows, int nb, int ldo,
                          float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
