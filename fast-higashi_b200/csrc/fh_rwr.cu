// Partial RWR imputation of one bin-block (reference: sparse_for_schic.py:279-320 densify,
// partial_rwr.py:45-175). The block-CSR of a range of cells is read once from HBM with 128-bit
// loads, staged through shared memory, and the imputed dense nb x w panel of every cell is written
// once; everything in between (conv'd panel A, A A^T, transition matrix P, the Q iterates) lives in
// an L2-sized workspace that is reused chunk after chunk.
#include <cuda_fp16.h>
#include "fh_common.cuh"
#include "../../include/fh_b200.h"

#define FH_FLOOR 1e-8f
#define FH_EPS 1e-15f

namespace {

constexpr int RT = 32;  // output rows per CTA in densify/conv

// ---------------------------------------------------------------------------------------------
// K1+K2: CSR -> dense (floor 1e-8) [-> 3x3 mean, zero padding counted, floor 1e-8]
// grid (ceil(nb/RT), ncell), 256 threads, smem (RT+2) x (ldw+8) floats + rowptr slice.
// Tile layout: window column c of tile row tr lives at tile[tr * tp + c + 4] (tp = ldw + 8, a multiple of 4), so that a
// group of four columns starting at a multiple of 4 is one aligned 128-bit shared load; index 3 is the left zero-padding
// column, indices >= w + 4 the right one.
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline size_t densify_smem_bytes(int ldw) { return (size_t)((RT + 2) * (ldw + 8) + RT + 4) * 4; }

// OUT16: the panel leaves as two binary16 planes (hi = rn16(s x), lo = rn16(s x - hi); rows of ld16 halves, the lo plane
// lo_plane halves after the hi plane) for the 3xFP16 kernel (fh_rwr_chain16.cu); s is the power of two that brings amax[cell]
// (the largest floored CSR value of the cell's rows of the block: an upper bound of every entry of its panel) into [2^13, 2^14).
__device__ __forceinline__ void split_pair16(float a, float b, unsigned& hi, unsigned& lo) {
	asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
	float fa, fb;
	asm("{\n.reg .b16 l, h;\nmov.b32 {l, h}, %2;\ncvt.f32.f16 %0, l;\ncvt.f32.f16 %1, h;\n}\n" : "=f"(fa), "=f"(fb) : "r"(hi));
	asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - fb), "f"(a - fa));
}
__device__ __forceinline__ void store4_16(__half* hi_row, long long lo_plane, float4 o, float sa) {
	uint2 vh, vl;
	split_pair16(o.x * sa, o.y * sa, vh.x, vl.x);
	split_pair16(o.z * sa, o.w * sa, vh.y, vl.y);
	*reinterpret_cast<uint2*>(hi_row) = vh;
	*reinterpret_cast<uint2*>(hi_row + lo_plane) = vl;
}

template <bool FROM_DENSE, bool OUT16>
__global__ void __launch_bounds__(256)
densify_conv_kernel(const int32_t* __restrict__ rowptr, const int16_t* __restrict__ col,
                    const float* __restrict__ val, long long nnz_total,
                    const float* __restrict__ dense_in, long long in_cell_stride,
                    int cell0, int nb, int w, int ldw, int do_conv,
                    float* __restrict__ out, long long out_cell_stride,
                    int ld16, long long lo_plane, const unsigned* __restrict__ amax, int pad16) {
	extern __shared__ __align__(16) float smem[];
	const int tp = ldw + 8;
	float* tile = smem;                                   // (RT+2) x tp
	int* rp = (int*)(smem + (RT + 2) * tp);               // RT+3 row pointers
	const int cell = blockIdx.y;
	const int r0 = blockIdx.x * RT;
	const int halo = do_conv ? 1 : 0;
	const int ra = max(r0 - halo, 0), rb = min(r0 + RT + halo, nb);  // staged global rows [ra, rb)
	const int tid = threadIdx.x;

	// 1. background: floor inside the block, 0 in the zero padding ring and the pad columns (128-bit stores; no integer
	// division in any per-element loop of this kernel - they dominated the first version's instruction count)
	const int lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
	for (int tr = warp; tr < RT + 2; tr += nwarp) {
		const int gr = r0 - 1 + tr;
		const bool rin = gr >= 0 && gr < nb;
		float4* trow = reinterpret_cast<float4*>(tile + tr * tp);
		for (int q = lane; q < (tp >> 2); q += 32) {
			const int c = 4 * q - 4;  // window column of the first element
			float4 v;
			v.x = (rin && c >= 0 && c < w) ? FH_FLOOR : 0.f;
			v.y = (rin && c + 1 >= 0 && c + 1 < w) ? FH_FLOOR : 0.f;
			v.z = (rin && c + 2 >= 0 && c + 2 < w) ? FH_FLOOR : 0.f;
			v.w = (rin && c + 3 >= 0 && c + 3 < w) ? FH_FLOOR : 0.f;
			trow[q] = v;
		}
	}
	if (!FROM_DENSE) {
		const long long base = (long long)(cell0 + cell) * nb;
		for (int i = tid; i <= rb - ra; i += blockDim.x) rp[i] = rowptr[base + ra + i];
	}
	__syncthreads();
	if (FROM_DENSE) {
		const float* src = dense_in + (long long)cell * in_cell_stride;
		for (int gr = ra + warp; gr < rb; gr += nwarp) {
			float* trow = tile + (gr - r0 + 1) * tp + 4;
			const float* srow = src + (long long)gr * ldw;
			for (int c = lane; c < w; c += 32) trow[c] = fmaxf(srow[c], FH_FLOOR);
		}
	} else {
		// 2. scatter the CSR entries of rows [ra, rb): 8 entries per thread and step through one
		// 128-bit load of column indices and two 128-bit loads of values
		const int lo = rp[0], hi = rp[rb - ra];
		const int nrow = rb - ra;
		const long long start = (long long)lo & ~7LL;
		for (long long u = start + 8LL * tid; u < hi; u += 8LL * blockDim.x) {
			__align__(16) short c8[8];
			__align__(16) float v8[8];
			if (u + 8 <= nnz_total) {
				*reinterpret_cast<int4*>(c8) = __ldg(reinterpret_cast<const int4*>(col + u));
				*reinterpret_cast<float4*>(v8) = __ldg(reinterpret_cast<const float4*>(val + u));
				*reinterpret_cast<float4*>(v8 + 4) = __ldg(reinterpret_cast<const float4*>(val + u + 4));
			} else {
#pragma unroll
				for (int e = 0; e < 8; ++e) {
					bool ok = u + e < nnz_total;
					c8[e] = ok ? col[u + e] : (short)0;
					v8[e] = ok ? val[u + e] : 0.f;
				}
			}
			// row of the first in-range entry by binary search, then walk
			long long first = u < lo ? lo : u;
			int r = 0;
			{
				int a = 0, b = nrow;  // find r with rp[r] <= first < rp[r+1]
				while (b - a > 1) {
					int m = (a + b) >> 1;
					if (rp[m] <= first) a = m; else b = m;
				}
				r = a;
			}
#pragma unroll
			for (int e = 0; e < 8; ++e) {
				long long idx = u + e;
				if (idx < lo || idx >= hi) continue;
				while (rp[r + 1] <= idx) ++r;
				int gr = ra + r;
				tile[(gr - r0 + 1) * tp + (int)c8[e] + 4] = fmaxf(v8[e], FH_FLOOR);
			}
		}
	}
	__syncthreads();
	// 3. stencil + write; pad columns [w, ldw) are written as 0. One work unit = FOUR columns x 8 rows: per tile row one
	// 128-bit shared load + the two neighbours, seven adds for the four horizontal triple sums, the 3x3 window slides down
	// the columns on them, one 128-bit global store per output row (ncu on the one-column version: issue slots 87 % busy,
	// ALU the busiest pipe - the kernel was instruction bound at 2.2 TB/s, not memory bound).
	float* dst = out + (long long)cell * out_cell_stride;
	// OUT16: strides in halves; window column c at plane column c + pad16 (pad16 = 0 or 4 leading zero columns)
	__half* dst16 = reinterpret_cast<__half*>(out) + (long long)cell * out_cell_stride + pad16;
	float sa = 1.f;
	if (OUT16) {
		const unsigned abits = max(amax[cell], __float_as_uint(FH_FLOOR));
		sa = __uint_as_float((267u - (abits >> 23)) << 23);
	}
	const int rows = min(RT, nb - r0);
	const int ngrp = ldw >> 2;
	if (do_conv) {
		const int nstrip = (rows + 7) >> 3;            // <= 4 strips of 8 rows
		const int per = nwarp / 4 > 0 ? nwarp / 4 : 1;  // warps per strip
		for (int strip = warp & 3; strip < nstrip; strip += 4) {
			const int tr0 = strip * 8;
			const int nr = min(8, rows - tr0);
			for (int q = (warp >> 2) * 32 + lane; q < ngrp; q += 32 * per) {
				const int c = 4 * q;
				float* drow = dst + (long long)(r0 + tr0) * ldw + c;
				__half* drow16 = dst16 + (long long)(r0 + tr0) * ld16 + c;
				const float* t = tile + tr0 * tp + c + 4;  // tile row tr0 = global row r0 + tr0 - 1
				float4 h0, h1;
				{
					const float4 a = *reinterpret_cast<const float4*>(t);
					const float l = t[-1], r_ = t[4];
					const float s01 = a.x + a.y, s23 = a.z + a.w;
					h0.x = l + s01; h0.y = s01 + a.z; h0.z = a.y + s23; h0.w = s23 + r_;
				}
				{
					const float4 a = *reinterpret_cast<const float4*>(t + tp);
					const float l = t[tp - 1], r_ = t[tp + 4];
					const float s01 = a.x + a.y, s23 = a.z + a.w;
					h1.x = l + s01; h1.y = s01 + a.z; h1.z = a.y + s23; h1.w = s23 + r_;
				}
				const bool m0 = c < w, m1 = c + 1 < w, m2 = c + 2 < w, m3 = c + 3 < w;
#pragma unroll
				for (int k = 0; k < 8; ++k) {
					if (k < nr) {
						const float* tn = t + (k + 2) * tp;
						const float4 a = *reinterpret_cast<const float4*>(tn);
						const float l = tn[-1], r_ = tn[4];
						const float s01 = a.x + a.y, s23 = a.z + a.w;
						float4 h2;
						h2.x = l + s01; h2.y = s01 + a.z; h2.z = a.y + s23; h2.w = s23 + r_;
						float4 o;  // 1 ulp from sum / 9: an IEEE division would be ~9 instructions per output
						o.x = m0 ? fmaxf(((h0.x + h1.x) + h2.x) * (1.0f / 9.0f), FH_FLOOR) : 0.f;
						o.y = m1 ? fmaxf(((h0.y + h1.y) + h2.y) * (1.0f / 9.0f), FH_FLOOR) : 0.f;
						o.z = m2 ? fmaxf(((h0.z + h1.z) + h2.z) * (1.0f / 9.0f), FH_FLOOR) : 0.f;
						o.w = m3 ? fmaxf(((h0.w + h1.w) + h2.w) * (1.0f / 9.0f), FH_FLOOR) : 0.f;
						if (OUT16) {
							store4_16(drow16 + (long long)k * ld16, lo_plane, o, sa);
							if (pad16 && c == 0) store4_16(drow16 + (long long)k * ld16 - 4, lo_plane, make_float4(0.f, 0.f, 0.f, 0.f), 0.f);
						} else *reinterpret_cast<float4*>(drow + (long long)k * ldw) = o;
						h0 = h1; h1 = h2;
					}
				}
			}
		}
	} else {
		for (int tr = warp; tr < rows; tr += nwarp) {
			float4* drow = reinterpret_cast<float4*>(dst + (long long)(r0 + tr) * ldw);
			__half* drow16 = dst16 + (long long)(r0 + tr) * ld16;
			const float4* trow = reinterpret_cast<const float4*>(tile + (tr + 1) * tp + 4);
			for (int q = lane; q < ngrp; q += 32) {  // pad columns hold the background's zeros
				if (OUT16) {
					store4_16(drow16 + 4 * q, lo_plane, trow[q], sa);
					if (pad16 && q == 0) store4_16(drow16 - 4, lo_plane, make_float4(0.f, 0.f, 0.f, 0.f), 0.f);
				} else drow[q] = trow[q];
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------
// K4: second-order affinity S2 = A A^T (diagonal ignored) + first-order block of A -> column
// stochastic P, in place over S2. grid (ncell), thread per column.
// ---------------------------------------------------------------------------------------------
// Each thread keeps its <= 32 rows of the column in registers, so S2 and the diagonal block of A are read
// once and P (and Q1 = 0.5 P + 0.5 I, the first RWR step, when Q1 != nullptr) are written once.
template <int MAXR>  // rows per thread: nb <= 8 * MAXR
__global__ void __launch_bounds__(256)
transition_kernel(const float* __restrict__ A, long long a_cell_stride, int ldw, int s,
                  float* __restrict__ SP, int nb, int ldp, float* __restrict__ Q1) {
	// 256 threads = 8 row groups x 32 columns; a column tile of 32 is reduced over the 8 groups in smem
	__shared__ float red[3][8][33];
	const int cell = blockIdx.x;
	const float* a = A + (long long)cell * a_cell_stride + s;
	float* p = SP + (long long)cell * nb * ldp;
	float* q1 = Q1 ? Q1 + (long long)cell * nb * ldp : nullptr;
	const int cx = threadIdx.x & 31, g = threadIdx.x >> 5;
	for (int j0 = 0; j0 < nb; j0 += 32) {
		const int j = j0 + cx;
		const bool ok = j < nb;
		float f[MAXR], h[MAXR];
		float cs1 = 0.f, cs2 = 0.f;
#pragma unroll
		for (int e = 0; e < MAXR; ++e) {
			const int i = g + 8 * e;
			f[e] = 0.f; h[e] = 0.f;
			if (ok && i < nb) {
				f[e] = a[(long long)i * ldw + j];
				h[e] = (i != j) ? p[i * ldp + j] : 0.f;
			}
			cs1 += f[e]; cs2 += h[e];
		}
		red[0][g][cx] = cs1; red[1][g][cx] = cs2;
		__syncthreads();
		cs1 = 0.f; cs2 = 0.f;
#pragma unroll
		for (int k = 0; k < 8; ++k) { cs1 += red[0][k][cx]; cs2 += red[1][k][cx]; }
		cs1 += FH_EPS; cs2 += FH_EPS;
		float csl = 0.f;
#pragma unroll
		for (int e = 0; e < MAXR; ++e) {
			const int i = g + 8 * e;
			float l = (f[e] / cs1) * 0.75f + ((i != j) ? (h[e] / cs2) * 0.25f : 0.f);
			if (!(ok && i < nb)) l = 0.f;
			f[e] = l;
			csl += l;
		}
		red[2][g][cx] = csl;
		__syncthreads();
		csl = 0.f;
#pragma unroll
		for (int k = 0; k < 8; ++k) csl += red[2][k][cx];
		// a zero column is unreachable after the 1e-8 floor; kept for parity (partial_rwr.py:96-97)
		const bool empty = (csl == 0.f);
		if (empty) csl += 1.f;
		csl += FH_EPS;
#pragma unroll
		for (int e = 0; e < MAXR; ++e) {
			const int i = g + 8 * e;
			if (ok && i < nb) {
				float l = f[e];
				if (empty && i == j) l += 1.f;
				const float pv = l / csl;
				p[i * ldp + j] = pv;
				if (q1) q1[i * ldp + j] = 0.5f * pv + ((i == j) ? 0.5f : 0.f);
			}
		}
		__syncthreads();
	}
	if (q1)  // pad columns of Q1
		for (int t = threadIdx.x; t < nb * (ldp - nb); t += blockDim.x) q1[(t / (ldp - nb)) * ldp + nb + t % (ldp - nb)] = 0.f;
}

// Q = 0.5 * P + 0.5 * I  (first RWR step: Q0 = I so bmm(Q0, P) = P exactly)
__global__ void first_step_kernel(const float* __restrict__ P, float* __restrict__ Q, int nb, int ldp,
                                  long long total) {
	long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (i >= total) return;
	int c = (int)(i % ldp);
	int r = (int)((i / ldp) % nb);
	float v = 0.5f * P[i];
	if (r == c) v += 0.5f;
	Q[i] = (c < nb) ? v : 0.f;
}

__global__ void identity_kernel(float* __restrict__ Q, int nb, int ldp, long long total) {
	long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
	if (i >= total) return;
	int c = (int)(i % ldp);
	int r = (int)((i / ldp) % nb);
	Q[i] = (r == c) ? 1.f : 0.f;
}

// per-cell Frobenius norm of Qa - Qb (Qa == nullptr: identity)
__global__ void __launch_bounds__(256)
delta_kernel(const float* __restrict__ Qa, const float* __restrict__ Qb, int nb, int ldp,
             float* __restrict__ delta) {
	__shared__ float red[32];
	const int cell = blockIdx.x;
	const long long base = (long long)cell * nb * ldp;
	float acc = 0.f;
	for (int i = threadIdx.x; i < nb * ldp; i += blockDim.x) {
		int c = i % ldp, r = i / ldp;
		if (c >= nb) continue;
		float a = Qa ? Qa[base + i] : (r == c ? 1.f : 0.f);
		float d = a - Qb[base + i];
		acc += d * d;
	}
	acc = fh_block_sum(acc, red);
	if (threadIdx.x == 0) delta[cell] = sqrtf(acc);
}

// do_col: Q <- rownorm(clamp0((Q + Q^T)/2))   (partial_rwr.py:132-134). warp per row.
__global__ void __launch_bounds__(256)
symnorm_kernel(const float* __restrict__ Q, float* __restrict__ out, int nb, int ldp) {
	const int cell = blockIdx.x;
	const float* q = Q + (long long)cell * nb * ldp;
	float* o = out + (long long)cell * nb * ldp;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
	for (int i = wid; i < nb; i += nw) {
		float rs = 0.f;
		for (int j = lane; j < nb; j += 32) {
			float v = fmaxf((q[i * ldp + j] + q[j * ldp + i]) * 0.5f, 0.f);
			rs += v;
		}
		rs = fh_warp_sum(rs) + FH_EPS;
		for (int j = lane; j < ldp; j += 32) {
			float v = 0.f;
			if (j < nb) v = fmaxf((q[i * ldp + j] + q[j * ldp + i]) * 0.5f, 0.f) / rs;
			o[i * ldp + j] = v;
		}
	}
}

__global__ void colsum_accum_kernel(const float* __restrict__ x, int nb, int w, int ldw,
                                    long long cell_stride, float* __restrict__ cov, long long cov_ld) {
	const int cell = blockIdx.y;
	const int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= w) return;
	const float* p = x + (long long)cell * cell_stride + c;
	float s = 0.f;
	for (int r = 0; r < nb; ++r) s += p[(long long)r * ldw];
	cov[(long long)cell * cov_ld + c] += s;
}

__global__ void avgpool_kernel(const float* __restrict__ x, int nb, int w, int ldw, long long cell_stride,
                               int ll, int orow, int ocol, float* __restrict__ out, long long out_cell_stride) {
	const int cell = blockIdx.y;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= orow * ocol) return;
	int pr = i / ocol, pc = i - pr * ocol;
	const float* p = x + (long long)cell * cell_stride + (long long)(pr * ll) * ldw + pc * ll;
	float s = 0.f;
	for (int a = 0; a < ll; ++a)
		for (int b = 0; b < ll; ++b) s += p[(long long)a * ldw + b];
	out[(long long)cell * out_cell_stride + i] = s / (float)(ll * ll);
}

__global__ void __launch_bounds__(256)
sqnorm_kernel(const float* __restrict__ x, const float* __restrict__ y, long long rows, long long cols,
              long long ldx, long long ldy, double* __restrict__ acc) {
	__shared__ double red[32];
	double a = 0.0;
	const long long total = rows * cols;
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
	     i += (long long)gridDim.x * blockDim.x) {
		long long r = i / cols, c = i - r * cols;
		float xv = x[r * ldx + c];
		float yv = y ? y[r * ldy + c] : xv;
		a += (double)xv * (double)yv;
	}
	a = fh_block_sum(a, red);
	if (threadIdx.x == 0) atomicAdd(acc, a);
}

// per cell: largest floored value of its CSR rows of the block -> amax[cell] (bits of a non-negative float; 0 = no entries).
// One CTA per cell (a cell's slice is a few thousand entries): no atomics, no initialisation pass.
__global__ void __launch_bounds__(128)
csr_absmax_kernel(const int32_t* __restrict__ rowptr, const float* __restrict__ val, int cell0, int nb,
                  unsigned* __restrict__ amax) {
	__shared__ float red[4];
	const long long r0 = (long long)(cell0 + blockIdx.x) * nb;
	const int lo = rowptr[r0], hi = rowptr[r0 + nb];
	float m = 0.f;
	for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) m = fmaxf(m, __ldg(val + i));
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
	__syncthreads();
	if (threadIdx.x == 0) {
		m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
		amax[blockIdx.x] = m > 0.f ? __float_as_uint(fmaxf(m, FH_FLOOR)) : 0u;
	}
}

}  // namespace
extern "C" void fh_count_tc_fallback(void);
size_t fh_rwr_chain_scratch_bytes();
int fh_rwr_chain16_pad(int s);
int fh_rwr_chain16(const void* Ahi, const unsigned* amax, float* out, int nb, int w, int ldw, int ld16, int s, int k,
                   int ncell, long long a_cell_stride, long long out_cell_stride, const float* bin_cov, long long bin_cov_ld,
                   void* stream);
int fh_rwr_chain(const float* P, const float* A, float* out, int nb, int w, int ldw, int ldp, int s, int k, int ncell,
                 long long p_cell_stride, long long a_cell_stride, long long out_cell_stride, float* scratch,
                 void* stream);
namespace {

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct RwrWs {
	float *A, *P, *Q0, *Q1, *delta, *scratch;
	unsigned* amax;
	size_t bytes;
};

RwrWs carve(const fh_rwr_desc* d, void* ws) {
	RwrWs r;
	const int ldp = (d->nb + 3) & ~3;
	// the conv'd panel: fp32 rows of ldw floats, or two binary16 planes with rows of round_up(w, 8) halves
	size_t a = align_up((size_t)d->ncell * d->nb * ((d->ldw + 4 + 7) & ~7) * 4, 256);
	size_t p = align_up((size_t)d->ncell * d->nb * ldp * 4, 256);
	char* b = (char*)ws;
	r.A = (float*)b; b += a;
	r.P = (float*)b; b += p;
	r.Q0 = (float*)b; b += p;
	r.Q1 = (float*)b; b += p;
	r.delta = (float*)b; b += align_up((size_t)d->ncell * 4, 256);
	r.scratch = (float*)b; b += align_up(fh_rwr_chain_scratch_bytes(), 256);
	r.amax = (unsigned*)b; b += align_up((size_t)d->ncell * 4, 256);  // 3xFP16 path: one scale word per cell
	r.bytes = (size_t)(b - (char*)ws);
	return r;
}

int gemm_f32(int use_tc, int M, int N, int K, int batch, const float* A, long long sa_m, long long sa_k,
             long long ba, const float* B, long long sb_k, long long sb_n, long long bb, float* C,
             long long ldc, long long bc, double alpha, int epi, double diag, const float* cscale,
             long long cs_batch, int recip, void* stream) {
	fh_gemm_desc g;
	memset(&g, 0, sizeof(g));
	g.M = M; g.N = N; g.K = K; g.batch = batch;
	g.sa_m = sa_m; g.sa_k = sa_k; g.sb_k = sb_k; g.sb_n = sb_n; g.ldc = ldc;
	g.batch_a = ba; g.batch_b = bb; g.batch_c = bc;
	g.alpha = alpha; g.beta = 0.0; g.dtype = use_tc ? FH_GEMM_TF32X3 : FH_GEMM_F32;
	g.epilogue = epi; g.diag = diag;
	g.cscale = cscale; g.cscale_batch = cs_batch; g.cscale_recip = recip;
	// batches are limited to 65535 per launch by gridDim.z
	for (int b0 = 0; b0 < batch; b0 += 32768) {
		int nbt = batch - b0 < 32768 ? batch - b0 : 32768;
		g.batch = nbt;
		g.cscale = cscale ? cscale + (long long)b0 * cs_batch : nullptr;
		int rc = fh_gemm_batched(&g, A + (long long)b0 * ba, B + (long long)b0 * bb, C + (long long)b0 * bc, stream);
		if (rc) return rc;
	}
	return FH_OK;
}

// FH_RWR_FUSED: 2 (default) S2 + transition + steps + Q A in one kernel; 1 steps + Q A fused; 0 per-step
// kernels (A/B measurements)
int rwr_fused_level() {
	static int v = -1;
	if (v < 0) { const char* e = getenv("FH_RWR_FUSED"); v = e ? atoi(e) : 2; }
	return v;
}

// FH_RWR_F16: 1 (default) the fused kernel works on binary16 operand pairs (3xFP16, fh_rwr_chain16.cu: the panel is
// densified straight into two binary16 planes); 0 the 3xTF32 kernel (fh_rwr_chain.cu)
int rwr_f16_enabled() {
	static int v = -1;
	if (v < 0) { const char* e = getenv("FH_RWR_F16"); v = e ? atoi(e) : 1; }
	return v;
}

// the RWR pipeline from the conv'd panel A (in ws.A, or already in `out` when !do_rwr)
int rwr_from_panel(const fh_rwr_desc* d, const RwrWs& ws, const float* bin_cov, long long bin_cov_ld,
                   float* out, long long out_cell_stride, int* host_n_iter, cudaStream_t st) {
	const int nb = d->nb, w = d->w, ldw = d->ldw, nc = d->ncell;
	const int ldp = (nb + 3) & ~3;
	const long long acs = (long long)nb * ldw, pcs = (long long)nb * ldp;
	const long long ptotal = (long long)nc * pcs;
	const int tc = d->use_tensor_cores;
	int rc;
	// forced step count without do_col on the tensor cores: one fused kernel (fh_rwr_chain.cu)
	const int fused = (tc && d->k >= 1 && !d->do_col && nb <= 128) ? rwr_fused_level() : 0;
	FH_CHECK_ARG(nb <= 256, "fh_rwr: bin block of %d rows (max 256: recommend_bs_bin, FastHigashi_Wrapper.py:501)", nb);
	if (fused >= 2) {
		rc = fh_rwr_chain(nullptr, ws.A, out, nb, w, ldw, ldp, d->s, d->k, nc, pcs, acs, out_cell_stride, ws.scratch, st);
		if (rc == FH_OK) {
			if (host_n_iter) *host_n_iter = d->k;
			return FH_OK;
		}
		if (rc != FH_ERR_UNSUPPORTED) return rc;
		fh_count_tc_fallback();
	}
	// S2 = A A^T  (partial_rwr.py:85)
	rc = gemm_f32(tc, nb, nb, w, nc, ws.A, ldw, 1, acs, ws.A, 1, ldw, acs, ws.P, ldp, pcs, 1.0, FH_EPI_NONE, 0.0,
	              nullptr, 0, 0, st);
	if (rc) return rc;
	const bool fuse_q1 = d->k >= 1 && fused != 1;  // forced mode: Q1 comes out of the transition kernel
	if (nb <= 128) transition_kernel<16><<<nc, 256, 0, st>>>(ws.A, acs, ldw, d->s, ws.P, nb, ldp, fuse_q1 ? ws.Q0 : nullptr);
	else transition_kernel<32><<<nc, 256, 0, st>>>(ws.A, acs, ldw, d->s, ws.P, nb, ldp, fuse_q1 ? ws.Q0 : nullptr);
	FH_LAUNCH_CHECK();
	if (fused == 1) {
		rc = fh_rwr_chain(ws.P, ws.A, out, nb, w, ldw, ldp, d->s, d->k, nc, pcs, acs, out_cell_stride, nullptr, st);
		if (rc == FH_OK) {
			if (host_n_iter) *host_n_iter = d->k;
			return FH_OK;
		}
		if (rc != FH_ERR_UNSUPPORTED) return rc;
		fh_count_tc_fallback();
		first_step_kernel<<<fh_cdiv(ptotal, 256), 256, 0, st>>>(ws.P, ws.Q0, nb, ldp, ptotal);  // Q1 for the GEMM chain
		FH_LAUNCH_CHECK();
	}
	float* Q = ws.Q0;
	float* Qn = ws.Q1;
	int n_iter = 0;
	const int tpb = 256;
	const int nblk = fh_cdiv(ptotal, tpb);
	if (d->k >= 0) {
		if (d->k == 0) {
			identity_kernel<<<nblk, tpb, 0, st>>>(Q, nb, ldp, ptotal);
			FH_LAUNCH_CHECK();
		} else {
			for (int it = 1; it < d->k; ++it) {
				rc = gemm_f32(tc, nb, nb, nb, nc, Q, ldp, 1, pcs, ws.P, ldp, 1, pcs, Qn, ldp, pcs, 0.5,
				              FH_EPI_DIAG_ADD, 0.5, nullptr, 0, 0, st);
				if (rc) return rc;
				float* t = Q; Q = Qn; Qn = t;
			}
		}
		n_iter = d->k;
	} else {
		// auto-stop (partial_rwr.py:99-126): apply the step, then stop once the largest per-cell
		// Frobenius change is < 0.01; the reported count excludes the breaking step.
		float* hdelta = (float*)malloc(sizeof(float) * nc);
		if (!hdelta) { fh_set_error("fh_rwr: host malloc failed"); return FH_ERR_ARG; }
		const float* prev = nullptr;  // identity
		int count = 0;
		for (int it = 0; it < 60; ++it) {
			if (it == 0) {
				first_step_kernel<<<nblk, tpb, 0, st>>>(ws.P, Qn, nb, ldp, ptotal);
				FH_LAUNCH_CHECK();
			} else {
				rc = gemm_f32(tc, nb, nb, nb, nc, Q, ldp, 1, pcs, ws.P, ldp, 1, pcs, Qn, ldp, pcs, 0.5,
				              FH_EPI_DIAG_ADD, 0.5, nullptr, 0, 0, st);
				if (rc) { free(hdelta); return rc; }
			}
			delta_kernel<<<nc, 256, 0, st>>>(prev, Qn, nb, ldp, ws.delta);
			fh_count_launch(1);
			cudaError_t e = cudaMemcpyAsync(hdelta, ws.delta, sizeof(float) * nc, cudaMemcpyDeviceToHost, st);
			if (e == cudaSuccess) e = cudaStreamSynchronize(st);
			if (e != cudaSuccess) { free(hdelta); fh_set_error("fh_rwr auto-stop: %s", cudaGetErrorString(e)); return FH_ERR_CUDA; }
			float mx = 0.f;
			for (int c = 0; c < nc; ++c) mx = hdelta[c] > mx ? hdelta[c] : mx;
			float* t = Q; Q = Qn; Qn = t;
			prev = Q;
			if (mx < 0.01f) break;
			count++;
		}
		free(hdelta);
		n_iter = count;
	}
	if (host_n_iter) *host_n_iter = n_iter;
	const float* Qfin = Q;
	if (d->do_col) {
		symnorm_kernel<<<nc, 256, 0, st>>>(Q, Qn, nb, ldp);
		FH_LAUNCH_CHECK();
		Qfin = Qn;
	}
	// x = Q A (/ bin_cov per window column when do_col)   (partial_rwr.py:135,138)
	rc = gemm_f32(tc, nb, w, nb, nc, Qfin, ldp, 1, pcs, ws.A, ldw, 1, acs, out, ldw, out_cell_stride, 1.0,
	              FH_EPI_NONE, 0.0, d->do_col ? bin_cov : nullptr, bin_cov_ld, 1, st);
	if (rc) return rc;
	if (ldw > w) {
		if (out_cell_stride == acs) {
			FH_CUDA(cudaMemset2DAsync(out + w, (size_t)ldw * 4, 0, (size_t)(ldw - w) * 4, (size_t)nc * nb, st));
		} else {
			for (int c = 0; c < nc; ++c)
				FH_CUDA(cudaMemset2DAsync(out + (long long)c * out_cell_stride + w, (size_t)ldw * 4, 0,
				                          (size_t)(ldw - w) * 4, (size_t)nb, st));
		}
	}
	return FH_OK;
}

int check_desc(const fh_rwr_desc* d) {
	FH_CHECK_ARG(d != nullptr, "fh_rwr: null descriptor");
	FH_CHECK_ARG(d->nb > 0 && d->w > 0 && d->ncell >= 0, "fh_rwr: bad sizes nb=%d w=%d ncell=%d", d->nb, d->w, d->ncell);
	FH_CHECK_ARG(d->ldw >= d->w && d->ldw % 4 == 0, "fh_rwr: ldw=%d must be >= w=%d and a multiple of 4", d->ldw, d->w);
	FH_CHECK_ARG(d->s >= 0 && d->s + d->nb <= d->w, "fh_rwr: diagonal block [%d,%d) outside window %d", d->s, d->s + d->nb, d->w);
	FH_CHECK_ARG(d->ncell <= 65535, "fh_rwr: ncell %d > 65535 per call", d->ncell);
	FH_CHECK_ARG(densify_smem_bytes(d->ldw) <= 200 * 1024, "fh_rwr: window %d too wide", d->w);
	return FH_OK;
}

int launch_densify(const fh_rwr_desc* d, bool from_dense, const int32_t* rowptr, const int16_t* col,
                   const float* val, const float* dense_in, long long in_cs, int do_conv, float* out,
                   long long out_cs, cudaStream_t st) {
	size_t smem = densify_smem_bytes(d->ldw);
	FH_CHECK_ARG(((uintptr_t)out & 15) == 0 && (out_cs & 3) == 0, "fh_rwr: output panels must be 16-byte aligned with a cell stride that is a multiple of 4 floats");
	dim3 grid(fh_cdiv(d->nb, RT), d->ncell);
	const int tmr = fh_time_begin(FH_TIME_DENSIFY, st);
	if (from_dense) {
		if (smem > 48 * 1024) FH_CUDA(cudaFuncSetAttribute(densify_conv_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		densify_conv_kernel<true, false><<<grid, 256, smem, st>>>(nullptr, nullptr, nullptr, 0, dense_in, in_cs, 0, d->nb, d->w,
		                                                         d->ldw, do_conv, out, out_cs, 0, 0, nullptr, 0);
	} else {
		if (smem > 48 * 1024) FH_CUDA(cudaFuncSetAttribute(densify_conv_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		densify_conv_kernel<false, false><<<grid, 256, smem, st>>>(rowptr, col, val, d->nnz, nullptr, 0, d->cell0, d->nb, d->w,
		                                                          d->ldw, do_conv, out, out_cs, 0, 0, nullptr, 0);
	}
	fh_time_end(tmr, st);
	FH_LAUNCH_CHECK();
	return FH_OK;
}

}  // namespace

extern "C" size_t fh_rwr_workspace_bytes(const fh_rwr_desc* d) {
	if (!d) return 0;
	return carve(d, nullptr).bytes;
}

extern "C" int fh_densify(const fh_rwr_desc* d, const int32_t* rowptr, const int16_t* col, const float* val,
                          float* out, long long out_cell_stride, void* stream) {
	int rc = check_desc(d);
	if (rc) return rc;
	if (d->ncell == 0) return FH_OK;
	return launch_densify(d, false, rowptr, col, val, nullptr, 0, 0, out, out_cell_stride, (cudaStream_t)stream);
}

extern "C" int fh_rwr_batched(const fh_rwr_desc* d, const int32_t* rowptr, const int16_t* col, const float* val,
                              const float* bin_cov, long long bin_cov_ld, float* out, long long out_cell_stride,
                              void* workspace, size_t workspace_bytes, int* host_n_iter, void* stream) {
	int rc = check_desc(d);
	if (rc) return rc;
	if (host_n_iter) *host_n_iter = 0;
	if (d->ncell == 0) return FH_OK;
	cudaStream_t st = (cudaStream_t)stream;
	const int conv = d->do_conv && d->nb > 1;  // partial_rwr.py:77
	if (!d->do_rwr)  // conv only, or neither: the (floored) densified block
		return launch_densify(d, false, rowptr, col, val, nullptr, 0, conv, out, out_cell_stride, st);
	FH_CHECK_ARG(workspace != nullptr && workspace_bytes >= fh_rwr_workspace_bytes(d),
	             "fh_rwr_batched: workspace too small (%zu < %zu)", workspace_bytes, fh_rwr_workspace_bytes(d));
	FH_CHECK_ARG(!d->do_col || bin_cov != nullptr, "fh_rwr_batched: do_col needs bin_cov");
	RwrWs ws = carve(d, workspace);
	// forced step count on the tensor cores (every call of the ALS sweep; do_col from two steps on): 3xFP16 fused kernel
	if (d->use_tensor_cores && d->k >= (d->do_col ? 2 : 1) && d->nb <= 128 && rwr_fused_level() >= 2 && rwr_f16_enabled() &&
	    ((uintptr_t)out & 15) == 0 && (out_cell_stride & 3) == 0 && (d->s & 3) == 0) {
		const int pad16 = fh_rwr_chain16_pad(d->s);  // 0 or 4: the diagonal block starts at a multiple of 8 plane columns
		const int ld16 = (d->ldw + pad16 + 7) & ~7;
		const long long acs16 = (long long)d->nb * ld16;
		csr_absmax_kernel<<<d->ncell, 128, 0, st>>>(rowptr, val, d->cell0, d->nb, ws.amax);
		FH_LAUNCH_CHECK();
		const size_t smem = densify_smem_bytes(d->ldw);
		if (smem > 48 * 1024) FH_CUDA(cudaFuncSetAttribute(densify_conv_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		dim3 grid(fh_cdiv(d->nb, RT), d->ncell);
		const int tmr = fh_time_begin(FH_TIME_DENSIFY, st);
		densify_conv_kernel<false, true><<<grid, 256, smem, st>>>(rowptr, col, val, d->nnz, nullptr, 0, d->cell0, d->nb, d->w, d->ldw,
		                                                         conv, ws.A, acs16, ld16, (long long)d->ncell * acs16, ws.amax, pad16);
		fh_time_end(tmr, st);
		FH_LAUNCH_CHECK();
		rc = fh_rwr_chain16(ws.A, ws.amax, out, d->nb, d->w, d->ldw, ld16, d->s, d->k, d->ncell, acs16, out_cell_stride,
		                    d->do_col ? bin_cov : nullptr, bin_cov_ld, st);
		if (rc == FH_OK) {
			if (host_n_iter) *host_n_iter = d->k;
			return FH_OK;
		}
		if (rc != FH_ERR_UNSUPPORTED) return rc;
		fh_count_tc_fallback();
	}
	rc = launch_densify(d, false, rowptr, col, val, nullptr, 0, conv, ws.A, (long long)d->nb * d->ldw, st);
	if (rc) return rc;
	return rwr_from_panel(d, ws, bin_cov, bin_cov_ld, out, out_cell_stride, host_n_iter, st);
}

extern "C" int fh_rwr_dense(const fh_rwr_desc* d, float* x, long long cell_stride, const float* bin_cov,
                            long long bin_cov_ld, void* workspace, size_t workspace_bytes, int* host_n_iter,
                            void* stream) {
	int rc = check_desc(d);
	if (rc) return rc;
	if (host_n_iter) *host_n_iter = 0;
	if (d->ncell == 0 || !(d->do_conv || d->do_rwr)) return FH_OK;
	cudaStream_t st = (cudaStream_t)stream;
	FH_CHECK_ARG(workspace != nullptr && workspace_bytes >= fh_rwr_workspace_bytes(d),
	             "fh_rwr_dense: workspace too small (%zu < %zu)", workspace_bytes, fh_rwr_workspace_bytes(d));
	FH_CHECK_ARG(!d->do_col || !d->do_rwr || bin_cov != nullptr, "fh_rwr_dense: do_col needs bin_cov");
	RwrWs ws = carve(d, workspace);
	const int conv = d->do_conv && d->nb > 1;
	rc = launch_densify(d, true, nullptr, nullptr, nullptr, x, cell_stride, conv, ws.A, (long long)d->nb * d->ldw, st);
	if (rc) return rc;
	if (!d->do_rwr) {
		FH_CUDA(cudaMemcpy2DAsync(x, (size_t)cell_stride * 4, ws.A, (size_t)d->nb * d->ldw * 4, (size_t)d->nb * d->ldw * 4,
		                          d->ncell, cudaMemcpyDeviceToDevice, st));
		return FH_OK;
	}
	return rwr_from_panel(d, ws, bin_cov, bin_cov_ld, x, cell_stride, host_n_iter, st);
}

extern "C" int fh_colsum_accum(const float* x, int ncell, int nb, int w, int ldw, long long cell_stride,
                               float* cov, long long cov_ld, void* stream) {
	if (ncell <= 0) return FH_OK;
	FH_CHECK_ARG(ncell <= 65535, "fh_colsum_accum: ncell > 65535");
	dim3 grid(fh_cdiv(w, 128), ncell);
	colsum_accum_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(x, nb, w, ldw, cell_stride, cov, cov_ld);
	FH_LAUNCH_CHECK();
	return FH_OK;
}

extern "C" int fh_avgpool(const float* x, int ncell, int nb, int w, int ldw, long long cell_stride, int ll,
                          float* out, long long out_cell_stride, void* stream) {
	if (ncell <= 0) return FH_OK;
	FH_CHECK_ARG(ll >= 1 && ncell <= 65535, "fh_avgpool: bad arguments");
	int orow = nb / ll, ocol = w / ll;
	if (orow * ocol == 0) return FH_OK;
	dim3 grid(fh_cdiv(orow * ocol, 128), ncell);
	avgpool_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(x, nb, w, ldw, cell_stride, ll, orow, ocol, out, out_cell_stride);
	FH_LAUNCH_CHECK();
	return FH_OK;
}

extern "C" int fh_sqnorm_accum(const float* x, long long rows, long long cols, long long ld, double* acc, void* stream) {
	if (rows * cols <= 0) return FH_OK;
	int grid = (int)((rows * cols + 256 * 8 - 1) / (256 * 8));
	if (grid > 148 * 8) grid = 148 * 8;
	sqnorm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, nullptr, rows, cols, ld, ld, acc);
	FH_LAUNCH_CHECK();
	return FH_OK;
}

extern "C" int fh_dot_accum(const float* x, const float* y, long long rows, long long cols, long long ldx,
                            long long ldy, double* acc, void* stream) {
	if (rows * cols <= 0) return FH_OK;
	int grid = (int)((rows * cols + 256 * 8 - 1) / (256 * 8));
	if (grid > 148 * 8) grid = 148 * 8;
	sqnorm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, rows, cols, ldx, ldy, acc);
	FH_LAUNCH_CHECK();
	return FH_OK;
}
