// Register-blocked one-sided Jacobi for the per-bin polar step (included by fh_polar.cu; uses jacobi_tan).
//
// Why: the scalar kernel (a quarter-warp per row pair, rows streamed from shared memory) is shared-memory bound - every
// row crosses shared memory ~2.5 times per round-robin step, n - 1 steps per sweep (ncu, round 1: fp64 pipe 21 %,
// 1.25 eligible warps per scheduler). Here a WARP owns a pair of row blocks (2 x 4 rows) and keeps them in registers
// (lane l holds elements l, l + 32, ... of each row: 8 x PL doubles), performs all 4 x 4 cross rotations on them - four
// rounds of four independent rotations whose dot products are reduced together (reduce4: 20 shuffles for four sums) -
// and the blocks move between warps in the odd-even transposition ordering: N = n / 4 blocks on a line, even steps pair
// positions (2g, 2g + 1), odd steps (2g + 1, 2g + 2), the two blocks of a pair swap positions after their visit. Every
// pair of blocks meets exactly once in N steps, and a warp keeps ONE of its two blocks from a step to the next, so a
// step moves one block per warp through shared memory: per sweep each row crosses shared memory ~N times instead of
// ~2.5 (n - 1) times (n = 144: 36 vs 357). The pairs inside a block are rotated in the first step of a sweep.
// Same rotation rule, thresholds and stopping test as the scalar sweeps (sweep counts equal on graded spectra to
// kappa = 2e7: numpy model of both orderings).
#pragma once

constexpr int kRB = 4;  // rows per block

__host__ __device__ inline int rb_blocks(int n) { const int N = (n + kRB - 1) / kRB; return N + (N & 1); }
__host__ __device__ inline int rb_rows(int n) { return rb_blocks(n) * kRB; }      // zero rows appended
__host__ __device__ inline int rb_pl(int n) { return (n + 31) >> 5; }              // row elements per lane
__host__ __device__ inline int rb_ld(int n) { return rb_pl(n) * 32 + 1; }          // odd pitch: column walks of the Cholesky phase hit all banks
__host__ __device__ inline int rb_threads(int n) { return 32 * (rb_blocks(n) / 2); }

// Block standing at line position `pos` before global step s (0 <= s < 2N): blocks that start on even positions walk
// right, the others left, both bounce off the ends (dwelling one step there): position f(p0 +- s), f(x) = x mod 2N
// folded at N. Exactly one of the four candidates is a valid start.
__device__ __forceinline__ int rb_block_at(const int pos, const int s, const int N) {
	const int M = 2 * N;
	int a = pos - s; if (a < 0) a += M;
	if (a < N && !(a & 1)) return a;
	int b = M - 1 - pos - s; if (b < 0) b += M;
	if (b < N && !(b & 1)) return b;
	int c = pos + s; if (c >= M) c -= M;
	if (c < N && (c & 1)) return c;
	int d = M - 1 - pos + s; if (d >= M) d -= M;
	return d;
}

// Four warp-wide sums at once: two exchange stages hand each lane one of the four values to finish, three butterfly
// stages finish it, four broadcasts return all sums to every lane (10 fp64 shuffles instead of 20; fixed order).
__device__ __forceinline__ void rb_reduce4(double& g0, double& g1, double& g2, double& g3, const int lane) {
	const unsigned full = 0xffffffffu;
	const bool h16 = lane & 16, h8 = lane & 8;
	double k0 = h16 ? g2 : g0, k1 = h16 ? g3 : g1;
	k0 += __shfl_xor_sync(full, h16 ? g0 : g2, 16);
	k1 += __shfl_xor_sync(full, h16 ? g1 : g3, 16);
	double k = h8 ? k1 : k0;
	k += __shfl_xor_sync(full, h8 ? k0 : k1, 8);
	k += __shfl_xor_sync(full, k, 4);
	k += __shfl_xor_sync(full, k, 2);
	k += __shfl_xor_sync(full, k, 1);
	g0 = __shfl_sync(full, k, 0);
	g1 = __shfl_sync(full, k, 8);
	g2 = __shfl_sync(full, k, 16);
	g3 = __shfl_sync(full, k, 24);
}

template <int PL>
__device__ __forceinline__ double rb_dot(const double (&x)[PL], const double (&y)[PL]) {
	double a = 0.0, b = 0.0;
#pragma unroll
	for (int e = 0; e < PL; ++e) {
		if (e & 1) b = fma(x[e], y[e], b); else a = fma(x[e], y[e], a);
	}
	return a + b;
}

// Rotate rows x (index ip) and y (index iq) given their inner product ga; squared norms tracked in nrm[].
template <int PL>
__device__ __forceinline__ void rb_rotate(double (&x)[PL], double (&y)[PL], const double ga, double* __restrict__ nrm, const int ip,
                                          const int iq, const double tol2, double& worst, const int lane) {
	const double al = nrm[ip], be = nrm[iq];
	const double g2 = ga * ga, mn = fmin(al, be);
	if (g2 > tol2 * mn * mn && ga != 0.0) {
		worst = fmax(worst, g2 / (al * be));
		const double tt = jacobi_tan(al, be, ga);
		const double cs = rsqrt(tt * tt + 1.0), sn = tt * cs;
#pragma unroll
		for (int e = 0; e < PL; ++e) {
			const double a = x[e], c = y[e];
			x[e] = cs * a - sn * c;
			y[e] = sn * a + cs * c;
		}
		if (lane == 0) {
			nrm[ip] = fmax(al - tt * ga, 0.0);
			nrm[iq] = be + tt * ga;
		}
	}
}

// Sweeps over the rows of R (rb_rows(n) x ld, pad rows / columns zero). Warps [0, N / 2) work, the others only meet
// the barriers. Returns the sweep count.
template <int PL>
__device__ __forceinline__ int jacobi_sweeps_rb(double* __restrict__ R, const int n, const int ld, double* __restrict__ nrm,
                                                double* __restrict__ red, const int max_sweeps, const double tol2) {
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
	const int N = rb_blocks(n), M = 2 * N;
	const bool worker = warp < (N >> 1);
	double X[kRB][PL], Y[kRB][PL];
	int curX = -1, curY = -1;  // block held by each register tile
	int s = 0;                 // global step modulo 2N (N even: a sweep starts on an even step)
	int sweep = 0;
	for (; sweep < max_sweeps; ++sweep) {
		double worst = 0.0;  // largest gamma^2 / (alpha beta) met in this sweep
		for (int step = 0; step < N; ++step) {
			const int pL = 2 * warp + (s & 1);
			if (worker && pL + 1 < N) {
				const int bL = rb_block_at(pL, s, N), bR = rb_block_at(pL + 1, s, N);
				if (curX != bL) {
					const double* src = R + (size_t)bL * kRB * ld + lane;
#pragma unroll
					for (int i = 0; i < kRB; ++i)
#pragma unroll
						for (int e = 0; e < PL; ++e) X[i][e] = src[i * ld + 32 * e];
					curX = bL;
				}
				if (curY != bR) {
					const double* src = R + (size_t)bR * kRB * ld + lane;
#pragma unroll
					for (int i = 0; i < kRB; ++i)
#pragma unroll
						for (int e = 0; e < PL; ++e) Y[i][e] = src[i * ld + 32 * e];
					curY = bR;
				}
				const int ix = bL * kRB, iy = bR * kRB;
				if (step == 0) {
					// fresh squared norms of the eight rows, then the six pairs inside each block (three rounds of
					// two disjoint pairs per block)
					double a0 = rb_dot<PL>(X[0], X[0]), a1 = rb_dot<PL>(X[1], X[1]), a2 = rb_dot<PL>(X[2], X[2]), a3 = rb_dot<PL>(X[3], X[3]);
					rb_reduce4(a0, a1, a2, a3, lane);
					double b0 = rb_dot<PL>(Y[0], Y[0]), b1 = rb_dot<PL>(Y[1], Y[1]), b2 = rb_dot<PL>(Y[2], Y[2]), b3 = rb_dot<PL>(Y[3], Y[3]);
					rb_reduce4(b0, b1, b2, b3, lane);
					if (lane == 0) {
						nrm[ix] = a0; nrm[ix + 1] = a1; nrm[ix + 2] = a2; nrm[ix + 3] = a3;
						nrm[iy] = b0; nrm[iy + 1] = b1; nrm[iy + 2] = b2; nrm[iy + 3] = b3;
					}
					__syncwarp();
#define RB_INTRA(P0, Q0, P1, Q1)                                                                                       \
	{                                                                                                                  \
		double g0 = rb_dot<PL>(X[P0], X[Q0]), g1 = rb_dot<PL>(X[P1], X[Q1]), g2 = rb_dot<PL>(Y[P0], Y[Q0]), g3 = rb_dot<PL>(Y[P1], Y[Q1]); \
		rb_reduce4(g0, g1, g2, g3, lane);                                                                              \
		rb_rotate<PL>(X[P0], X[Q0], g0, nrm, ix + P0, ix + Q0, tol2, worst, lane);                                     \
		rb_rotate<PL>(X[P1], X[Q1], g1, nrm, ix + P1, ix + Q1, tol2, worst, lane);                                     \
		rb_rotate<PL>(Y[P0], Y[Q0], g2, nrm, iy + P0, iy + Q0, tol2, worst, lane);                                     \
		rb_rotate<PL>(Y[P1], Y[Q1], g3, nrm, iy + P1, iy + Q1, tol2, worst, lane);                                     \
		__syncwarp();                                                                                                  \
	}
					RB_INTRA(0, 1, 2, 3)
					RB_INTRA(0, 2, 1, 3)
					RB_INTRA(0, 3, 1, 2)
#undef RB_INTRA
				}
				// the 16 cross pairs: round r rotates (X[i], Y[(i + r) & 3]), i = 0..3 - disjoint rows, one reduction
#define RB_CROSS(RR)                                                                                                   \
	{                                                                                                                  \
		double g0 = rb_dot<PL>(X[0], Y[(0 + RR) & 3]), g1 = rb_dot<PL>(X[1], Y[(1 + RR) & 3]),                         \
		       g2 = rb_dot<PL>(X[2], Y[(2 + RR) & 3]), g3 = rb_dot<PL>(X[3], Y[(3 + RR) & 3]);                         \
		rb_reduce4(g0, g1, g2, g3, lane);                                                                              \
		rb_rotate<PL>(X[0], Y[(0 + RR) & 3], g0, nrm, ix + 0, iy + ((0 + RR) & 3), tol2, worst, lane);                 \
		rb_rotate<PL>(X[1], Y[(1 + RR) & 3], g1, nrm, ix + 1, iy + ((1 + RR) & 3), tol2, worst, lane);                 \
		rb_rotate<PL>(X[2], Y[(2 + RR) & 3], g2, nrm, ix + 2, iy + ((2 + RR) & 3), tol2, worst, lane);                 \
		rb_rotate<PL>(X[3], Y[(3 + RR) & 3], g3, nrm, ix + 3, iy + ((3 + RR) & 3), tol2, worst, lane);                 \
		__syncwarp();                                                                                                  \
	}
				RB_CROSS(0)
				RB_CROSS(1)
				RB_CROSS(2)
				RB_CROSS(3)
#undef RB_CROSS
			}
			// tiles whose block is wanted elsewhere (or nowhere) in the next step go back to shared memory; everything
			// goes back at the end of a sweep (it may be the last one)
			const int s1 = (s + 1 == M) ? 0 : s + 1;
			const int nL = 2 * warp + (s1 & 1);
			const bool keep = worker && nL + 1 < N && step + 1 < N;
			const int nbL = keep ? rb_block_at(nL, s1, N) : -1, nbR = keep ? rb_block_at(nL + 1, s1, N) : -1;
			if (curX >= 0 && curX != nbL) {
				double* dst = R + (size_t)curX * kRB * ld + lane;
#pragma unroll
				for (int i = 0; i < kRB; ++i)
#pragma unroll
					for (int e = 0; e < PL; ++e) dst[i * ld + 32 * e] = X[i][e];
				curX = -1;
			}
			if (curY >= 0 && curY != nbR) {
				double* dst = R + (size_t)curY * kRB * ld + lane;
#pragma unroll
				for (int i = 0; i < kRB; ++i)
#pragma unroll
					for (int e = 0; e < PL; ++e) dst[i * ld + 32 * e] = Y[i][e];
				curY = -1;
			}
			__syncthreads();
			s = s1;
		}
		if (lane == 0) red[warp] = worst;
		__syncthreads();
		double wmax = 0.0;
		for (int i = 0; i < nw; ++i) wmax = fmax(wmax, red[i]);
		__syncthreads();
		// same stopping rule as the scalar sweeps: the largest scaled off-diagonal met was <= 3e-6, its rotations leave
		// ~1e-11 behind (quadratic convergence), far below the fp32 output rounding
		if (wmax <= 1e-11) { ++sweep; break; }
	}
	return sweep;
}

// Two warp-wide sums at once (7 fp64 shuffles).
__device__ __forceinline__ void rb_reduce2(double& g0, double& g1, const int lane) {
	const unsigned full = 0xffffffffu;
	const bool h16 = lane & 16;
	double k = h16 ? g1 : g0;
	k += __shfl_xor_sync(full, h16 ? g0 : g1, 16);
	k += __shfl_xor_sync(full, k, 8);
	k += __shfl_xor_sync(full, k, 4);
	k += __shfl_xor_sync(full, k, 2);
	k += __shfl_xor_sync(full, k, 1);
	g0 = __shfl_sync(full, k, 0);
	g1 = __shfl_sync(full, k, 16);
}

template <int PL>
__device__ __forceinline__ void rb_load4(double (&X)[kRB][PL], const double* __restrict__ src, const int ld) {
#pragma unroll
	for (int i = 0; i < kRB; ++i)
#pragma unroll
		for (int e = 0; e < PL; ++e) X[i][e] = src[i * ld + 32 * e];
}
template <int PL>
__device__ __forceinline__ void rb_store4(const double (&X)[kRB][PL], double* __restrict__ dst, const int ld) {
#pragma unroll
	for (int i = 0; i < kRB; ++i)
#pragma unroll
		for (int e = 0; e < PL; ++e) dst[i * ld + 32 * e] = X[i][e];
}

// fresh squared norms of the four rows of a tile, then its six inner pairs (three rounds of two disjoint pairs)
template <int PL>
__device__ __forceinline__ void rb_intra4(double (&X)[kRB][PL], double* __restrict__ nrm, const int ix, const double tol2,
                                          double& worst, const int lane) {
	double a0 = rb_dot<PL>(X[0], X[0]), a1 = rb_dot<PL>(X[1], X[1]), a2 = rb_dot<PL>(X[2], X[2]), a3 = rb_dot<PL>(X[3], X[3]);
	rb_reduce4(a0, a1, a2, a3, lane);
	if (lane == 0) { nrm[ix] = a0; nrm[ix + 1] = a1; nrm[ix + 2] = a2; nrm[ix + 3] = a3; }
	__syncwarp();
#define RB_INTRA2(P0, Q0, P1, Q1)                                                    \
	{                                                                                \
		double g0 = rb_dot<PL>(X[P0], X[Q0]), g1 = rb_dot<PL>(X[P1], X[Q1]);         \
		rb_reduce2(g0, g1, lane);                                                    \
		rb_rotate<PL>(X[P0], X[Q0], g0, nrm, ix + P0, ix + Q0, tol2, worst, lane);   \
		rb_rotate<PL>(X[P1], X[Q1], g1, nrm, ix + P1, ix + Q1, tol2, worst, lane);   \
		__syncwarp();                                                                \
	}
	RB_INTRA2(0, 1, 2, 3)
	RB_INTRA2(0, 2, 1, 3)
	RB_INTRA2(0, 3, 1, 2)
#undef RB_INTRA2
}

// Rows longer than 128 elements (PL = 5): eight rows of 160 doubles do not fit the 96 registers a 20-warp CTA leaves a
// thread, so only the left block of a pair is held in registers and the right block passes through them two rows at a
// time (rounds of two independent rotations); nothing stays in registers between steps - four block passes through
// shared memory per step instead of two, still ~9 times fewer per sweep than the scalar kernel.
template <int PL>
__device__ __forceinline__ int jacobi_sweeps_rb_half(double* __restrict__ R, const int n, const int ld, double* __restrict__ nrm,
                                                     double* __restrict__ red, const int max_sweeps, const double tol2) {
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
	const int N = rb_blocks(n), M = 2 * N;
	const bool worker = warp < (N >> 1);
	int s = 0, sweep = 0;
	for (; sweep < max_sweeps; ++sweep) {
		double worst = 0.0;
		for (int step = 0; step < N; ++step) {
			const int pL = 2 * warp + (s & 1);
			if (worker && pL + 1 < N) {
				const int bL = rb_block_at(pL, s, N), bR = rb_block_at(pL + 1, s, N);
				const int ix = bL * kRB, iy = bR * kRB;
				double X[kRB][PL];
				if (step == 0) {  // the pairs inside the right block first (it is only half resident below)
					rb_load4<PL>(X, R + (size_t)iy * ld + lane, ld);
					rb_intra4<PL>(X, nrm, iy, tol2, worst, lane);
					rb_store4<PL>(X, R + (size_t)iy * ld + lane, ld);
					__syncwarp();
				}
				rb_load4<PL>(X, R + (size_t)ix * ld + lane, ld);
				if (step == 0) rb_intra4<PL>(X, nrm, ix, tol2, worst, lane);
#pragma unroll 1
				for (int h = 0; h < 2; ++h) {
					double Y0[PL], Y1[PL];
					double* yr = R + (size_t)(iy + 2 * h) * ld + lane;
#pragma unroll
					for (int e = 0; e < PL; ++e) { Y0[e] = yr[32 * e]; Y1[e] = yr[ld + 32 * e]; }
					const int j0 = iy + 2 * h, j1 = j0 + 1;
#define RB_HALF(P0, P1, YA, JA, YB, JB)                                              \
	{                                                                                \
		double g0 = rb_dot<PL>(X[P0], YA), g1 = rb_dot<PL>(X[P1], YB);               \
		rb_reduce2(g0, g1, lane);                                                    \
		rb_rotate<PL>(X[P0], YA, g0, nrm, ix + P0, JA, tol2, worst, lane);           \
		rb_rotate<PL>(X[P1], YB, g1, nrm, ix + P1, JB, tol2, worst, lane);           \
		__syncwarp();                                                                \
	}
					RB_HALF(0, 1, Y0, j0, Y1, j1)
					RB_HALF(1, 0, Y0, j0, Y1, j1)
					RB_HALF(2, 3, Y0, j0, Y1, j1)
					RB_HALF(3, 2, Y0, j0, Y1, j1)
#undef RB_HALF
#pragma unroll
					for (int e = 0; e < PL; ++e) { yr[32 * e] = Y0[e]; yr[ld + 32 * e] = Y1[e]; }
				}
				rb_store4<PL>(X, R + (size_t)ix * ld + lane, ld);
			}
			__syncthreads();
			s = (s + 1 == M) ? 0 : s + 1;
		}
		if (lane == 0) red[warp] = worst;
		__syncthreads();
		double wmax = 0.0;
		for (int i = 0; i < nw; ++i) wmax = fmax(wmax, red[i]);
		__syncthreads();
		if (wmax <= 1e-11) { ++sweep; break; }
	}
	return sweep;
}
