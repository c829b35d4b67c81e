// Register-blocked one-sided Jacobi for the per-bin polar step (included by fh_polar.cu inside its anonymous namespace).
//
// Why: the scalar kernel of round 1 (a quarter-warp per row pair, rows streamed from shared memory) was shared-memory
// bound - every row crossed shared memory ~2.5 times per round-robin step, n - 1 steps per sweep (ncu: fp64 pipe 21 %,
// 1.25 eligible warps per scheduler). Here a WARP owns a pair of row blocks (2 x 4 rows) and keeps them in registers
// (lane l holds elements l, l + 32, ... of each row: 8 x PL doubles), performs all 4 x 4 cross rotations on them - four
// rounds of four independent rotations - and the blocks move between warps in the odd-even transposition ordering:
// N = n / 4 blocks on a line, even steps pair positions (2g, 2g + 1), odd steps (2g + 1, 2g + 2), the two blocks of a
// pair swap positions after their visit. Every pair of blocks meets exactly once in N steps, and a warp keeps ONE of its
// two blocks from a step to the next, so a step moves one block per warp through shared memory: per sweep each row
// crosses shared memory ~N times instead of ~2.5 (n - 1) times (n = 144: 36 vs 357). The pairs inside a block are
// rotated in the first step of a sweep.
//
// A round's scalar work is done ONCE per warp, not once per rotation: the four inner products are reduced together
// (two exchange stages + three butterfly stages leave gamma_j in the eight lanes of group j), each lane group evaluates
// the angle of ITS pair (first measured version: every lane evaluated all four angles in turn - ~120 instructions per
// rotation against ~25 of vector work, ncu: 15.9 M warp instructions per n = 144 problem), and (cos, sin, t gamma) of
// the four pairs are broadcast back with shuffles.
//
// The tracked squared norms of the rows live in REGISTERS, identical in every lane (same inputs, same operations);
// they go through the shared-memory table only together with their block, between two __syncthreads. (First version:
// lane 0 updated the table after every rotation and the other lanes re-read it after __syncwarp - measured wrong on the
// B200: orthogonality defects growing with n.)
//
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

template <int PL>
__device__ __forceinline__ double rb_dot(const double (&x)[PL], const double (&y)[PL]) {
	double a = 0.0, b = 0.0;
#pragma unroll
	for (int e = 0; e < PL; ++e) {
		if (e & 1) b = fma(x[e], y[e], b); else a = fma(x[e], y[e], a);
	}
	return a + b;
}

// Four warp-wide sums at once (fixed order): afterwards the eight lanes of group j = (lane >> 3) & 3 hold sum j.
__device__ __forceinline__ double rb_reduce4_grouped(const double g0, const double g1, const double g2, const double g3, const int lane) {
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
	return k;
}
// Two sums: lanes 0-15 end with sum 0, lanes 16-31 with sum 1.
__device__ __forceinline__ double rb_reduce2_grouped(const double g0, const double g1, const int lane) {
	const unsigned full = 0xffffffffu;
	const bool h16 = lane & 16;
	double k = h16 ? g1 : g0;
	k += __shfl_xor_sync(full, h16 ? g0 : g1, 16);
	k += __shfl_xor_sync(full, k, 8);
	k += __shfl_xor_sync(full, k, 4);
	k += __shfl_xor_sync(full, k, 2);
	k += __shfl_xor_sync(full, k, 1);
	return k;
}
__device__ __forceinline__ double rb_sel4(const int j, const double v0, const double v1, const double v2, const double v3) {
	return (j & 2) ? ((j & 1) ? v3 : v2) : ((j & 1) ? v1 : v0);
}

__device__ __forceinline__ float rb_rcpf(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rb_sqrtf(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float rb_rsqrtf(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// The rotation of this lane's pair: inner product ga, tracked squared norms al / be. Identity (cs 1, sn 0, dl 0) when
// the pair is left alone. dl = t gamma is the amount that moves from |x|^2 to |y|^2. `open` is set when the pair was
// still further from orthogonal than the stopping level.
// t = tan of the annihilating rotation = sgn 2|gamma| / (|d| + sqrt(d^2 + 4 gamma^2)), d = beta - alpha (the smaller root:
// |t| <= 1), evaluated in fp32 on operands scaled by one power of two (an fp32-accurate angle leaves a 1e-7 relative
// residual, removed quadratically by the next sweep); cos = (1 + t^2)^(-1/2) from an fp32 seed and two Newton steps in
// fp64, so that every rotation is orthogonal to fp64 rounding whatever t is. No IEEE division, no library call:
// ncu on the first version showed this scalar part, not the row updates, dominating the instruction count.
__device__ __forceinline__ void rb_angle(const double ga, const double al, const double be, const double tol2, bool& open, double& cs,
                                         double& sn, double& dl) {
	cs = 1.0; sn = 0.0; dl = 0.0;
	const double g2 = ga * ga, mn = al < be ? al : be;
	if (g2 > tol2 * mn * mn) {  // gamma = 0 never passes
		open = open || (g2 > 1e-11 * (al * be));
		const double d = be - al;
		const double ad = fabs(d), ag = 2.0 * fabs(ga);
		const double big = ad > ag ? ad : ag;
		const int ex = (__double2hiint(big) >> 20) & 0x7ff;
		const double scale = __hiloint2double((2046 - ex) << 20, 0);  // 2^(1023 - ex): the larger operand lands in [1, 2)
		const float ds = (float)(ad * scale), gs = (float)(ag * scale);
		float t = gs * rb_rcpf(ds + rb_sqrtf(fmaf(ds, ds, gs * gs)));
		t = t < 1.f ? t : 1.f;
		const double tt = ((d >= 0.0) == (ga >= 0.0)) ? (double)t : -(double)t;
		const double u = fma(tt, tt, 1.0);
		double y = (double)rb_rsqrtf((float)u);
		y = fma(0.5 * y, fma(-u * y, y, 1.0), y);
		y = fma(0.5 * y, fma(-u * y, y, 1.0), y);
		cs = y;
		sn = tt * y;
		dl = tt * ga;
	}
}

// x <- cs x - sn y, y <- sn x + cs y and the norm bookkeeping; (cs, sn, dl) are warp-uniform
template <int PL>
__device__ __forceinline__ void rb_apply(double (&x)[PL], double (&y)[PL], double& al, double& be, const double cs, const double sn,
                                         const double dl) {
	if (sn != 0.0) {
#pragma unroll
		for (int e = 0; e < PL; ++e) {
			const double a = x[e], c = y[e];
			x[e] = cs * a - sn * c;
			y[e] = sn * a + cs * c;
		}
		const double a2 = al - dl;
		al = a2 > 0.0 ? a2 : 0.0;
		be = be + dl;
	}
}

// One round of four disjoint pairs (xa, ya) .. (xd, yd) with tracked norms (na, ma) .. (nd, md).
#define RB_ROUND4(xa, ya, na, ma, xb, yb, nb_, mb, xc, yc, nc, mc, xd, yd, nd, md)                                      \
	{                                                                                                                  \
		const double gk = rb_reduce4_grouped(rb_dot<PL>(xa, ya), rb_dot<PL>(xb, yb), rb_dot<PL>(xc, yc), rb_dot<PL>(xd, yd), lane); \
		double cs_, sn_, dl_;                                                                                          \
		rb_angle(gk, rb_sel4(grp, na, nb_, nc, nd), rb_sel4(grp, ma, mb, mc, md), tol2, open, cs_, sn_, dl_);          \
		rb_apply<PL>(xa, ya, na, ma, __shfl_sync(0xffffffffu, cs_, 0), __shfl_sync(0xffffffffu, sn_, 0), __shfl_sync(0xffffffffu, dl_, 0));     \
		rb_apply<PL>(xb, yb, nb_, mb, __shfl_sync(0xffffffffu, cs_, 8), __shfl_sync(0xffffffffu, sn_, 8), __shfl_sync(0xffffffffu, dl_, 8));    \
		rb_apply<PL>(xc, yc, nc, mc, __shfl_sync(0xffffffffu, cs_, 16), __shfl_sync(0xffffffffu, sn_, 16), __shfl_sync(0xffffffffu, dl_, 16));  \
		rb_apply<PL>(xd, yd, nd, md, __shfl_sync(0xffffffffu, cs_, 24), __shfl_sync(0xffffffffu, sn_, 24), __shfl_sync(0xffffffffu, dl_, 24));  \
	}
// One round of two disjoint pairs.
#define RB_ROUND2(xa, ya, na, ma, xb, yb, nb_, mb)                                                                     \
	{                                                                                                                  \
		const double gk = rb_reduce2_grouped(rb_dot<PL>(xa, ya), rb_dot<PL>(xb, yb), lane);                            \
		double cs_, sn_, dl_;                                                                                          \
		rb_angle(gk, (lane & 16) ? nb_ : na, (lane & 16) ? mb : ma, tol2, open, cs_, sn_, dl_);                        \
		rb_apply<PL>(xa, ya, na, ma, __shfl_sync(0xffffffffu, cs_, 0), __shfl_sync(0xffffffffu, sn_, 0), __shfl_sync(0xffffffffu, dl_, 0));     \
		rb_apply<PL>(xb, yb, nb_, mb, __shfl_sync(0xffffffffu, cs_, 16), __shfl_sync(0xffffffffu, sn_, 16), __shfl_sync(0xffffffffu, dl_, 16)); \
	}

template <int PL>
__device__ __forceinline__ void rb_load4(double (&X)[kRB][PL], double (&nx)[kRB], const double* src, const double* nsrc, const int ld) {
#pragma unroll
	for (int i = 0; i < kRB; ++i) {
		nx[i] = nsrc[i];
#pragma unroll
		for (int e = 0; e < PL; ++e) X[i][e] = src[i * ld + 32 * e];
	}
}
template <int PL>
__device__ __forceinline__ void rb_store4(const double (&X)[kRB][PL], const double (&nx)[kRB], double* dst, double* ndst, const int ld,
                                          const int lane) {
#pragma unroll
	for (int i = 0; i < kRB; ++i) {
		if (lane == 0) ndst[i] = nx[i];
#pragma unroll
		for (int e = 0; e < PL; ++e) dst[i * ld + 32 * e] = X[i][e];
	}
}

// fresh squared norms of the four rows of a tile, in every lane
template <int PL>
__device__ __forceinline__ void rb_norms4(const double (&X)[kRB][PL], double (&nx)[kRB], const int lane) {
	const double k = rb_reduce4_grouped(rb_dot<PL>(X[0], X[0]), rb_dot<PL>(X[1], X[1]), rb_dot<PL>(X[2], X[2]), rb_dot<PL>(X[3], X[3]), lane);
	nx[0] = __shfl_sync(0xffffffffu, k, 0);
	nx[1] = __shfl_sync(0xffffffffu, k, 8);
	nx[2] = __shfl_sync(0xffffffffu, k, 16);
	nx[3] = __shfl_sync(0xffffffffu, k, 24);
}

// block-wide "some pair was still open" (every thread gets the answer)
__device__ __forceinline__ bool rb_any_open(const bool open, double* red, const int lane, const int warp, const int nw) {
	const bool w = __any_sync(0xffffffffu, open);
	if (lane == 0) red[warp] = w ? 1.0 : 0.0;
	__syncthreads();
	bool any = false;
	for (int i = 0; i < nw; ++i) any = any || (red[i] != 0.0);
	__syncthreads();
	return any;
}

// Sweeps over the rows of R (rb_rows(n) x ld, pad rows / columns zero). Warps [0, N / 2) work, the others only meet
// the barriers. Returns the sweep count.
template <int PL>
__device__ __forceinline__ int jacobi_sweeps_rb(double* R, const int n, const int ld, double* nrm, double* red, const int max_sweeps,
                                                const double tol2) {
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
	const int grp = (lane >> 3) & 3;
	const int N = rb_blocks(n), M = 2 * N;
	const bool worker = warp < (N >> 1);
	double X[kRB][PL], Y[kRB][PL], nx[kRB], ny[kRB];
	int curX = -1, curY = -1;  // block held by each register tile
	int s = 0;                 // global step modulo 2N (N even: a sweep starts on an even step)
	int sweep = 0;
	for (; sweep < max_sweeps; ++sweep) {
		bool open = false;  // a pair further from orthogonal than the stopping level was met in this sweep
		for (int step = 0; step < N; ++step) {
			const int pL = 2 * warp + (s & 1);
			if (worker && pL + 1 < N) {
				const int bL = rb_block_at(pL, s, N), bR = rb_block_at(pL + 1, s, N);
				if (curX != bL) { rb_load4<PL>(X, nx, R + (size_t)bL * kRB * ld + lane, nrm + bL * kRB, ld); curX = bL; }
				if (curY != bR) { rb_load4<PL>(Y, ny, R + (size_t)bR * kRB * ld + lane, nrm + bR * kRB, ld); curY = bR; }
				if (step == 0) {
					// fresh squared norms of the eight rows, then the six pairs inside each block (three rounds of two
					// disjoint pairs per block)
					rb_norms4<PL>(X, nx, lane);
					rb_norms4<PL>(Y, ny, lane);
					RB_ROUND4(X[0], X[1], nx[0], nx[1], X[2], X[3], nx[2], nx[3], Y[0], Y[1], ny[0], ny[1], Y[2], Y[3], ny[2], ny[3])
					RB_ROUND4(X[0], X[2], nx[0], nx[2], X[1], X[3], nx[1], nx[3], Y[0], Y[2], ny[0], ny[2], Y[1], Y[3], ny[1], ny[3])
					RB_ROUND4(X[0], X[3], nx[0], nx[3], X[1], X[2], nx[1], nx[2], Y[0], Y[3], ny[0], ny[3], Y[1], Y[2], ny[1], ny[2])
				}
				// the 16 cross pairs: round r rotates (X[i], Y[(i + r) & 3]), i = 0..3 - disjoint rows
				RB_ROUND4(X[0], Y[0], nx[0], ny[0], X[1], Y[1], nx[1], ny[1], X[2], Y[2], nx[2], ny[2], X[3], Y[3], nx[3], ny[3])
				RB_ROUND4(X[0], Y[1], nx[0], ny[1], X[1], Y[2], nx[1], ny[2], X[2], Y[3], nx[2], ny[3], X[3], Y[0], nx[3], ny[0])
				RB_ROUND4(X[0], Y[2], nx[0], ny[2], X[1], Y[3], nx[1], ny[3], X[2], Y[0], nx[2], ny[0], X[3], Y[1], nx[3], ny[1])
				RB_ROUND4(X[0], Y[3], nx[0], ny[3], X[1], Y[0], nx[1], ny[0], X[2], Y[1], nx[2], ny[1], X[3], Y[2], nx[3], ny[2])
			}
			// tiles whose block is wanted elsewhere (or nowhere) in the next step go back to shared memory; everything
			// goes back at the end of a sweep (it may be the last one)
			const int s1 = (s + 1 == M) ? 0 : s + 1;
			const int nL = 2 * warp + (s1 & 1);
			const bool keep = worker && nL + 1 < N && step + 1 < N;
			const int nbL = keep ? rb_block_at(nL, s1, N) : -1, nbR = keep ? rb_block_at(nL + 1, s1, N) : -1;
			if (curX >= 0 && curX != nbL) { rb_store4<PL>(X, nx, R + (size_t)curX * kRB * ld + lane, nrm + curX * kRB, ld, lane); curX = -1; }
			if (curY >= 0 && curY != nbR) { rb_store4<PL>(Y, ny, R + (size_t)curY * kRB * ld + lane, nrm + curY * kRB, ld, lane); curY = -1; }
			__syncthreads();
			s = s1;
		}
		// same stopping rule as the scalar sweeps: the largest scaled off-diagonal met was <= 3e-6 (cos^2 <= 1e-11), its
		// rotations leave ~1e-11 behind (quadratic convergence), far below the fp32 output rounding
		if (!rb_any_open(open, red, lane, warp, nw)) { ++sweep; break; }
	}
	return sweep;
}

// Rows longer than 128 elements (PL = 5): eight rows of 160 doubles do not fit the 96 registers a 20-warp CTA leaves a
// thread, so only the left block of a pair is held in registers and the right block passes through them two rows at a
// time (rounds of two independent rotations); nothing stays in registers between steps - four block passes through
// shared memory per step instead of two, still ~9 times fewer per sweep than the scalar kernel.
template <int PL>
__device__ __forceinline__ int jacobi_sweeps_rb_half(double* R, const int n, const int ld, double* nrm, double* red, const int max_sweeps,
                                                     const double tol2) {
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
	const int N = rb_blocks(n), M = 2 * N;
	const bool worker = warp < (N >> 1);
	int s = 0, sweep = 0;
	for (; sweep < max_sweeps; ++sweep) {
		bool open = false;
		for (int step = 0; step < N; ++step) {
			const int pL = 2 * warp + (s & 1);
			if (worker && pL + 1 < N) {
				const int bL = rb_block_at(pL, s, N), bR = rb_block_at(pL + 1, s, N);
				const int ix = bL * kRB, iy = bR * kRB;
				double X[kRB][PL], nx[kRB];
				if (step == 0) {  // the pairs inside the right block first (it is only half resident below)
					rb_load4<PL>(X, nx, R + (size_t)iy * ld + lane, nrm + iy, ld);
					rb_norms4<PL>(X, nx, lane);
					RB_ROUND2(X[0], X[1], nx[0], nx[1], X[2], X[3], nx[2], nx[3])
					RB_ROUND2(X[0], X[2], nx[0], nx[2], X[1], X[3], nx[1], nx[3])
					RB_ROUND2(X[0], X[3], nx[0], nx[3], X[1], X[2], nx[1], nx[2])
					rb_store4<PL>(X, nx, R + (size_t)iy * ld + lane, nrm + iy, ld, lane);
					__syncwarp();
				}
				rb_load4<PL>(X, nx, R + (size_t)ix * ld + lane, nrm + ix, ld);
				if (step == 0) {
					rb_norms4<PL>(X, nx, lane);
					RB_ROUND2(X[0], X[1], nx[0], nx[1], X[2], X[3], nx[2], nx[3])
					RB_ROUND2(X[0], X[2], nx[0], nx[2], X[1], X[3], nx[1], nx[3])
					RB_ROUND2(X[0], X[3], nx[0], nx[3], X[1], X[2], nx[1], nx[2])
				}
#pragma unroll 1
				for (int h = 0; h < 2; ++h) {
					double Y0[PL], Y1[PL];
					double* yr = R + (size_t)(iy + 2 * h) * ld + lane;
					volatile double* yn = nrm + iy + 2 * h;  // written by lane 0 a few lines up in the first step: re-read, never cached
					double n0 = yn[0], n1 = yn[1];
#pragma unroll
					for (int e = 0; e < PL; ++e) { Y0[e] = yr[32 * e]; Y1[e] = yr[ld + 32 * e]; }
					RB_ROUND2(X[0], Y0, nx[0], n0, X[1], Y1, nx[1], n1)
					RB_ROUND2(X[1], Y0, nx[1], n0, X[0], Y1, nx[0], n1)
					RB_ROUND2(X[2], Y0, nx[2], n0, X[3], Y1, nx[3], n1)
					RB_ROUND2(X[3], Y0, nx[3], n0, X[2], Y1, nx[2], n1)
					if (lane == 0) { yn[0] = n0; yn[1] = n1; }
#pragma unroll
					for (int e = 0; e < PL; ++e) { yr[32 * e] = Y0[e]; yr[ld + 32 * e] = Y1[e]; }
				}
				rb_store4<PL>(X, nx, R + (size_t)ix * ld + lane, nrm + ix, ld, lane);
			}
			__syncthreads();
			s = (s + 1 == M) ? 0 : s + 1;
		}
		if (!rb_any_open(open, red, lane, warp, nw)) { ++sweep; break; }
	}
	return sweep;
}
#undef RB_ROUND4
#undef RB_ROUND2
