// Block one-sided Jacobi for the per-bin polar step (alternative inner loop of chol_jacobi_kernel,
// fh_polar.cu; opt-in with FH_POLAR_BLOCK=1 until it has been timed on a B200).
//
// Why: the scalar kernel moves every row of R through shared memory ~2.5 times per round-robin step,
// n-1 steps per sweep (ncu: 24 ms for 938 problems of n ~ 140, i.e. ~3.3 us per step of which ~2 us is
// the 525 KB of shared-memory traffic). Here rows are grouped in blocks of 8; a step pairs blocks, and
// for every pair (16 rows X):  S = X X^T (16 x 16)  ->  S = Z Lambda Z^T by a two-sided cyclic Jacobi on
// the small matrix  ->  X <- Z^T X.  One sweep has ceil(n/8)-1 steps instead of n-1, each moving the rows
// three times: ~7x less shared-memory traffic per sweep. Numerics checked on the CPU (numpy prototype and
// the host emulation of THIS file, tests/test_polar_block_emulation.py): same outer sweep count (7-8) and
// the same polar-factor error as the scalar kernel on graded spectra up to kappa 3e6 with ONE cyclic sweep over
// the 16 x 16 problem per visit (iterating it to convergence, 2.7 inner sweeps on average, saves no outer sweep).
//
// The file is written in barrier-separated PHASES whose bodies depend only on the thread / lane index and
// on shared memory (no warp shuffles, no per-lane state across a barrier). With FH_EMU defined the phase
// macros become plain loops over threads, so the same source runs on the host (tests/emu/) - in forward and
// reverse thread order, which must give bit-identical results: any dependence between two threads inside
// one phase (a missing barrier) shows up as a difference.
#pragma once

#ifdef FH_EMU
#include <cmath>
struct fh_d2 { double x, y; };
extern int fh_emu_reverse;
#define FH_DEV inline
#define FH_FOR_THREADS(v, cnt) for (int _i##v = 0, v = 0; _i##v < (cnt) && ((v = fh_emu_reverse ? (cnt) - 1 - _i##v : _i##v), true); ++_i##v)
#define FH_FOR_WARPS(v, cnt) FH_FOR_THREADS(v, cnt)
#define FH_FOR_LANES(v) FH_FOR_THREADS(v, 32)
#define FH_WARP_SYNC() ((void)0)
#define FH_CTA_SYNC() ((void)0)
#define FH_LD2(p) (*(const fh_d2*)(p))
#else
typedef double2 fh_d2;
#define FH_DEV __device__ __forceinline__
#define FH_FOR_THREADS(v, cnt) for (int v = threadIdx.x, _o##v = 1; _o##v; _o##v = 0)
#define FH_FOR_WARPS(v, cnt) for (int v = threadIdx.x >> 5, _o##v = 1; _o##v; _o##v = 0)
#define FH_FOR_LANES(v) for (int v = threadIdx.x & 31, _o##v = 1; _o##v; _o##v = 0)
#define FH_WARP_SYNC() __syncwarp()
#define FH_CTA_SYNC() __syncthreads()
#define FH_LD2(p) (*reinterpret_cast<const double2*>(p))
#endif

constexpr int kBJRows = 8;            // rows per block
constexpr int kBJSlot = 536;          // doubles of scratch per block pair: S 256 | Z 256 | cs 16 | worst 1 | ctl 7 (14 ints)
#ifdef FH_EMU
// the emulation can vary the inner-solve policy and counts phases (tests/emu, design studies)
extern int fh_emu_cross_only;
extern long long fh_emu_count[4];      // [0] block-pair visits, [1] J1 phases, [2] J2+J3 phase pairs
#define kBJCrossOnly fh_emu_cross_only
#define FH_EMU_COUNT(i) (++fh_emu_count[i])
#else
// 1: after the first step of a sweep only the 64 cross pairs (row of block P, row of block Q) are rotated - the
// within-block pairs have had their visit of this sweep at step 0, exactly as in a cyclic scalar sweep.
constexpr int kBJCrossOnly = 1;
#define FH_EMU_COUNT(i) ((void)0)
#endif
constexpr double kBJInnerTol = 1e-22;   // rotation threshold on s_pq^2 / (s_pp s_qq)
constexpr int kBJMaxSide = 152;       // largest Gram side whose rows + scratch fit 227 KB of shared memory

struct BJSlot {
	double *S, *Z, *cs, *worst;
	int* ctl;  // [2] pair active, [3] valid rows (8 | 16), [4],[5] block ids
};
FH_DEV BJSlot bj_slot(double* scratch, int t) {
	double* b = scratch + (size_t)t * kBJSlot;
	BJSlot s;
	s.S = b; s.Z = b + 256; s.cs = b + 512; s.worst = b + 528; s.ctl = (int*)(b + 529);
	return s;
}
inline __host__ __device__ int bj_blocks(int n) { return (n + kBJRows - 1) / kBJRows; }
inline __host__ __device__ int bj_slots(int n) { return (bj_blocks(n) + 1) >> 1; }
inline __host__ __device__ int bj_rows(int n) { return bj_blocks(n) * kBJRows; }     // rows of R incl. zero padding
inline __host__ __device__ int bj_ld(int n) { return (n + 1) & ~1; }                 // even: 16-byte row pairs
inline __host__ __device__ size_t bj_scratch_doubles(int n) { return (size_t)bj_slots(n) * kBJSlot + 2; }

// tan of the rotation angle for alpha = s_pp, beta = s_qq, gamma = s_pq. Device: the scalar kernel's jacobi_tan (fp32 angle
// with approximate div / sqrt - the residual it leaves is removed quadratically by the next visit); the host emulation
// evaluates the same formula exactly. cos = rsqrt(1 + t^2) in fp64 keeps every rotation orthogonal either way.
#ifdef FH_EMU
extern int fh_emu_fp32_angle;  // 1: round the angle to fp32 as the device's approximate evaluation does
inline double bj_tan(double al, double be, double ga) {
	const double zeta = (be - al) / (2.0 * ga);
	const double t = (zeta == 0.0) ? 1.0 : std::copysign(1.0, zeta) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
	return fh_emu_fp32_angle ? (double)((float)t * (1.0f + 2e-7f)) : t;
}
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
#else
#define bj_tan jacobi_tan
#endif

// pair `t` of round-robin step `s` among m = mm + 1 players (player mm fixed): p < q
FH_DEV void bj_pair(int t, int s, int mm, int& p, int& q) {
	if (t == 0) { p = mm; q = s; }
	else {
		p = s + t; if (p >= mm) p -= mm;
		q = s - t + mm; if (q >= mm) q -= mm;
	}
	if (p > q) { int x = p; p = q; q = x; }
}

// rotation `t` (0..7) of inner step `s`: the round-robin over all 16 rows (15 steps), or the cross pairs only (8 steps)
FH_DEV void bj_inner_pair(int t, int s, bool cross, int& p, int& q) {
	if (cross) { p = t; q = 8 + ((t + s) & 7); }
	else bj_pair(t, s, 15, p, q);
}

// Orthogonalises the rows of R (n x ld, rows n..bj_rows(n) and the pad columns zero) in place.
// scratch: bj_scratch_doubles(n) doubles, 16-byte aligned. Returns the number of sweeps.
FH_DEV int block_jacobi_sweeps(double* R, const int n, const int ld, double* scratch, const int nthreads,
                               const int max_sweeps, const double skip_tol) {
	const int nw = nthreads >> 5;
	const int nblk = bj_blocks(n);
	const int m = (nblk + 1) & ~1, mm = m - 1, nslot = m >> 1;
	double* sweep_worst = scratch + (size_t)nslot * kBJSlot;  // [2], ping-pong by sweep parity
	const int nchunk = (ld + 63) >> 6;  // 64-column chunks of the apply phase
	FH_FOR_THREADS(tid, nthreads) {
		if (tid < 2) sweep_worst[tid] = 0.0;
	}
	FH_CTA_SYNC();
	int sweep = 0;
	for (; sweep < max_sweeps; ++sweep) {
		double* sw_cur = sweep_worst + (sweep & 1);
		for (int step = 0; step < mm; ++step) {
			// ---------------- P1: Gram of every block pair of this step, two work items per pair over all warps ----------------
			// item 2t    : rows of block P against the rows of P and Q (8 x 16 entries, a 2 x 2 tile per lane)
			// item 2t + 1: rows of block Q against the rows of Q      (8 x 8 entries, 1 x 2 per lane)
			// The Q x P block is the transpose of P x Q: the pair's warp mirrors it at the start of P2 (3/4 of the fp64 work of
			// the full 16 x 16 product, spread over all warps instead of one warp per pair).
			FH_FOR_WARPS(w, nw) {
				for (int item = w; item < 2 * nslot; item += nw) {
					const int t = item >> 1, half = item & 1;
					BJSlot sl = bj_slot(scratch, t);
					int P, Q;
					bj_pair(t, step, mm, P, Q);
					const bool hasQ = Q < nblk;  // false: the dummy player of an odd block count, P is alone
					FH_FOR_LANES(lane) {
						if (half == 0) {
							const int a = lane >> 3, l = 2 * (lane & 7);
							const double* pi = R + (size_t)(P * 8 + 2 * a) * ld;
							const bool jv = l < 8 || hasQ;
							const double* pj = R + (size_t)((l < 8) ? P * 8 + l : (hasQ ? Q * 8 + (l - 8) : 0)) * ld;
							double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
							if (jv) {
								for (int c = 0; c < ld; c += 2) {
									const fh_d2 x0 = FH_LD2(pi + c), x1 = FH_LD2(pi + ld + c), y0 = FH_LD2(pj + c), y1 = FH_LD2(pj + ld + c);
									a00 = fma(x0.x, y0.x, a00); a00 = fma(x0.y, y0.y, a00);
									a01 = fma(x0.x, y1.x, a01); a01 = fma(x0.y, y1.y, a01);
									a10 = fma(x1.x, y0.x, a10); a10 = fma(x1.y, y0.y, a10);
									a11 = fma(x1.x, y1.x, a11); a11 = fma(x1.y, y1.y, a11);
								}
							}
							sl.S[(2 * a) * 16 + l] = a00; sl.S[(2 * a) * 16 + l + 1] = a01;
							sl.S[(2 * a + 1) * 16 + l] = a10; sl.S[(2 * a + 1) * 16 + l + 1] = a11;
						} else {
							const int i = lane >> 2, l = 2 * (lane & 3);
							double a0 = 0.0, a1 = 0.0;
							if (hasQ) {
								const double* pi = R + (size_t)(Q * 8 + i) * ld;
								const double* pj = R + (size_t)(Q * 8 + l) * ld;
								for (int c = 0; c < ld; c += 2) {
									const fh_d2 x = FH_LD2(pi + c), y0 = FH_LD2(pj + c), y1 = FH_LD2(pj + ld + c);
									a0 = fma(x.x, y0.x, a0); a0 = fma(x.y, y0.y, a0);
									a1 = fma(x.x, y1.x, a1); a1 = fma(x.y, y1.y, a1);
								}
							}
							sl.S[(8 + i) * 16 + 8 + l] = a0; sl.S[(8 + i) * 16 + 8 + l + 1] = a1;
						}
					}
				}
			}
			FH_CTA_SYNC();
			// ---------------- P2: per pair: activity flags, then the rotations of S (two-sided cyclic Jacobi), one warp ----------------
			FH_FOR_WARPS(w, nw) {
				for (int t = w; t < nslot; t += nw) {
					BJSlot sl = bj_slot(scratch, t);
					int P, Q;
					bj_pair(t, step, mm, P, Q);
					const int nv = (Q < nblk) ? 16 : 8;
					const bool cross = kBJCrossOnly && step != 0;
					FH_FOR_LANES(lane) {  // Q x P block = (P x Q block)^T
#pragma unroll
						for (int e = 0; e < 2; ++e) {
							const int idx = lane * 2 + e, i = 8 + (idx >> 3), j = idx & 7;
							sl.S[i * 16 + j] = sl.S[j * 16 + i];
						}
					}
					FH_WARP_SYNC();
					// does any entry still need a rotation (scalar kernel's rule: s_ij^2 > skip * min(d_i, d_j)^2), and is any of those
					// above the stopping level s_ij^2 > 1e-11 d_i d_j? Two flags per row, no divisions.
					FH_FOR_LANES(lane) {
						if (lane < 16) {
							const double di = sl.S[lane * 16 + lane];
							int flags = 0;
							for (int j = 0; j < 16; ++j) {
								if (j == lane || (cross && ((j < 8) == (lane < 8)))) continue;
								const double sij = sl.S[lane * 16 + j], dj = sl.S[j * 16 + j];
								const double mn = fmin(di, dj), g2 = sij * sij;
								if (g2 > skip_tol * mn * mn) flags |= (g2 > 1e-11 * (di * dj)) ? 3 : 1;
							}
							sl.cs[lane] = (double)flags;
						}
					}
					FH_WARP_SYNC();
					FH_FOR_LANES(lane) {
						if (lane == 0) {
							int flags = 0;
							for (int i = 0; i < 16; ++i) flags |= (int)sl.cs[i];
							sl.worst[0] = (flags & 2) ? 1.0 : 0.0;  // 1: this visit met an off-diagonal above the stopping level
							sl.ctl[2] = flags & 1;
							sl.ctl[3] = nv; sl.ctl[4] = P; sl.ctl[5] = Q;
						}
					}
					FH_WARP_SYNC();
					if (sl.ctl[2] == 0) continue;
					FH_EMU_COUNT(0);
					FH_FOR_LANES(lane) {
						for (int e = lane; e < 256; e += 32) sl.Z[e] = ((e >> 4) == (e & 15)) ? 1.0 : 0.0;
					}
					FH_WARP_SYNC();
					// ONE cyclic sweep over the pair's rotations per visit (measured on the CPU: iterating the 16 x 16 problem to
					// convergence triples the phases per visit and does not save a single outer sweep)
					const int nsteps = cross ? 8 : 15;
					for (int s = 0; s < nsteps; ++s) {
						// J1: the 8 disjoint rotations of this step
						FH_FOR_LANES(lane) {
							if (lane < 8) {
								int p, q;
								bj_inner_pair(lane, s, cross, p, q);
								const double app = sl.S[p * 16 + p], aqq = sl.S[q * 16 + q], apq = sl.S[p * 16 + q];
								double c = 1.0, sn = 0.0;
								if (apq != 0.0 && apq * apq > kBJInnerTol * fabs(app * aqq)) {
									const double tt = bj_tan(app, aqq, apq);
									c = rsqrt(tt * tt + 1.0);
									sn = tt * c;
								}
								sl.cs[2 * lane] = c;
								sl.cs[2 * lane + 1] = sn;
							}
						}
						FH_WARP_SYNC();
						FH_EMU_COUNT(1);
						// a step without any rotation: nothing to apply. cs is rewritten by the next J1, so its readers must be a
						// barrier ahead of it.
						bool any = false;
						for (int e = 0; e < 8; ++e) any = any || (sl.cs[2 * e + 1] != 0.0);
						if (!any) { FH_WARP_SYNC(); continue; }
						FH_EMU_COUNT(2);
						// J2: S <- J^T S J and Z <- Z J in one phase. The 8 disjoint pairs tile S into 8 x 8 blocks of 2 x 2 entries
						// ({p_a, q_a} x {p_b, q_b}); a block only needs its own entries and the two rotations: 2 blocks per lane.
						FH_FOR_LANES(lane) {
#pragma unroll
							for (int e = 0; e < 2; ++e) {
								const int bidx = lane * 2 + e, a = bidx >> 3, b = bidx & 7;
								const double ca = sl.cs[2 * a], sa = sl.cs[2 * a + 1], cb = sl.cs[2 * b], sb = sl.cs[2 * b + 1];
								if (sa == 0.0 && sb == 0.0) continue;
								int pa, qa, pb, qb;
								bj_inner_pair(a, s, cross, pa, qa);
								bj_inner_pair(b, s, cross, pb, qb);
								const double bpp = sl.S[pa * 16 + pb], bpq = sl.S[pa * 16 + qb], bqp = sl.S[qa * 16 + pb], bqq = sl.S[qa * 16 + qb];
								const double tpp = ca * bpp - sa * bqp, tpq = ca * bpq - sa * bqq;  // J_a^T B
								const double tqp = sa * bpp + ca * bqp, tqq = sa * bpq + ca * bqq;
								sl.S[pa * 16 + pb] = tpp * cb - tpq * sb;                             // ... J_b
								sl.S[pa * 16 + qb] = tpp * sb + tpq * cb;
								sl.S[qa * 16 + pb] = tqp * cb - tqq * sb;
								sl.S[qa * 16 + qb] = tqp * sb + tqq * cb;
							}
#pragma unroll
							for (int e = 0; e < 4; ++e) {
								const int combo = lane * 4 + e, i = combo >> 3, tt = combo & 7;
								const double c = sl.cs[2 * tt], sn = sl.cs[2 * tt + 1];
								if (sn == 0.0) continue;
								int p, q;
								bj_inner_pair(tt, s, cross, p, q);
								const double zp = sl.Z[i * 16 + p], zq = sl.Z[i * 16 + q];
								sl.Z[i * 16 + p] = c * zp - sn * zq;
								sl.Z[i * 16 + q] = sn * zp + c * zq;
							}
						}
						FH_WARP_SYNC();
					}
				}
			}
			FH_CTA_SYNC();
			// ---------------- P3: X <- Z^T X for every active pair, all warps, two columns (c, c + 32) per lane ----------------
			// (the Z entries are warp-uniform broadcast loads: one load feeds the FMAs of both columns, which halves the
			// shared-memory instructions per fp64 FMA)
			FH_FOR_THREADS(tid, nthreads) {
				const int w = tid >> 5, lane = tid & 31;
				for (int item = w; item < nslot * nchunk; item += nw) {
					const int t = item / nchunk, c0 = (item - t * nchunk) * 64 + lane, c1 = c0 + 32;
					BJSlot sl = bj_slot(scratch, t);
					if (sl.ctl[2] == 0 || c0 >= ld) continue;
					const bool two = c1 < ld;
					const int nv = sl.ctl[3], P = sl.ctl[4], Q = sl.ctl[5];
					double x0[16], x1[16];
#pragma unroll
					for (int k = 0; k < 16; ++k) {
						const double* src = R + (size_t)((k < 8) ? P * 8 + k : Q * 8 + (k - 8)) * ld;
						x0[k] = (k < nv) ? src[c0] : 0.0;
						x1[k] = (k < nv && two) ? src[c1] : 0.0;
					}
#pragma unroll
					for (int h = 0; h < 4; ++h) {
						double o0 = 0.0, o1 = 0.0, o2 = 0.0, o3 = 0.0, q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
#pragma unroll
						for (int k = 0; k < 16; ++k) {
							const fh_d2 za = FH_LD2(sl.Z + k * 16 + 4 * h), zb = FH_LD2(sl.Z + k * 16 + 4 * h + 2);
							o0 = fma(za.x, x0[k], o0); o1 = fma(za.y, x0[k], o1);
							o2 = fma(zb.x, x0[k], o2); o3 = fma(zb.y, x0[k], o3);
							q0 = fma(za.x, x1[k], q0); q1 = fma(za.y, x1[k], q1);
							q2 = fma(zb.x, x1[k], q2); q3 = fma(zb.y, x1[k], q3);
						}
						const int l = 4 * h;
						if (l < nv) {
							double* dst = R + (size_t)((l < 8) ? P * 8 + l : Q * 8 + (l - 8)) * ld;
							dst[c0] = o0; dst[ld + c0] = o1; dst[2 * (size_t)ld + c0] = o2; dst[3 * (size_t)ld + c0] = o3;
							if (two) { dst[c1] = q0; dst[ld + c1] = q1; dst[2 * (size_t)ld + c1] = q2; dst[3 * (size_t)ld + c1] = q3; }
						}
					}
				}
				if (tid == 0) {
					double wm = *sw_cur;
					for (int t = 0; t < nslot; ++t) wm = fmax(wm, bj_slot(scratch, t).worst[0]);
					*sw_cur = wm;
					if (step == 0) sweep_worst[(sweep + 1) & 1] = 0.0;  // its last reader is two barriers behind
				}
			}
			FH_CTA_SYNC();
		}
		// no rotated pair of the whole sweep was above the stopping level (CTA-uniform read): same rule as the scalar kernel
		if (*sw_cur < 0.5) { ++sweep; break; }
	}
	return sweep;
}
