// Fused RWR imputation of one bin block from the conv'd panel A (reference: partial_rwr.py:80-126, :138):
//   S2 = A A^T;  P = colnorm(3/4 colnorm(A_diag_block) + 1/4 colnorm(S2 - diag))     [PANEL mode only]
//   Q_1 = 0.5 P + 0.5 I,   Q_{t+1} = 0.5 Q_t P + 0.5 I  (t = 1 .. k-1),   X = Q_k A
// for every cell of the chunk in ONE persistent tcgen05 kernel. The per-launch GEMM chain it replaces
// (k-1 launches of Q P plus Q A) is HBM-bound: each 128 x 128 x 128 step re-reads Q and P and re-writes Q
// (53 KB each per cell, measured 4.2 TB/s of traffic for ~90 TF/s of useful maths). Here Q never
// leaves the SM:
//   * Q (hi / lo TF32 halves) lives in TENSOR MEMORY and is the A operand of every MMA
//     (tcgen05.mma with A from TMEM); a step's accumulator is read back (tcgen05.ld), turned into
//     0.5 d + 0.5 I, split and written straight back (tcgen05.st) by the drain warps - no shared
//     memory round trip.
//   * P (the B operand of the chain) is loaded once per cell by TMA into 4 slots of a 6-slot
//     shared-memory ring and stays there for all steps; the remaining slots already prefetch the
//     first A tiles. During X = Q A the whole ring streams A (MN-major B tiles), so the next cell's P
//     arrives while this cell's product is still running.
//   * HBM traffic per cell: P + A in, X out (343 KB at nb=115, w=316) instead of ~1.1 MB.
//   * PANEL mode also forms S2 on the tensor cores (each landed K-major A tile is BOTH operands) and runs
//     the transition-matrix normalisation in the drain warps: thread = row; the first-order column sums
//     (through the warps' private staging tiles, fixed order: deterministic) are taken while the S2 MMAs run,
//     the symmetric S2 column sums are row sums, the column sum of the blend is analytic. P goes
//     to a per-CTA 64 KB global scratch (L2 resident) and comes back through TMA as the chain's B operand;
//     Q_1 is written straight into TMEM. HBM traffic per cell: A twice in (second read mostly L2), X out.
// Arithmetic is the same 3xTF32 split, MMA order and fp32 epilogue as fh_gemm_tc.cu (K = nb <= 128 is a
// single accumulation chunk), so results match the unfused path.
// TMEM (512 columns): Q_hi [0,128) | Q_lo [128,256) | accumulator 0 [256,384) | accumulator 1 [384,512)
// Roles: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 hi/lo splitters,
// warps 8-15 drain (accumulator -> Q or -> X).
#include "fh_tc.cuh"
#include "../../include/fh_b200.h"

namespace {
using namespace fh_tc;

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int TILE_BYTES = BK * BN * 4;         // 16 KB: 32 k-rows x 128 n (four 32-wide TMA boxes) or 128 rows x 32 k
constexpr int SLOT_BYTES = 2 * TILE_BYTES;      // hi (raw fp32 as landed), lo
constexpr int EPI_BYTES = 8 * 32 * 32 * 4;      // 8 drain warps x (32 x 32 floats, XOR-swizzled)
constexpr int XCH_FLOATS = 4 * 128 + 2 * 128 + 4 * 128;  // PANEL: column partials [4][128], row sums [2][128], cs1/w1/w2/flag [128]
constexpr int NTHREADS = 512;
#ifndef FH_CHAIN_CHUNK_KB
#define FH_CHAIN_CHUNK_KB 4                     // compile-time knob for A/B builds (FH_NVCC_EXTRA, scripts/gpu_session.sh)
#endif
constexpr int CHUNK_KB = FH_CHAIN_CHUNK_KB;     // S2: k-blocks accumulated in TMEM before a drain (as fh_gemm_tc.cu)
constexpr float EPS = 1e-15f;                   // partial_rwr.py:88-97
template <bool PANEL> struct Cfg {
	static constexpr int SLOTS = PANEL ? 5 : 6;
	static constexpr int SMEM_BYTES = SLOTS * SLOT_BYTES + EPI_BYTES + (PANEL ? XCH_FLOATS * 4 : 0) + 1024 /*align*/ + 256 /*barriers*/;
};
constexpr uint32_t TM_QHI = 0, TM_QLO = 128, TM_ACC = 256;

struct ChainP {
	int nb, w, ldw, ldp, k, ncell;
	long long p_cell_stride, out_cell_stride;
	const float* P;        // !PANEL: transition matrices (Q_1 source). PANEL: per-CTA scratch (gridDim x 128 x 128)
	const float* A;        // PANEL: the panel, for the first-order block A[:, s:s+nb]
	long long a_cell_stride;
	int s;
	float* out;
	int vec_ok;
	int tma_out;           // X tiles leave through TMA stores (tmO valid: 16-byte aligned rows)
	int dbg;               // FH_CHAIN_DEBUG bits (experiments): 1 no L2 prefetch, 2 no TMA stores of X
	long long* trace;      // FH_CHAIN_TRACE=1: clock64 stamps of CTA 0's 4th cell (debug)
};
#define FH_TRACE(slot)                                                         \
	do {                                                                       \
		if (p.trace && blockIdx.x == 0 && cell == 3 * (int)gridDim.x) p.trace[slot] = clock64(); \
	} while (0)

// tmP: P as (k rows x n) MN-major boxes (PANEL: the scratch, one "cell" per CTA); tmA: the panel as MN-major
// B tiles of Q A; tmK (PANEL): the panel as K-major 128 x 32 tiles for S2; tmO: X as 32 x 32 store boxes;
// tmS (PANEL): the scratch as 32 x 32 store boxes
template <bool PANEL>
__global__ void __launch_bounds__(NTHREADS, 1)
rwr_chain_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmA,
                 const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmO,
                 const __grid_constant__ CUtensorMap tmS, ChainP p) {
	constexpr int SLOTS = Cfg<PANEL>::SLOTS;
	extern __shared__ uint8_t smem_raw[];
	uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
	uint8_t* stagebuf = smem + SLOTS * SLOT_BYTES;
	float* xch = (float*)(stagebuf + EPI_BYTES);
	uint64_t* bars = (uint64_t*)(stagebuf + EPI_BYTES + (PANEL ? XCH_FLOATS * 4 : 0));
	uint64_t* raw_full = bars;                    // TMA landed              (count 1 + tx)
	uint64_t* split_full = bars + SLOTS;          // lo half written         (count 4: splitter warps)
	uint64_t* empty = bars + 2 * SLOTS;           // MMAs reading the slot done (tcgen05.commit)
	uint64_t* acc_full = bars + 3 * SLOTS;        // [2] accumulator complete (tcgen05.commit)
	uint64_t* acc_empty = bars + 3 * SLOTS + 2;   // [2] accumulator drained  (count 8: drain warps)
	uint64_t* q_ready = bars + 3 * SLOTS + 4;     // Q hi/lo stored in TMEM   (count 8: drain warps)
	uint64_t* p_written = bars + 3 * SLOTS + 5;   // PANEL: P of the current cell is in the global scratch (count 1)
	uint32_t* tmem_holder = (uint32_t*)(bars + 3 * SLOTS + 6);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int nkb = (p.nb + BK - 1) / BK;         // k-blocks of the bin dimension (K of every product)
	const int NT = (p.ldw + BN - 1) / BN;         // 128-column tiles of the window
	const int nkw = PANEL ? (p.w + BK - 1) / BK : 0;  // k-blocks of the window (K of S2)
	const int n_last = (p.ldw - (NT - 1) * BN + 15) & ~15;  // MMA width of the last window tile (multiple of 16)
	const int box_last = (n_last + 31) / 32;                // its 32-column TMA boxes
	const bool chain = p.k > 1;

	if (threadIdx.x == 0) {
		for (int s = 0; s < SLOTS; ++s) {
			mbar_init(&raw_full[s], 1);
			mbar_init(&split_full[s], 4);
			mbar_init(&empty[s], 1);
		}
		mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
		mbar_init(&acc_empty[0], 8); mbar_init(&acc_empty[1], 8);
		mbar_init(q_ready, 8);
		mbar_init(p_written, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 0 && lane == 0) {
		asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmP) : "memory");
		asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
		if (PANEL) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmK) : "memory");
		if (p.tma_out) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmO) : "memory");
		if (PANEL) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmS) : "memory");
	}
	if (warp == 2) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(512) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem = *tmem_holder;

	if (warp == 0) {
		// ------------------------------------------------------------------ TMA producer
		if (lane == 0) {
			long long it = 0, ncell_done = 0;
			for (int cell = blockIdx.x; cell < p.ncell; cell += gridDim.x, ++ncell_done) {
				const int np = chain ? nkb : 0;
				const int nslots = nkw + np + NT * nkb;
				for (int j = 0; j < nslots; ++j, ++it) {
					const int s = (int)(it % SLOTS);
					mbar_wait(&empty[s], (uint32_t)(((it / SLOTS) & 1) ^ 1));
					uint8_t* dst = smem + s * SLOT_BYTES;
					const int nbox = (j >= nslots - nkb) ? box_last : BN / 32;  // the last window tile may be narrower
					mbar_expect_tx(&raw_full[s], j < nkw ? TILE_BYTES : nbox * (BK * 128));
					if (j < nkw) {  // S2: one K-major box of 128 rows x 32 window columns
						if (j == 0) FH_TRACE(0);
						tma_load_3d(dst, &tmK, &raw_full[s], j * BK, 0, cell);
						if (j == nkw - 1) FH_TRACE(1);
						continue;
					}
					const CUtensorMap* tm;
					int n0, k0, z;
					if (j < nkw + np) {
						if (PANEL && j == nkw) { mbar_wait(p_written, (uint32_t)(ncell_done & 1)); FH_TRACE(2); }  // transition done
						tm = &tmP; n0 = 0; k0 = (j - nkw) * BK; z = PANEL ? (int)blockIdx.x : cell;
					} else {
						const int t = j - nkw - np;
						tm = &tmA; n0 = (t / nkb) * BN; k0 = (t % nkb) * BK; z = cell;
					}
#pragma unroll
					for (int b = 0; b < BN / 32; ++b)
						if (b < nbox) tma_load_3d(dst + b * (BK * 128), tm, &raw_full[s], n0 + 32 * b, k0, z);
					if (j == nslots - 1) FH_TRACE(3);
					// next cell's panel -> L2 once this cell's P is on its way: the TMA unit is idle during the step
					// chain (issued at the top of the cell the prefetches queue AHEAD of the S2 loads: measured +4k clk)
					if (PANEL && !(p.dbg & 1) && j == nkw + np - 1 && cell + (int)gridDim.x < p.ncell) {
						for (int jj = 0; jj < nkw; ++jj) tma_prefetch_3d(&tmK, jj * BK, 0, cell + gridDim.x);
					}
				}
			}
		}
	} else if (warp == 1) {
		// ------------------------------------------------------------------ MMA issuer
		if (lane == 0) {
			const uint32_t idesc = make_idesc_tf32(false, true, BN, BM);
			// MN-major B tile (SWIZZLE_128B_BASE32B): LBO = BK*128 between 32-wide n groups, SBO = 512,
			// a K=8 step = 8 rows = +1024 B
			const uint32_t b_lbo = BK * 128, b_sbo = 512, b_step = 1024, b_lt = 1;
			long long it = 0, ch = 0, qn = 0;
			const uint32_t idesc_last = make_idesc_tf32(false, true, n_last, BM);
			auto issue_kb = [&](uint32_t acc, int kb, uint32_t slot_addr, uint32_t idesc) {
#pragma unroll
				for (int k4 = 0; k4 < BK / 8; ++k4) {
					const uint32_t a_hi = tmem + TM_QHI + (uint32_t)(kb * BK + k4 * 8), a_lo = tmem + TM_QLO + (uint32_t)(kb * BK + k4 * 8);
					const uint64_t dbh = make_desc(slot_addr + k4 * b_step, b_lbo, b_sbo, b_lt);
					const uint64_t dbl = make_desc(slot_addr + TILE_BYTES + k4 * b_step, b_lbo, b_sbo, b_lt);
					umma_tf32_ts(acc, a_lo, dbh, idesc, (kb | k4) ? 1u : 0u);  // small terms first
					umma_tf32_ts(acc, a_hi, dbl, idesc, 1u);
					umma_tf32_ts(acc, a_hi, dbh, idesc, 1u);
				}
			};
			const uint32_t idesc_kk = make_idesc_tf32(false, false, BN, BM);
			for (int cell = blockIdx.x; cell < p.ncell; cell += gridDim.x) {
				FH_TRACE(8);
				if (PANEL) {
					// S2 = A A^T: the landed K-major tile (SWIZZLE_128B: 128-byte rows, 8-row groups 1024 B apart, a K=8
					// step = +32 B) is both operands; accumulator chunks of CHUNK_KB k-blocks alternate buffers
					for (int kb = 0; kb < nkw; ++kb, ++it) {
						const int cb = (int)(ch & 1);
						const long long tw2 = p.trace ? clock64() : 0;
						if (kb % CHUNK_KB == 0) {
							mbar_wait(&acc_empty[cb], (uint32_t)(((ch >> 1) & 1) ^ 1));
							asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
						}
						const long long tw3 = p.trace ? clock64() : 0;
						const uint32_t acc = tmem + TM_ACC + (uint32_t)(cb * BN);
						const int s = (int)(it % SLOTS);
						mbar_wait(&split_full[s], (uint32_t)((it / SLOTS) & 1));
						if (p.trace && blockIdx.x == 0 && cell == 3 * (int)gridDim.x) { p.trace[6] += tw3 - tw2; p.trace[5] += clock64() - tw3; }
						asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
						const uint32_t hi = smem_u32(smem + s * SLOT_BYTES), lo = hi + TILE_BYTES;
#pragma unroll
						for (int k4 = 0; k4 < BK / 8; ++k4) {
							const uint64_t dh = make_desc(hi + k4 * 32, 16, 1024, 2), dl = make_desc(lo + k4 * 32, 16, 1024, 2);
							umma_tf32(acc, dl, dh, idesc_kk, ((kb % CHUNK_KB) | k4) ? 1u : 0u);  // small terms first
							umma_tf32(acc, dh, dl, idesc_kk, 1u);
							umma_tf32(acc, dh, dh, idesc_kk, 1u);
						}
						umma_commit(&empty[s]);
						if (kb % CHUNK_KB == CHUNK_KB - 1 || kb == nkw - 1) {
							umma_commit(&acc_full[cb]);
							++ch;
						}
					}
				}
				FH_TRACE(9);
				if (chain) {
					const long long p_it = it;  // the cell's P slots: p_it .. p_it + nkb - 1
					for (int step = 1; step < p.k; ++step) {
						mbar_wait(q_ready, (uint32_t)(qn & 1)); ++qn;
						const int cb = (int)(ch & 1);
						mbar_wait(&acc_empty[cb], (uint32_t)(((ch >> 1) & 1) ^ 1));
						asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
						const uint32_t acc = tmem + TM_ACC + (uint32_t)(cb * BN);
						for (int kb = 0; kb < nkb; ++kb) {
							const long long si = p_it + kb;
							const int s = (int)(si % SLOTS);
							if (step == 1) {
								mbar_wait(&split_full[s], (uint32_t)((si / SLOTS) & 1));
								asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
							}
							issue_kb(acc, kb, smem_u32(smem + s * SLOT_BYTES), idesc);
						}
						umma_commit(&acc_full[cb]);
						++ch;
					}
					for (int kb = 0; kb < nkb; ++kb) umma_commit(&empty[(int)((p_it + kb) % SLOTS)]);  // P no longer needed
					it += nkb;
				}
				FH_TRACE(10);
				mbar_wait(q_ready, (uint32_t)(qn & 1)); ++qn;   // Q_k
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				FH_TRACE(11);
				for (int nt = 0; nt < NT; ++nt) {
					const int cb = (int)(ch & 1);
					const long long tw1 = p.trace ? clock64() : 0;
					mbar_wait(&acc_empty[cb], (uint32_t)(((ch >> 1) & 1) ^ 1));
					if (p.trace && blockIdx.x == 0 && cell == 3 * (int)gridDim.x) p.trace[14] += clock64() - tw1;
					asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
					const uint32_t acc = tmem + TM_ACC + (uint32_t)(cb * BN);
					for (int kb = 0; kb < nkb; ++kb, ++it) {
						const int s = (int)(it % SLOTS);
						const long long tw0 = p.trace ? clock64() : 0;
						mbar_wait(&split_full[s], (uint32_t)((it / SLOTS) & 1));
						if (p.trace && blockIdx.x == 0 && cell == 3 * (int)gridDim.x) p.trace[13] += clock64() - tw0;
						asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
						issue_kb(acc, kb, smem_u32(smem + s * SLOT_BYTES), nt == NT - 1 ? idesc_last : idesc);
						umma_commit(&empty[s]);
					}
					umma_commit(&acc_full[cb]);
					++ch;
				}
				FH_TRACE(12);
			}
		}
	} else if (warp >= 4 && warp < 8) {
		// ------------------------------------------------------------------ splitters
		const int t = threadIdx.x - 128;  // 0..127
		long long it = 0;
		for (int cell = blockIdx.x; cell < p.ncell; cell += gridDim.x) {
			const int nslots = nkw + (chain ? nkb : 0) + NT * nkb;
			for (int j = 0; j < nslots; ++j, ++it) {
				const int s = (int)(it % SLOTS);
				const long long tw = (p.trace && t == 0) ? clock64() : 0;
				mbar_wait(&raw_full[s], (uint32_t)((it / SLOTS) & 1));
				if (p.trace && t == 0 && blockIdx.x == 0 && cell == 3 * (int)gridDim.x) p.trace[j < nkw ? 7 : 15] += clock64() - tw;
				const float4* hi = (const float4*)(smem + s * SLOT_BYTES);
				uint4* lo = (uint4*)(smem + s * SLOT_BYTES + TILE_BYTES);
				const int ni = (j >= nkw && j >= nslots - nkb) ? 2 * box_last : TILE_BYTES / 16 / 128;  // landed boxes only
#pragma unroll
				for (int i = 0; i < TILE_BYTES / 16 / 128; ++i) {
					if (i >= ni) break;
					const float4 v = hi[t + 128 * i];
					uint4 l;
					l.x = tf32_lo(v.x); l.y = tf32_lo(v.y); l.z = tf32_lo(v.z); l.w = tf32_lo(v.w);
					lo[t + 128 * i] = l;
				}
				asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (UMMA)
				__syncwarp();
				if (lane == 0) mbar_arrive(&split_full[s]);
			}
		}
	} else if (warp >= 8) {
		// ------------------------------------------------------------------ drain: accumulator -> P, Q / X
		const int q = warp & 3;             // TMEM lane quarter of this warp (rows 32q .. 32q+31)
		const int h = (warp - 8) >> 2;      // column half (64 columns)
		const int m = q * 32 + lane;        // this thread's row
		const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
		float* tile_s = (float*)(stagebuf + (warp - 8) * (32 * 32 * 4));
		// 32 values of Q (this thread's row, columns 64h + 32c ..) -> TMEM hi / lo
		auto store_q_chunk = [&](int c, const float (&qv)[32]) {
			uint32_t hi[32], lo[32];
#pragma unroll
			for (int j = 0; j < 32; ++j) { hi[j] = __float_as_uint(qv[j]); lo[j] = tf32_lo(qv[j]); }
			tmem_st32(tmem + lane_addr + TM_QHI + (uint32_t)(h * 64 + c * 32), hi);
			tmem_st32(tmem + lane_addr + TM_QLO + (uint32_t)(h * 64 + c * 32), lo);
		};
		auto publish_q = [&]() {
			tmem_st_wait();
			asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
			__syncwarp();
			if (lane == 0) mbar_arrive(q_ready);
		};
		// Q_1 = 0.5 P + 0.5 I from the transition matrix in global memory (first RWR step, Q_0 = I)   [!PANEL]
		auto first_q = [&](int cell) {
			const float* prow = p.P + (long long)cell * p.p_cell_stride + (long long)m * p.ldp + h * 64;
#pragma unroll
			for (int c = 0; c < 2; ++c) {
				float qv[32];
#pragma unroll
				for (int g = 0; g < 8; ++g) {
					const int col = h * 64 + c * 32 + 4 * g;
					float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
					if (m < p.nb && col < p.ldp) v = *reinterpret_cast<const float4*>(prow + c * 32 + 4 * g);
					const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
					for (int x = 0; x < 4; ++x) {
						float r = 0.f;
						if (m < p.nb && col + x < p.nb) r = 0.5f * e[x] + ((m == col + x) ? 0.5f : 0.f);
						qv[4 * g + x] = r;
					}
				}
				store_q_chunk(c, qv);
			}
			publish_q();
		};
		auto drain_sync = [&]() { asm volatile("bar.sync 1, 256;" ::: "memory"); };  // the 8 drain warps
		float* colpart = xch;                  // [4][128] per-row-quarter column sums
		float* rowsum = xch + 4 * 128;         // [2][128] S2 row sums of the two column halves
		float* cs1raw = xch + 6 * 128;         // column sums of the first-order block
		float* w1 = xch + 7 * 128;             // per-column weights of the first / second order parts of P
		float* w2 = xch + 8 * 128;
		float* cflag = xch + 9 * 128;          // empty column of the blend: its diagonal entry (partial_rwr.py:96-97)
		const int td = threadIdx.x - 256;      // 0..255 among the drain warps
		long long ch = 0;
		if (!PANEL && (int)blockIdx.x < p.ncell) first_q(blockIdx.x);
		for (int cell = blockIdx.x; cell < p.ncell; cell += gridDim.x) {
			if (PANEL) {
				// first-order block A[32q + r][s + 64h + 32c + lane], r = 0..31: COALESCED (a warp reads one 128-byte
				// row segment per instruction; one row per thread would touch 32 lines per instruction - measured 5 us)
				const float* ablk = p.A + (long long)cell * p.a_cell_stride + (long long)(q * 32) * p.ldw + p.s + h * 64 + lane;
				// ---- A (independent of S2: runs under the S2 MMAs): column sums of the first-order block; the block
				// itself is transposed to one row per lane through the staging tile (word (r, j) at r*32 + (j ^ r))
				// and parked in the Q_lo columns of TMEM, which are dead until this cell's Q_1 is written.
				// The two 32-column passes stay a loop: unrolled, ptxas hoists the second pass's 32 loads over the first
				// pass's transpose and spills (436 B of spill stores per thread; rolled: 12 B). Measured on a B200
				// (round 2): RWR stage 54.7 -> 51.8 ms per sweep, parity tests unchanged.
#pragma unroll 1
				for (int c = 0; c < 2; ++c) {
					const bool colok = h * 64 + c * 32 + lane < p.nb;
					float v[32];
#pragma unroll
					for (int r = 0; r < 32; ++r) v[r] = (colok && q * 32 + r < p.nb) ? ablk[(long long)r * p.ldw + c * 32] : 0.f;
					float cs = 0.f;
#pragma unroll
					for (int r = 0; r < 32; ++r) cs += v[r];
					colpart[q * 128 + h * 64 + c * 32 + lane] = cs;
					if (lane == 0) tma_store_wait_read();  // the tile may still feed an X store
					__syncwarp();
#pragma unroll
					for (int r = 0; r < 32; ++r) tile_s[r * 32 + (lane ^ r)] = v[r];
					__syncwarp();
					uint32_t fu[32];
#pragma unroll
					for (int j = 0; j < 32; ++j) fu[j] = __float_as_uint(tile_s[lane * 32 + (j ^ lane)]);
					__syncwarp();
					tmem_st32(tmem + lane_addr + TM_QLO + (uint32_t)(h * 64 + c * 32), fu);
				}
				tmem_st_wait();
				drain_sync();
				if (td < 128) cs1raw[td] = (colpart[td] + colpart[128 + td]) + (colpart[256 + td] + colpart[384 + td]);
				// ---- B: S2 row (this warp's 64 columns), chunks summed with round-to-nearest adds
				float sum[64];
#pragma unroll
				for (int j = 0; j < 64; ++j) sum[j] = 0.f;
				const int nchunk = (nkw + CHUNK_KB - 1) / CHUNK_KB;
				for (int chunk = 0; chunk < nchunk; ++chunk, ++ch) {
					const int cb = (int)(ch & 1);
					mbar_wait(&acc_full[cb], (uint32_t)((ch >> 1) & 1));
					asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
					for (int c = 0; c < 2; ++c) {
						uint32_t v[32];
						tmem_ld32(tmem + lane_addr + TM_ACC + (uint32_t)(cb * BN + h * 64 + c * 32), v);
#pragma unroll
						for (int j = 0; j < 32; ++j) sum[c * 32 + j] += __uint_as_float(v[j]);
					}
					asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
					__syncwarp();
					if (lane == 0) mbar_arrive(&acc_empty[cb]);
				}
				if (td == 0) FH_TRACE(16);
				// second-order affinity without its diagonal; S2 is symmetric, so its column sums are row sums
				float rs = 0.f;
#pragma unroll
				for (int j = 0; j < 64; ++j) {
					const int col = h * 64 + j;
					const float x = (m < p.nb && col < p.nb && col != m) ? sum[j] : 0.f;
					sum[j] = x;
					rs += x;
				}
				rowsum[h * 128 + m] = rs;
				drain_sync();
				if (td == 0) FH_TRACE(17);
				// per column j: P[i][j] = (3/4 f/(cs1+eps) + 1/4 h/(cs2+eps)) / (csl+eps) = f w1[j] + h w2[j]. The column
				// sum of the blend is taken analytically, csl = 3/4 cs1/(cs1+eps) + 1/4 cs2/(cs2+eps) (the reference
				// sums the rounded entries: same value to fp32 rounding), partial_rwr.py:88-97
				if (td < 128) {
					const float c1 = cs1raw[td], c2 = rowsum[td] + rowsum[128 + td];
					const float r1 = 1.f / (c1 + EPS), r2 = 1.f / (c2 + EPS);
					float csl = 0.75f * (c1 * r1) + 0.25f * (c2 * r2);
					const bool empty = (csl == 0.f) && td < p.nb;  // unreachable after the 1e-8 floor; kept for parity
					if (empty) csl = 1.f;
					const float rl = 1.f / (csl + EPS);
					w1[td] = 0.75f * r1 * rl;
					w2[td] = 0.25f * r2 * rl;
					cflag[td] = empty ? rl : 0.f;
				}
				drain_sync();
				if (td == 0) FH_TRACE(18);
				// ---- C: P -> global scratch (B operand of the chain, back through TMA); Q_1 = 0.5 P + 0.5 I -> TMEM
#pragma unroll
				for (int c = 0; c < 2; ++c) {
					uint32_t fu[32];
					tmem_ld32(tmem + lane_addr + TM_QLO + (uint32_t)(h * 64 + c * 32), fu);
					float pv[32];
#pragma unroll
					for (int j = 0; j < 32; ++j) {
						const int col = h * 64 + c * 32 + j;
						float x = __uint_as_float(fu[j]) * w1[col] + sum[c * 32 + j] * w2[col];  // 0 outside nb x nb
						if (m == col) x += cflag[col];
						pv[j] = x;
					}
					if (chain) {  // 128-byte-swizzled rows -> one TMA store of the 32 x 32 chunk
						if (lane == 0) tma_store_wait_read();
						__syncwarp();
						float4* rowp = reinterpret_cast<float4*>(tile_s) + lane * 8;
#pragma unroll
						for (int g = 0; g < 8; ++g) rowp[g ^ (lane & 7)] = make_float4(pv[4 * g], pv[4 * g + 1], pv[4 * g + 2], pv[4 * g + 3]);
						asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
						__syncwarp();
						if (lane == 0) {
							tma_store_3d(&tmS, tile_s, h * 64 + c * 32, q * 32, blockIdx.x);
							tma_store_commit();
						}
					}
					if (td == 0 && c == 0) FH_TRACE(29);
#pragma unroll
					for (int j = 0; j < 32; ++j) pv[j] = 0.5f * pv[j] + ((m == h * 64 + c * 32 + j && m < p.nb) ? 0.5f : 0.f);
					store_q_chunk(c, pv);
					if (td == 0 && c == 0) FH_TRACE(30);
				}
				if (td == 0) FH_TRACE(31);
				if (chain && lane == 0) tma_store_wait_all();  // P chunks written (visible to the producer's TMA loads)
				publish_q();
				if (td == 0) FH_TRACE(19);
				if (chain) {
					drain_sync();
					if (td == 0) mbar_arrive(p_written);
				}
			}
			for (int step = 1; step < p.k; ++step) {
				const int cb = (int)(ch & 1);
				mbar_wait(&acc_full[cb], (uint32_t)((ch >> 1) & 1));
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
				for (int c = 0; c < 2; ++c) {
					uint32_t v[32];
					tmem_ld32(tmem + lane_addr + TM_ACC + (uint32_t)(cb * BN + h * 64 + c * 32), v);
					float qv[32];
#pragma unroll
					for (int j = 0; j < 32; ++j) {
						float r = 0.5f * __uint_as_float(v[j]);
						if (m == h * 64 + c * 32 + j && m < p.nb) r += 0.5f;
						qv[j] = r;
					}
					store_q_chunk(c, qv);  // every MMA that read the old Q completed before acc_full fired
				}
				asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
				__syncwarp();
				if (lane == 0) mbar_arrive(&acc_empty[cb]);
				++ch;
				publish_q();
			}
			if (td == 0) FH_TRACE(20);
			float* ob = p.out + (long long)cell * p.out_cell_stride;
			for (int nt = 0; nt < NT; ++nt) {
				const int cb = (int)(ch & 1);
				mbar_wait(&acc_full[cb], (uint32_t)((ch >> 1) & 1));
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				// !PANEL: the last tile's completion frees Q: hand the next cell's Q_1 to the MMA warp first (its
				// first accumulator is the other buffer), then drain this tile
				if (!PANEL && nt == NT - 1 && cell + (int)gridDim.x < p.ncell) first_q(cell + gridDim.x);
				float sum[64];
#pragma unroll
				for (int c = 0; c < 2; ++c) {
					uint32_t v[32];
					tmem_ld32(tmem + lane_addr + TM_ACC + (uint32_t)(cb * BN + h * 64 + c * 32), v);
#pragma unroll
					for (int j = 0; j < 32; ++j) sum[c * 32 + j] = __uint_as_float(v[j]);
				}
				asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
				__syncwarp();
				if (lane == 0) mbar_arrive(&acc_empty[cb]);
				++ch;
				if (td == 0) FH_TRACE(21 + nt);
				if (p.tma_out) {
					// each lane lays its row into the warp's staging tile in the 128-byte-swizzle pattern (16-byte chunk g
					// of row r at g ^ (r & 7): four wavefronts per 512-byte store, the minimum), one TMA store per
					// 32 x 32 chunk; rows >= nb and columns >= ldw are clipped by the tensor map
#pragma unroll
					for (int c = 0; c < 2; ++c) {
						const int n0 = nt * BN + h * 64 + c * 32;
						if (lane == 0) tma_store_wait_read();
						__syncwarp();
						float4* rowp = reinterpret_cast<float4*>(tile_s) + lane * 8;
#pragma unroll
						for (int g = 0; g < 8; ++g)
							rowp[g ^ (lane & 7)] = make_float4(sum[c * 32 + 4 * g], sum[c * 32 + 4 * g + 1], sum[c * 32 + 4 * g + 2], sum[c * 32 + 4 * g + 3]);
						asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
						__syncwarp();
						if (lane == 0 && n0 < p.ldw && q * 32 < p.nb) {
							tma_store_3d(&tmO, tile_s, n0, q * 32, cell);
							tma_store_commit();
						}
					}
					continue;
				}
				// fallback (unaligned output): 32 x 32 transposes through an XOR-swizzled private buffer (word (r, j) at
				// r*32 + (j ^ r)), then each lane owns 4 consecutive columns of 8 rows
				const int cg = lane & 7, ro = lane >> 3;
#pragma unroll
				for (int c = 0; c < 2; ++c) {
					__syncwarp();
#pragma unroll
					for (int j = 0; j < 32; ++j) tile_s[lane * 32 + (j ^ lane)] = sum[c * 32 + j];
					__syncwarp();
					const int n = nt * BN + h * 64 + c * 32 + 4 * cg;  // first of this lane's 4 columns
					if (n < p.ldw) {  // pad columns [w, ldw) receive the exact zeros of the out-of-bounds B columns
#pragma unroll 1
						for (int itr = 0; itr < 8; ++itr) {
							const int rr = itr * 4 + ro;
							const int r = q * 32 + rr;
							if (r >= p.nb) continue;
							float4 v = reinterpret_cast<const float4*>(tile_s)[rr * 8 + (cg ^ itr)];
							if (ro & 1) { float t0 = v.x; v.x = v.y; v.y = t0; t0 = v.z; v.z = v.w; v.w = t0; }
							if (ro & 2) { float t0 = v.x; v.x = v.z; v.z = t0; t0 = v.y; v.y = v.w; v.w = t0; }
							float* cp = ob + (long long)r * p.ldw + n;
							if (p.vec_ok) *reinterpret_cast<float4*>(cp) = v;
							else { cp[0] = v.x; cp[1] = v.y; cp[2] = v.z; cp[3] = v.w; }
						}
					}
				}
			}
			if (td == 0) FH_TRACE(25);
		}
		if (lane == 0) tma_store_wait_all();
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 2) {
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
	}
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

// A: (ncell, nb, ldw) conv'd panels; out: cell c at out + c * out_cell_stride, rows of ldw floats.
// scratch == nullptr: P (ncell, nb, ldp) holds the column-stochastic transition matrices (chain + Q A only).
// scratch != nullptr (fh_rwr_chain_scratch_bytes() bytes): P is ignored, S2 and the transition matrix
// are formed in the kernel from A and the diagonal block offset s.
// Returns FH_ERR_UNSUPPORTED (nothing launched) when the shape is outside the fused kernel's range -
// the caller runs the per-step kernels.
size_t fh_rwr_chain_scratch_bytes() { return (size_t)256 * 128 * 128 * 4; }

int fh_rwr_chain(const float* P, const float* A, float* out, int nb, int w, int ldw, int ldp, int s, int k, int ncell,
                 long long p_cell_stride, long long a_cell_stride, long long out_cell_stride, float* scratch,
                 void* stream) {
	if (ncell <= 0) return FH_OK;
	const bool panel = scratch != nullptr;
	if (nb > BM || k < 1 || (ldw & 3) || (a_cell_stride & 3) || !aligned16(A) || ncell > 65535 ||
	    (!panel && ((ldp & 3) || (p_cell_stride & 3) || !aligned16(P))) || (panel && !aligned16(scratch))) {
		fh_set_error("fh_rwr_chain: shape outside the fused kernel (nb <= 128, k >= 1, 16-byte aligned rows)");
		return FH_ERR_UNSUPPORTED;
	}
	static int num_sms = 0;
	if (!num_sms) {
		int dev = 0;
		FH_CUDA(cudaGetDevice(&dev));
		FH_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
	}
	int grid = ncell < num_sms ? ncell : num_sms;
	if (grid > 256) grid = 256;  // scratch holds 256 CTAs
	CUtensorMap tp, ta, tk, to, ts;
	// (k rows x n contiguous) MN-major operands; the contiguous extent is the LOGICAL width, so pad columns and
	// rows beyond nb read as zeros whatever the buffers hold
	bool ok = make_map(&ta, A, w, nb, ldw, ncell, a_cell_stride, 32, BK, true);
	if (panel) {
		ok = ok && make_map(&tp, scratch, nb, nb, 128, grid, 128 * 128, 32, BK, true) &&
		     make_map(&ts, scratch, 128, 128, 128, grid, 128 * 128, 32, 32, false) &&
		     make_map(&tk, A, w, nb, ldw, ncell, a_cell_stride, BK, BM, false);
	} else {
		ok = ok && make_map(&tp, P, nb, nb, ldp, ncell, p_cell_stride, 32, BK, true);
		tk = ta; ts = ta;
	}
	if (!ok) {
		fh_set_error("fh_rwr_chain: cuTensorMapEncodeTiled failed");
		return FH_ERR_UNSUPPORTED;
	}
	ChainP p;
	p.nb = nb; p.w = w; p.ldw = ldw; p.ldp = ldp; p.k = k; p.ncell = ncell;
	p.p_cell_stride = p_cell_stride; p.out_cell_stride = out_cell_stride;
	p.P = panel ? scratch : P; p.A = A; p.a_cell_stride = a_cell_stride; p.s = s; p.out = out;
	p.vec_ok = aligned16(out) && (out_cell_stride % 4 == 0);
	p.tma_out = p.vec_ok && make_map(&to, out, ldw, nb, ldw, ncell, out_cell_stride, 32, 32, false);
	static int dbg = -1;
	if (dbg < 0) { const char* e = getenv("FH_CHAIN_DEBUG"); dbg = e ? atoi(e) : 0; }
	p.dbg = dbg;
	if (dbg & 2) p.tma_out = 0;
	if (!p.tma_out) to = ta;
	cudaStream_t st = (cudaStream_t)stream;
	static int trace_on = -1;
	if (trace_on < 0) { const char* e = getenv("FH_CHAIN_TRACE"); trace_on = (e && e[0] == '1') ? 1 : 0; }
	p.trace = nullptr;
	if (trace_on) {
		FH_CUDA(cudaMalloc(&p.trace, 32 * sizeof(long long)));
		FH_CUDA(cudaMemsetAsync(p.trace, 0, 32 * sizeof(long long), st));
	}
	if (panel) FH_CUDA(cudaFuncSetAttribute(rwr_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<true>::SMEM_BYTES));
	else FH_CUDA(cudaFuncSetAttribute(rwr_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<false>::SMEM_BYTES));
	const int tmr = fh_time_begin(FH_TIME_RWR_CHAIN, st);
	if (panel) rwr_chain_kernel<true><<<grid, NTHREADS, Cfg<true>::SMEM_BYTES, st>>>(tp, ta, tk, to, ts, p);
	else rwr_chain_kernel<false><<<grid, NTHREADS, Cfg<false>::SMEM_BYTES, st>>>(tp, ta, tk, to, ts, p);
	fh_time_end(tmr, st);
	FH_LAUNCH_CHECK();
	if (trace_on) {  // debug only: synchronises and prints the phase stamps (SM clocks relative to stamp 0 / 8)
		long long h[32];
		FH_CUDA(cudaMemcpyAsync(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost, st));
		FH_CUDA(cudaStreamSynchronize(st));
		FH_CUDA(cudaFree(p.trace));
		long long t0 = h[8] ? h[8] : h[0];
		fprintf(stderr, "[chain trace] nb=%d w=%d k=%d:", nb, w, k);
		for (int i = 0; i < 32; ++i)  // 5/6/13/14 are accumulated wait times, the rest stamps
			if (h[i]) fprintf(stderr, " %d:%lld", i, (i == 5 || i == 6 || i == 7 || i == 13 || i == 14 || i == 15) ? h[i] : h[i] - t0);
		fprintf(stderr, "\n");
	}
	return FH_OK;
}
