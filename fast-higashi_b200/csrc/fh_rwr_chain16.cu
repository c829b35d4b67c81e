// Fused RWR imputation of one bin block from the conv'd panel A (reference: partial_rwr.py:80-126, :138), 3xFP16 variant:
//   S2 = A A^T;  P = colnorm(3/4 colnorm(A_diag_block) + 1/4 colnorm(S2 - diag))
//   Q_1 = 0.5 P + 0.5 I,   Q_{t+1} = 0.5 Q_t P + 0.5 I  (t = 1 .. k-1),   X = Q_k A
// for every cell of the chunk in ONE persistent tcgen05 kernel, like fh_rwr_chain.cu, but every operand is a pair of
// binary16 values (hi = rn16(s x), lo = rn16(s x - hi), s a power of two: 22 significand bits, the same as the two TF32
// halves) and the three products hi hi + hi lo + lo hi run as kind::f16 MMAs: twice the tensor-pipe rate of kind::tf32 and
// half the shared-memory bytes per operand (scripts/split_precision_study.py: operand error 1.2e-7 on the RWR chain).
// What that changes in the data flow:
//   * the conv'd panel arrives ALREADY SPLIT: densify_conv_kernel<.., true> (fh_rwr.cu) writes it as two binary16 planes
//     (same bytes as fp32), scaled per cell by the power of two that brings its largest CSR value into [2^13, 2^14)
//     (device words `amax[cell]`); TMA lands the tiles in the layouts the MMAs read (K-major SWIZZLE_128B for S2, MN-major
//     SWIZZLE_128B for X = Q A) - there are no splitter warps and no generic-proxy pass over the ring;
//   * P never leaves the SM: the drain warps write its hi / lo tiles (scaled by 2^14) straight into a dedicated
//     shared-memory operand (MN-major SWIZZLE_128B, conflict-free 16-byte stores) - no global scratch round trip;
//   * Q (scaled by 2^14) lives in TENSOR MEMORY as packed halves (two per 32-bit column) and is the A operand of every
//     MMA of the chain and of X = Q A.
//   * the first-order block A[:, s:s+nb] is TMA-loaded (both planes) into the shared-memory region that P will occupy:
//     the K-major tile of a 64-column box and the MN-major P operand put row m of column half h at the same 128 bytes,
//     so every drain thread turns its own row of the block into its own row of P in place (no transposes through
//     staging tiles, no global loads in the drain warps: those cost 14k of a cell's 47k cycles in the first version).
//   * S2 of the NEXT cell is issued at the top of a cell's transition into its own TMEM accumulator, and the column sums of
//     the first-order block come from a ones x F product: the tensor pipe works while the drain warps build P.
// Measured on a B200 (FH_CHAIN_TRACE=1, nb = 115, w = 315, k = 4, SM cycles per cell): 47k in the first version -> 37k
// (3xTF32 kernel: 64k). What a cell costs now: S2 (next cell) ~10k under the transition (~14k: the drain warps' shared-memory
// traffic for P competes with the UMMA operand reads, both at the 128 B/clk limit), 3 chain steps 3 x 3.1k (24 MMAs ~2k +
// drain ~1k), X = Q A ~9.5k (epilogue paced: TMEM read, staging writes, TMA stores) + 3k tail. Shared-memory bytes per cell
// (operand reads + TMA + staging) ~2.1 MB = 16k cycles at 128 B/clk: the floor of this design.
// TMEM (512 columns): Q_hi [0,64) | Q_lo [64,128) | S2 accumulator [128,256) | accumulators [256,384) [384,512)
// Roles: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-19 drain (lane quarter x column quarter).
// Every product has K <= ~330 (S2) or K = nb <= 128 (chain, X): one accumulation in TMEM each (fh_gemm_tc.cu drains long-K
// sums every 128 because the tensor core's accumulation truncates; at these K the effect is <= 2e-6).
#include <cuda_fp16.h>
#include "fh_tc.cuh"
#include "../../include/fh_b200.h"

namespace {
using namespace fh_tc;

constexpr int BM = 128, BN = 128, BK = 64;      // k-block: 64 halves = one 128-byte swizzled row
constexpr int PLANE_BYTES = BK * BN * 2;        // 16 KB: the hi (or lo) tile of a slot
constexpr int SLOT_BYTES = 2 * PLANE_BYTES;     // hi | lo
constexpr int SLOTS = 3;
constexpr int P_PLANE = 128 * 128 * 2;          // P hi (or lo): [n group (2)][k row (128)][128 B]
constexpr int EPI_BYTES = 16 * 32 * 16 * 4;     // 16 drain warps x (32 rows x 16 floats)
constexpr int XCH_FLOATS = 12 * 128;            // row sums [4][128], cs1 / w1 / w2 / flag [128], do_col: 1 / bin_cov [512]
constexpr int ONES_BYTES = 4096;                // 128 x 16 halves of 1.0: the A operand of the column-sum MMAs
constexpr int NTHREADS = 640;
constexpr int SMEM_BYTES = SLOTS * SLOT_BYTES + 2 * P_PLANE + EPI_BYTES + XCH_FLOATS * 4 + ONES_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr float EPS = 1e-15f;                   // partial_rwr.py:88-97
constexpr uint32_t TM_QHI = 0, TM_QLO = 64, TM_S2 = 128, TM_ACC = 256;
constexpr float QS = 16384.f;                   // scale of Q and P (entries in [0, 1])
constexpr float QS_INV = 1.f / 16384.f;

struct Chain16P {
	int nb, w, ldw, ld16, k, ncell;
	int pad;               // plane column of window column 0: (s + pad) % 8 == 0 (TMA box starts must be 16-byte aligned)
	long long a_cell_stride, out_cell_stride;  // halves / floats
	const __half* Ahi;     // planes: hi at Ahi, lo at Ahi + ncell * a_cell_stride
	const unsigned* amax;  // [ncell] bits of the largest (floored) CSR value of each cell's slice: fixes that panel's scale
	int s;
	float* out;
	const float* bin_cov;  // do_col: [ncell][>= w] per-cell coverage of the window columns (row stride cov_ld), else nullptr
	long long cov_ld;
	long long* trace;      // FH_CHAIN_TRACE=1: clock64 stamps of CTA 0's 4th cell (debug)
};
#define FH_TRACE(slot)                                                         \
	do {                                                                       \
		if (p.trace && blockIdx.x == 0 && cell == 3 * (int)gridDim.x) p.trace[slot] = clock64(); \
	} while (0)

// tmK: the planes as K-major boxes of 64 window columns x 128 rows (S2); tmA: as MN-major boxes of 64 window columns x
// 64 bin rows (B tiles of X = Q A); plane / cell folded: z = cell (hi), ncell + cell (lo); tmO: X as store boxes of 32 rows x 16 columns
__global__ void __launch_bounds__(NTHREADS, 1)
rwr_chain16_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmA,
                   const __grid_constant__ CUtensorMap tmO, Chain16P p) {
	extern __shared__ uint8_t smem_raw[];
	uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
	uint8_t* pbuf = smem + SLOTS * SLOT_BYTES;                 // P hi | P lo
	uint8_t* stagebuf = pbuf + 2 * P_PLANE;
	float* xch = (float*)(stagebuf + EPI_BYTES);
	uint8_t* ones = stagebuf + EPI_BYTES + XCH_FLOATS * 4;
	uint64_t* bars = (uint64_t*)(ones + ONES_BYTES);
	uint64_t* full = bars;                        // TMA landed                  (count 1 + tx)
	uint64_t* empty = bars + SLOTS;               // MMAs reading the slot done  (tcgen05.commit)
	uint64_t* acc_full = bars + 2 * SLOTS;        // [2] accumulator complete    (tcgen05.commit)
	uint64_t* acc_empty = bars + 2 * SLOTS + 2;   // [2] accumulator drained     (count 16: drain warps)
	uint64_t* q_ready = bars + 2 * SLOTS + 4;     // Q hi / lo stored in TMEM    (count 16: drain warps)
	uint64_t* p_written = bars + 2 * SLOTS + 5;   // P of the current cell is in shared memory (count 1)
	uint64_t* f_full = bars + 2 * SLOTS + 6;      // first-order block landed in the P region (count 1 + tx)
	uint64_t* f_free = bars + 2 * SLOTS + 7;      // P region no longer read: chain MMAs done (tcgen05.commit) / k = 1, do_col: drain
	uint64_t* s2_full = bars + 2 * SLOTS + 8;     // S2 of a cell complete in its own accumulator (tcgen05.commit)
	uint64_t* s2_empty = bars + 2 * SLOTS + 9;    // ... and read by the drain warps (count 16)
	uint32_t* tmem_holder = (uint32_t*)(bars + 2 * SLOTS + 10);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int nkb = (p.nb + BK - 1) / BK;         // k-blocks of the bin dimension (K of X = Q A)
	const int NT = (p.ldw + p.pad + BN - 1) / BN; // 128-column tiles of the (shifted) window
	const int nkw = (p.w + p.pad + BK - 1) / BK;  // k-blocks of the (shifted) window (K of S2; the pad columns are zeros)
	const int n_last = (p.ldw + p.pad - (NT - 1) * BN + 15) & ~15;  // MMA width of the last window tile (multiple of 16)
	const int box_last = (n_last + 63) / 64;                // its 64-column TMA boxes
	const int n_step = (p.nb + 15) & ~15;                   // MMA width / K extent of the chain's products
	const bool chain = p.k > 1;

	if (threadIdx.x == 0) {
		for (int s = 0; s < SLOTS; ++s) {
			mbar_init(&full[s], 1);
			mbar_init(&empty[s], 1);
		}
		mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
		mbar_init(&acc_empty[0], 16); mbar_init(&acc_empty[1], 16);
		mbar_init(q_ready, 16);
		mbar_init(p_written, 1);
		mbar_init(f_full, 1);
		mbar_init(f_free, 1);
		mbar_init(s2_full, 1);
		mbar_init(s2_empty, 16);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 0 && lane == 0) {
		asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmK) : "memory");
		asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
		asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmO) : "memory");
	}
	for (int i = threadIdx.x; i < ONES_BYTES / 4; i += NTHREADS) reinterpret_cast<uint32_t*>(ones)[i] = 0x3C003C00u;  // (1.0, 1.0)
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
			// first-order block of a cell: two 64-column boxes per plane at window column s -> the P region
			auto load_first_order = [&](int cell) {
				mbar_expect_tx(f_full, 4 * PLANE_BYTES);
				for (int g = 0; g < 2; ++g) {
					tma_load_3d(pbuf + g * PLANE_BYTES, &tmK, f_full, p.s + p.pad + 64 * g, 0, cell);
					tma_load_3d(pbuf + P_PLANE + g * PLANE_BYTES, &tmK, f_full, p.s + p.pad + 64 * g, 0, p.ncell + cell);
				}
			};
			// S2 tiles of a cell: one K-major box of 128 rows x 64 window columns per plane and slot
			auto load_s2 = [&](int cell) {
				for (int j = 0; j < nkw; ++j, ++it) {
					const int s = (int)(it % SLOTS);
					mbar_wait(&empty[s], (uint32_t)(((it / SLOTS) & 1) ^ 1));
					uint8_t* dst = smem + s * SLOT_BYTES;
					mbar_expect_tx(&full[s], SLOT_BYTES);
					tma_load_3d(dst, &tmK, &full[s], j * BK, 0, cell);
					tma_load_3d(dst + PLANE_BYTES, &tmK, &full[s], j * BK, 0, p.ncell + cell);
				}
				// the panel after it -> L2 (its S2 loads follow one cell later)
				if (cell + (int)gridDim.x < p.ncell)
					for (int jj = 0; jj < nkw; ++jj) {
						tma_prefetch_3d(&tmK, jj * BK, 0, cell + gridDim.x);
						tma_prefetch_3d(&tmK, jj * BK, 0, p.ncell + cell + gridDim.x);
					}
			};
			// ring order = the MMA warp's order: S2(first); per cell: S2(next cell), then the X = Q A tiles of the cell
			if ((int)blockIdx.x < p.ncell) {
				load_first_order(blockIdx.x);
				load_s2(blockIdx.x);
			}
			for (int cell = blockIdx.x; cell < p.ncell; cell += gridDim.x, ++ncell_done) {
				if (cell == blockIdx.x + 3 * (int)gridDim.x) FH_TRACE(0);
				if (cell + (int)gridDim.x < p.ncell) load_s2(cell + gridDim.x);
				FH_TRACE(1);
				for (int t = 0; t < NT * nkb; ++t, ++it) {
					const int s = (int)(it % SLOTS);
					mbar_wait(&empty[s], (uint32_t)(((it / SLOTS) & 1) ^ 1));
					uint8_t* dst = smem + s * SLOT_BYTES;
					const int nt = t / nkb, kb = t % nkb;
					const int nbox = (nt == NT - 1) ? box_last : BN / 64;  // the last window tile may be narrower
					mbar_expect_tx(&full[s], 2 * nbox * (BK * 128));
					for (int g = 0; g < nbox; ++g) {
						tma_load_3d(dst + g * (BK * 128), &tmA, &full[s], nt * BN + 64 * g, kb * BK, cell);
						tma_load_3d(dst + PLANE_BYTES + g * (BK * 128), &tmA, &full[s], nt * BN + 64 * g, kb * BK, p.ncell + cell);
					}
				}
				FH_TRACE(3);
				// the next cell's first-order block, once this cell's chain no longer reads P (long past by now)
				if (cell + (int)gridDim.x < p.ncell) {
					mbar_wait(f_free, (uint32_t)(ncell_done & 1));
					load_first_order(cell + gridDim.x);
				}
			}
		}
	} else if (warp == 1) {
		// ------------------------------------------------------------------ MMA issuer
		if (lane == 0) {
			const uint32_t idesc_kk = make_idesc_f16(false, false, BN, BM);        // S2: both operands K-major
			const uint32_t idesc_st = make_idesc_f16(false, true, n_step, BM);     // chain: B = P, MN-major
			const uint32_t idesc_x = make_idesc_f16(false, true, BN, BM);          // X: B = panel tile, MN-major
			const uint32_t idesc_xl = make_idesc_f16(false, true, n_last, BM);
			const uint32_t p_hi = smem_u32(pbuf), p_lo = p_hi + P_PLANE;
			long long it = 0, ch = 0, qn = 0, ncell_done = 0;
			// S2 = A A^T of the n-th cell of this CTA into its own accumulator: the landed K-major tiles (SWIZZLE_128B: 128-byte
			// rows, 8-row groups 1024 B apart, a K = 16 step = +32 B) are both operands. K = w <= ~330 is accumulated in one go
			// (the tensor core's truncating accumulation costs ~7e-9 K relative, 2e-6 here, on the quarter-weight term of P).
			// It is issued ONE CELL AHEAD, at the top of the previous cell's transition, so that the tensor pipe works on it
			// while the drain warps build that cell's P (they idle the pipe for ~9k of a cell's ~43k cycles otherwise).
			auto issue_s2 = [&](long long n) {
				mbar_wait(s2_empty, (uint32_t)((n & 1) ^ 1));  // the drain warps have read the previous cell's S2
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				const uint32_t acc = tmem + TM_S2;
				for (int kb = 0; kb < nkw; ++kb, ++it) {
					const int s = (int)(it % SLOTS);
					mbar_wait(&full[s], (uint32_t)((it / SLOTS) & 1));
					asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
					const uint32_t hi = smem_u32(smem + s * SLOT_BYTES), lo = hi + PLANE_BYTES;
					const int nk = min(BK, p.w + p.pad - kb * BK);
					const int nk4 = (nk + 15) >> 4;
					for (int k4 = 0; k4 < nk4; ++k4) {
						const uint64_t dh = make_desc(hi + k4 * 32, 16, 1024, 2), dl = make_desc(lo + k4 * 32, 16, 1024, 2);
						umma_f16(acc, dl, dh, idesc_kk, (kb | k4) ? 1u : 0u);  // small terms first
						umma_f16(acc, dh, dl, idesc_kk, 1u);
						umma_f16(acc, dh, dh, idesc_kk, 1u);
					}
					umma_commit(&empty[s]);
				}
				umma_commit(s2_full);
			};
			if ((int)blockIdx.x < p.ncell) issue_s2(0);
			for (int cell = blockIdx.x; cell < p.ncell; cell += gridDim.x, ++ncell_done) {
				FH_TRACE(8);
				{
					// column sums of the first-order block on the tensor core: ones (128 x K) x F (K = rows, MN-major, the layout P
					// will have) -> every accumulator row holds the column sums (hi and lo planes summed); 2 nb / 16 MMAs instead
					// of a pass over the block by the drain warps (3k cycles of their critical path)
					mbar_wait(f_full, (uint32_t)(ncell_done & 1));
					const int cb = (int)(ch & 1);
					mbar_wait(&acc_empty[cb], (uint32_t)(((ch >> 1) & 1) ^ 1));
					asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
					const uint32_t acc = tmem + TM_ACC + (uint32_t)(cb * BN);
					const uint64_t da = make_desc(smem_u32(ones), 128, 256, 0);  // no swizzle: 8-row x 16-byte core matrices
					for (int ks = 0; ks < (n_step >> 4); ++ks) {
						umma_f16(acc, da, make_desc(p_lo + ks * 2048, 128 * 128, 1024, 2), idesc_st, ks ? 1u : 0u);
						umma_f16(acc, da, make_desc(p_hi + ks * 2048, 128 * 128, 1024, 2), idesc_st, 1u);
					}
					umma_commit(&acc_full[cb]);
					++ch;
				}
				if (cell + (int)gridDim.x < p.ncell) issue_s2(ncell_done + 1);
				FH_TRACE(9);
				if (chain) {
					// Q_{t+1} = 0.5 Q_t P + 0.5 I: A = Q from TMEM (K = 16 step = +8 columns), B = P in shared memory (MN-major:
					// 64-wide n groups 16 KB apart, 8-row k groups 1024 B apart, a K = 16 step = +2048 B)
					mbar_wait(p_written, (uint32_t)(ncell_done & 1));
					asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
					const int nks = n_step >> 4;
					for (int step = 1; step < p.k; ++step) {
						mbar_wait(q_ready, (uint32_t)(qn & 1)); ++qn;
						if (step == 2) FH_TRACE(30);
						const int cb = (int)(ch & 1);
						mbar_wait(&acc_empty[cb], (uint32_t)(((ch >> 1) & 1) ^ 1));
						asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
						const uint32_t acc = tmem + TM_ACC + (uint32_t)(cb * BN);
						for (int ks = 0; ks < nks; ++ks) {
							const uint32_t a_hi = tmem + TM_QHI + (uint32_t)(8 * ks), a_lo = tmem + TM_QLO + (uint32_t)(8 * ks);
							const uint64_t dbh = make_desc(p_hi + ks * 2048, 128 * 128, 1024, 2);
							const uint64_t dbl = make_desc(p_lo + ks * 2048, 128 * 128, 1024, 2);
							umma_f16_ts(acc, a_lo, dbh, idesc_st, ks ? 1u : 0u);  // small terms first
							umma_f16_ts(acc, a_hi, dbl, idesc_st, 1u);
							umma_f16_ts(acc, a_hi, dbh, idesc_st, 1u);
						}
						umma_commit(&acc_full[cb]);
						++ch;
						if (step == 1) FH_TRACE(29);
					}
					if (!p.bin_cov) umma_commit(f_free);  // P is no longer read once these MMAs have completed (do_col: the drain
					                                      // warps reuse the region for the transpose of Q and release it)
				}
				FH_TRACE(10);
				mbar_wait(q_ready, (uint32_t)(qn & 1)); ++qn;   // Q_k
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				FH_TRACE(11);
				for (int nt = 0; nt < NT; ++nt) {
					const int cb = (int)(ch & 1);
					mbar_wait(&acc_empty[cb], (uint32_t)(((ch >> 1) & 1) ^ 1));
					asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
					const uint32_t acc = tmem + TM_ACC + (uint32_t)(cb * BN);
					const uint32_t idesc = nt == NT - 1 ? idesc_xl : idesc_x;
					for (int kb = 0; kb < nkb; ++kb, ++it) {
						const int s = (int)(it % SLOTS);
						mbar_wait(&full[s], (uint32_t)((it / SLOTS) & 1));
						asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
						const uint32_t hi = smem_u32(smem + s * SLOT_BYTES), lo = hi + PLANE_BYTES;
						const int nk4 = (min(BK, n_step - kb * BK) + 15) >> 4;
						for (int k4 = 0; k4 < nk4; ++k4) {
							const uint32_t qc = (uint32_t)(kb * (BK / 2) + 8 * k4);
							const uint64_t dbh = make_desc(hi + k4 * 2048, BK * 128, 1024, 2);
							const uint64_t dbl = make_desc(lo + k4 * 2048, BK * 128, 1024, 2);
							umma_f16_ts(acc, tmem + TM_QLO + qc, dbh, idesc, (kb | k4) ? 1u : 0u);  // small terms first
							umma_f16_ts(acc, tmem + TM_QHI + qc, dbl, idesc, 1u);
							umma_f16_ts(acc, tmem + TM_QHI + qc, dbh, idesc, 1u);
						}
						umma_commit(&empty[s]);
					}
					umma_commit(&acc_full[cb]);
					++ch;
				}
				FH_TRACE(12);
			}
		}
	} else if (warp >= 4) {
		// ------------------------------------------------------------------ drain: accumulator -> P, Q / X
		// 16 warps: lane quarter q (rows 32q ..) x column quarter h (columns 32h ..). The drain phases are chains of dependent
		// ALU work between TMEM / shared-memory accesses (ncu: ~5 cycles per instruction and warp, the tensor pipe idle
		// meanwhile), so they are spread over four warps per scheduler instead of two, 32 columns per thread.
		const int q = warp & 3;             // TMEM lane quarter of this warp (rows 32q .. 32q+31)
		const int h = (warp - 4) >> 2;      // column quarter (32 columns)
		const int m = q * 32 + lane;        // this thread's row
		const bool diag_chunk = (q == h);   // warp-uniform: rows 32q.. meet columns 32h.. (column - first column = lane)
		const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
		float* tile_s = (float*)(stagebuf + (warp - 4) * (32 * 16 * 4));  // one 32 x 16 staging buffer per warp
		float sa_inv, s2_scale, x_scale;  // the current cell's panel scale (amax[cell] * sa in [2^13, 2^14)) undone
		auto publish_q = [&]() {
			tmem_st_wait();
			asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
			__syncwarp();
			if (lane == 0) mbar_arrive(q_ready);
		};
		auto drain_sync = [&]() { asm volatile("bar.sync 1, 512;" ::: "memory"); };  // the 16 drain warps
		float* rowsum = xch;                   // [4][128] S2 row sums of the four column quarters
		float* cs1raw = xch + 4 * 128;         // column sums of the first-order block
		float* w1 = xch + 5 * 128;             // per-column weights of the first / second order parts of P
		float* w2 = xch + 6 * 128;
		float* cflag = xch + 7 * 128;          // empty column of the blend: its diagonal entry (partial_rwr.py:96-97)
		float* rcov = xch + 8 * 128;           // do_col: 1 / bin_cov by (shifted) window column, <= 512
		const int td = threadIdx.x - 128;      // 0..511 among the drain warps
		// this thread's row of the first-order block / of P in the shared-memory operand: row m of the 64-column group
		// h >> 1; its 16-byte chunks (8 columns) 4 (h & 1) + g at (chunk ^ (m & 7)); lo plane P_PLANE bytes further
		uint8_t* prow = pbuf + (h >> 1) * (128 * 128) + m * 128;
		const int chunk0 = (h & 1) * 4;
		auto halves_to_floats = [&](const uint4& vh, const uint4& vl, float (&f)[8]) {
			const uint32_t hh[4] = {vh.x, vh.y, vh.z, vh.w}, ll[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
			for (int e = 0; e < 4; ++e) {
				const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hh[e]));
				const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&ll[e]));
				f[2 * e] = a.x + b.x;
				f[2 * e + 1] = a.y + b.y;
			}
		};
		long long ch = 0, ncell_done = 0;
		for (int cell = blockIdx.x; cell < p.ncell; cell += gridDim.x, ++ncell_done) {
			{
				const unsigned abits = max(p.amax[cell], __float_as_uint(1e-8f));
				sa_inv = __uint_as_float(((abits >> 23) - 13u) << 23);
				s2_scale = sa_inv * sa_inv;
				x_scale = sa_inv * (1.f / QS);
			}
			// ---- A: column sums of the first-order block from the tensor core's ones x F product: all accumulator rows are
			// equal, lane j of the quarter-0 warps keeps column 32h + j (rows >= nb and window columns >= w are the TMA's
			// zeros; columns >= nb are masked)
			mbar_wait(f_full, (uint32_t)(ncell_done & 1));
			{
				const int cb = (int)(ch & 1);
				mbar_wait(&acc_full[cb], (uint32_t)((ch >> 1) & 1));
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				if (q == 0) {
					uint32_t v[32];
					tmem_ld32(tmem + TM_ACC + (uint32_t)(cb * BN + h * 32), v);
					float cs = 0.f;
#pragma unroll
					for (int j = 0; j < 32; ++j)
						if (lane == j) cs = __uint_as_float(v[j]);
					cs1raw[h * 32 + lane] = h * 32 + lane < p.nb ? cs * sa_inv : 0.f;
				}
				asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
				__syncwarp();
				if (lane == 0) mbar_arrive(&acc_empty[cb]);
				++ch;
			}
			if (td == 0) FH_TRACE(26);
			// ---- B: S2 row (this warp's 32 columns)
			float sum[32];
			{
				mbar_wait(s2_full, (uint32_t)(ncell_done & 1));
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				uint32_t v[32];
				tmem_ld32(tmem + lane_addr + TM_S2 + (uint32_t)(h * 32), v);
#pragma unroll
				for (int j = 0; j < 32; ++j) sum[j] = __uint_as_float(v[j]);
				asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
				__syncwarp();
				if (lane == 0) mbar_arrive(s2_empty);
			}
			if (td == 0) FH_TRACE(16);
			// second-order affinity without its diagonal; S2 is symmetric, so its column sums are row sums
			float rs = 0.f;
			{
				const float rscale = m < p.nb ? s2_scale : 0.f;  // rows beyond the block are zero anyway (TMA zero fill)
#pragma unroll
				for (int j = 0; j < 32; ++j) sum[j] *= rscale;
				if (h * 32 + 32 > p.nb) {  // warp-uniform: only the quarters at / beyond the block's last column mask columns
#pragma unroll
					for (int j = 0; j < 32; ++j)
						if (h * 32 + j >= p.nb) sum[j] = 0.f;
				}
				if (diag_chunk) {
#pragma unroll
					for (int j = 0; j < 32; ++j)
						if (lane == j) sum[j] = 0.f;
				}
#pragma unroll
				for (int j = 0; j < 32; ++j) rs += sum[j];
			}
			rowsum[h * 128 + m] = rs;
			drain_sync();
			if (td == 0) FH_TRACE(17);
			if (p.bin_cov) {  // do_col: 1 / bin_cov of this cell's window columns (partial_rwr.py:135), by shifted window column.
				// After the cell's first drain barrier: no warp still reads the previous cell's table in its epilogue.
				const float* cv = p.bin_cov + (long long)cell * p.cov_ld;
				for (int i = td; i < NT * BN; i += 512) {
					const int n = i - p.pad;
					rcov[i] = (n >= 0 && n < p.w) ? 1.f / cv[n] : 0.f;
				}
			}
			// per column j: P[i][j] = (3/4 f/(cs1+eps) + 1/4 h/(cs2+eps)) / (csl+eps) = f w1[j] + h w2[j]. The column
			// sum of the blend is taken analytically, csl = 3/4 cs1/(cs1+eps) + 1/4 cs2/(cs2+eps) (the reference
			// sums the rounded entries: same value to fp32 rounding), partial_rwr.py:88-97
			if (td < 128) {
				const float c1 = cs1raw[td], c2 = (rowsum[td] + rowsum[128 + td]) + (rowsum[256 + td] + rowsum[384 + td]);
				const float r1 = 1.f / (c1 + EPS), r2 = 1.f / (c2 + EPS);
				float csl = 0.75f * (c1 * r1) + 0.25f * (c2 * r2);
				const bool empty_col = (csl == 0.f) && td < p.nb;  // unreachable after the 1e-8 floor; kept for parity
				if (empty_col) csl = 1.f;
				const float rl = 1.f / (csl + EPS);
				// pre-scaled for the binary16 operand (x 2^14; w1 also undoes the panel's scale), 0 beyond the block
				w1[td] = td < p.nb ? 0.75f * r1 * rl * (QS * sa_inv) : 0.f;
				w2[td] = td < p.nb ? 0.25f * r2 * rl * QS : 0.f;
				cflag[td] = empty_col ? rl * QS : 0.f;
			}
			drain_sync();
			if (td == 0) FH_TRACE(18);
			// ---- C: this thread's row of the first-order block -> its row of P, IN PLACE in the shared-memory B operand of the
			// chain (hi / lo halves scaled by 2^14); Q_1 = 0.5 P + 0.5 I -> TMEM. Away from the diagonal the halves of Q_1 are
			// the halves of P times 0.5 (exact), so only the warps that hold a diagonal chunk split again.
			{
				uint32_t qh[16], ql[16];
				const float cf = diag_chunk ? cflag[m] : 0.f;  // the diagonal entry of this thread's row
#pragma unroll
				for (int g = 0; g < 4; ++g) {
					const int off = ((chunk0 + g) ^ (m & 7)) * 16;
					float f[8];
					halves_to_floats(*reinterpret_cast<const uint4*>(prow + off), *reinterpret_cast<const uint4*>(prow + P_PLANE + off), f);
					const int col0 = h * 32 + 8 * g;
					const float4 wa = *reinterpret_cast<const float4*>(w1 + col0), wb = *reinterpret_cast<const float4*>(w1 + col0 + 4);
					const float4 va = *reinterpret_cast<const float4*>(w2 + col0), vb = *reinterpret_cast<const float4*>(w2 + col0 + 4);
					const float ww[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w}, vv[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
					float pv[8];
#pragma unroll
					for (int e = 0; e < 8; ++e) pv[e] = f[e] * ww[e] + sum[8 * g + e] * vv[e];  // rows >= nb: zeros
					if (diag_chunk) {
#pragma unroll
						for (int e = 0; e < 8; ++e)
							if (lane == 8 * g + e) pv[e] += cf;
					}
					uint4 vh, vl;
					f16_split2(pv[0], pv[1], vh.x, vl.x);
					f16_split2(pv[2], pv[3], vh.y, vl.y);
					f16_split2(pv[4], pv[5], vh.z, vl.z);
					f16_split2(pv[6], pv[7], vh.w, vl.w);
					if (chain) {
						*reinterpret_cast<uint4*>(prow + off) = vh;
						*reinterpret_cast<uint4*>(prow + P_PLANE + off) = vl;
					}
					if (diag_chunk) {
#pragma unroll
						for (int e = 0; e < 4; ++e) {
							const float a = 0.5f * pv[2 * e] + ((lane == 8 * g + 2 * e && m < p.nb) ? 0.5f * QS : 0.f);
							const float b = 0.5f * pv[2 * e + 1] + ((lane == 8 * g + 2 * e + 1 && m < p.nb) ? 0.5f * QS : 0.f);
							f16_split2(a, b, qh[4 * g + e], ql[4 * g + e]);
						}
					} else {
						const uint32_t ph[4] = {vh.x, vh.y, vh.z, vh.w}, pl[4] = {vl.x, vl.y, vl.z, vl.w};
						const __half2 half2c = __floats2half2_rn(0.5f, 0.5f);
#pragma unroll
						for (int e = 0; e < 4; ++e) {
							const __half2 a2 = __hmul2(*reinterpret_cast<const __half2*>(&ph[e]), half2c);
							const __half2 b2 = __hmul2(*reinterpret_cast<const __half2*>(&pl[e]), half2c);
							qh[4 * g + e] = *reinterpret_cast<const uint32_t*>(&a2);
							ql[4 * g + e] = *reinterpret_cast<const uint32_t*>(&b2);
						}
					}
				}
				tmem_st16(tmem + lane_addr + TM_QHI + (uint32_t)(h * 16), qh);
				tmem_st16(tmem + lane_addr + TM_QLO + (uint32_t)(h * 16), ql);
			}
			if (chain) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes of P -> UMMA
			publish_q();
			if (td == 0) FH_TRACE(19);
			drain_sync();
			if (td == 0) mbar_arrive(chain ? p_written : f_free);  // k = 1: nothing else reads the region
			for (int step = 1; step < p.k; ++step) {
				const int cb = (int)(ch & 1);
				mbar_wait(&acc_full[cb], (uint32_t)((ch >> 1) & 1));
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				if (td == 0 && step == 1) FH_TRACE(27);
				{
					uint32_t v[32];
					tmem_ld32(tmem + lane_addr + TM_ACC + (uint32_t)(cb * BN + h * 32), v);
					float qv[32];
#pragma unroll
					for (int j = 0; j < 32; ++j) qv[j] = (0.5f * QS_INV) * __uint_as_float(v[j]);  // Q scaled by 2^14 = 0.5 acc / 2^14 (+ 0.5 * 2^14 I)
					if (diag_chunk) {
#pragma unroll
						for (int j = 0; j < 32; ++j)
							if (lane == j && m < p.nb) qv[j] += 0.5f * QS;
					}
					if (h * 32 + 32 > n_step) {  // accumulator columns beyond the MMA's N were never written
#pragma unroll
						for (int j = 0; j < 32; ++j)
							if (h * 32 + j >= n_step) qv[j] = 0.f;
					}
					if (p.bin_cov && step == p.k - 1) {
						// do_col (partial_rwr.py:131-134): Q <- rownorm(max((Q + Q^T) / 2, 0)). The transpose goes through the P
						// region (dead: the chain's MMAs have completed) as a 128 x 128 fp32 tile, 16-byte chunk c of row r at
						// c ^ (r & 7): conflict-free row writes and conflict-free column reads.
						float* tq = reinterpret_cast<float*>(pbuf);
#pragma unroll
						for (int g = 0; g < 8; ++g)
							*reinterpret_cast<float4*>(tq + m * 128 + (((8 * h + g) ^ (m & 7)) << 2)) = make_float4(qv[4 * g], qv[4 * g + 1], qv[4 * g + 2], qv[4 * g + 3]);
						drain_sync();
						float prs = 0.f;
#pragma unroll
						for (int j = 0; j < 32; ++j) {
							const int r = h * 32 + j;  // Q^T[m][r] = Q[r][m]
							const float t = tq[r * 128 + ((((m >> 2) ^ (r & 7)) << 2) | (m & 3))];
							qv[j] = fmaxf(0.5f * (qv[j] + t), 0.f);
							prs += qv[j];
						}
						rowsum[h * 128 + m] = prs;
						drain_sync();  // every transposed read is done: the region may take the next cell's first-order block
						if (td == 0) mbar_arrive(f_free);
						const float tot = ((rowsum[m] + rowsum[128 + m]) + (rowsum[256 + m] + rowsum[384 + m])) * QS_INV + EPS;
						const float rinv = 1.f / tot;
#pragma unroll
						for (int j = 0; j < 32; ++j) qv[j] *= rinv;  // (S / 2^14) / (sum / 2^14 + eps) * 2^14
					}
					// every MMA that read the old Q completed before acc_full fired
					uint32_t hi[16], lo[16];
#pragma unroll
					for (int j = 0; j < 16; ++j) f16_split2(qv[2 * j], qv[2 * j + 1], hi[j], lo[j]);
					tmem_st16(tmem + lane_addr + TM_QHI + (uint32_t)(h * 16), hi);
					tmem_st16(tmem + lane_addr + TM_QLO + (uint32_t)(h * 16), lo);
				}
				asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
				__syncwarp();
				if (lane == 0) mbar_arrive(&acc_empty[cb]);
				++ch;
				publish_q();
				if (td == 0 && step == 1) FH_TRACE(28);
			}
			if (td == 0) FH_TRACE(20);
			for (int nt = 0; nt < NT; ++nt) {
				const int cb = (int)(ch & 1);
				mbar_wait(&acc_full[cb], (uint32_t)((ch >> 1) & 1));
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				if (td == 0) FH_TRACE(21 + nt);
				float xs[32];
				{
					uint32_t v[32];
					tmem_ld32(tmem + lane_addr + TM_ACC + (uint32_t)(cb * BN + h * 32), v);
#pragma unroll
					for (int j = 0; j < 32; ++j) xs[j] = __uint_as_float(v[j]) * x_scale;
					if (p.bin_cov) {
#pragma unroll
						for (int g = 0; g < 8; ++g) {
							const float4 rc = *reinterpret_cast<const float4*>(rcov + nt * BN + h * 32 + 4 * g);
							xs[4 * g] *= rc.x; xs[4 * g + 1] *= rc.y; xs[4 * g + 2] *= rc.z; xs[4 * g + 3] *= rc.w;
						}
					}
				}
				asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
				__syncwarp();
				if (lane == 0) mbar_arrive(&acc_empty[cb]);
				++ch;
				// each lane lays 16 columns of its row into the warp's staging buffer in the 64-byte-swizzle pattern (16-byte
				// chunk g of row r at g ^ ((r >> 1) & 3): conflict-free), one TMA store per 32 x 16 chunk; rows >= nb and
				// columns >= ldw are clipped by the tensor map
#pragma unroll
				for (int cc = 0; cc < 2; ++cc) {
					const int n0 = nt * BN + h * 32 + cc * 16 - p.pad;  // window column of the chunk
					if (n0 < 0) {
						// the first chunk of a shifted window starts at column -pad: a TMA store may not start at a negative
						// coordinate (illegal instruction, compute-sanitizer on a B200), so its 16 - pad valid columns leave
						// through 128-bit stores, one row per lane
						if (m < p.nb) {
							float* xrow = p.out + (long long)cell * p.out_cell_stride + (long long)m * p.ldw;
#pragma unroll
							for (int g = 1; g < 4; ++g)
								if (4 * g - p.pad + 4 <= p.ldw)
									*reinterpret_cast<float4*>(xrow + 4 * g - p.pad) = make_float4(xs[4 * g], xs[4 * g + 1], xs[4 * g + 2], xs[4 * g + 3]);
						}
						continue;
					}
					if (lane == 0) tma_store_wait_read();
					__syncwarp();
					float4* rowp = reinterpret_cast<float4*>(tile_s) + lane * 4;
#pragma unroll
					for (int g = 0; g < 4; ++g)
						rowp[g ^ ((lane >> 1) & 3)] = make_float4(xs[cc * 16 + 4 * g], xs[cc * 16 + 4 * g + 1], xs[cc * 16 + 4 * g + 2], xs[cc * 16 + 4 * g + 3]);
					asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
					__syncwarp();
					if (lane == 0 && n0 < p.ldw && q * 32 < p.nb) {
						tma_store_3d(&tmO, tile_s, n0, q * 32, cell);
						tma_store_commit();
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

// Ahi: two binary16 planes (hi, then lo at + ncell * a_cell_stride halves) of (ncell, nb, ld16) conv'd panels scaled as
// described at the top (written by densify_conv_kernel<.., true>), window column c at plane column c + pad with
// pad = fh_rwr_chain16_pad(s) leading zero columns; amax: [ncell] device words that fix the scales;
// bin_cov (do_col, partial_rwr.py:131-135; nullptr otherwise): [ncell][>= w] coverage of the window columns;
// out: cell c at out + c * out_cell_stride, rows of ldw floats (16-byte aligned). Returns FH_ERR_UNSUPPORTED (nothing
// launched) when the shape is outside the kernel's range - the caller runs the TF32 kernels.
int fh_rwr_chain16_pad(int s) { return (8 - (s & 7)) & 7; }

int fh_rwr_chain16(const void* Ahi, const unsigned* amax, float* out, int nb, int w, int ldw, int ld16, int s, int k,
                   int ncell, long long a_cell_stride, long long out_cell_stride, const float* bin_cov, long long bin_cov_ld,
                   void* stream) {
	if (ncell <= 0) return FH_OK;
	const int pad = fh_rwr_chain16_pad(s);
	// do_col (bin_cov != nullptr): the symmetrisation runs in the last chain step's drain (k >= 2); the reciprocal coverage
	// table holds 512 shifted window columns
	if ((bin_cov && (k < 2 || ((ldw + pad + BN - 1) / BN) * BN > 512)) || nb > BM || k < 1 || (pad & 3) || ld16 < ldw + pad || (ldw & 3) || (ld16 & 7) || (a_cell_stride & 7) || !aligned16(Ahi) || !aligned16(out) ||
	    (out_cell_stride & 3)) {
		fh_set_error("fh_rwr_chain16: shape outside the fused kernel (nb <= 128, k >= 1, 16-byte aligned rows)");
		return FH_ERR_UNSUPPORTED;
	}
	static int num_sms = 0;
	if (!num_sms) {
		int dev = 0;
		FH_CUDA(cudaGetDevice(&dev));
		FH_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
	}
	const int grid = ncell < num_sms ? ncell : num_sms;
	CUtensorMap tk, ta, to;
	// the contiguous extent is the LOGICAL width, so pad columns and rows beyond nb read as zeros whatever the buffers hold
	bool ok = make_map16(&tk, Ahi, w + pad, nb, ld16, 2LL * ncell, a_cell_stride, BK, BM) &&
	          make_map16(&ta, Ahi, w + pad, nb, ld16, 2LL * ncell, a_cell_stride, 64, BK) &&
	          make_map_sw(&to, out, ldw, nb, ldw, ncell, out_cell_stride, 16, 32, CU_TENSOR_MAP_SWIZZLE_64B);
	if (!ok) {
		fh_set_error("fh_rwr_chain16: cuTensorMapEncodeTiled failed");
		return FH_ERR_UNSUPPORTED;
	}
	Chain16P p;
	p.nb = nb; p.w = w; p.ldw = ldw; p.ld16 = ld16; p.k = k; p.ncell = ncell; p.pad = pad;
	p.a_cell_stride = a_cell_stride; p.out_cell_stride = out_cell_stride;
	p.Ahi = (const __half*)Ahi; p.amax = amax; p.s = s; p.out = out;
	p.bin_cov = bin_cov; p.cov_ld = bin_cov_ld;
	cudaStream_t st = (cudaStream_t)stream;
	static int trace_on = -1;
	if (trace_on < 0) { const char* e = getenv("FH_CHAIN_TRACE"); trace_on = (e && e[0] == '1') ? 1 : 0; }
	p.trace = nullptr;
	if (trace_on) {
		FH_CUDA(cudaMalloc(&p.trace, 32 * sizeof(long long)));
		FH_CUDA(cudaMemsetAsync(p.trace, 0, 32 * sizeof(long long), st));
	}
	static bool attr_set = false;
	if (!attr_set) {
		FH_CUDA(cudaFuncSetAttribute(rwr_chain16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
		attr_set = true;
	}
	const int tmr = fh_time_begin(FH_TIME_RWR_CHAIN, st);
	rwr_chain16_kernel<<<grid, NTHREADS, SMEM_BYTES, st>>>(tk, ta, to, p);
	fh_time_end(tmr, st);
	FH_LAUNCH_CHECK();
	if (trace_on) {  // debug only: synchronises and prints the phase stamps (SM clocks relative to stamp 8)
		long long h[32];
		FH_CUDA(cudaMemcpyAsync(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost, st));
		FH_CUDA(cudaStreamSynchronize(st));
		FH_CUDA(cudaFree(p.trace));
		long long t0 = h[8] ? h[8] : h[0];
		fprintf(stderr, "[chain16 trace] nb=%d w=%d k=%d:", nb, w, k);
		for (int i = 0; i < 32; ++i)
			if (h[i]) fprintf(stderr, " %d:%lld", i, h[i] - t0);
		fprintf(stderr, "\n");
	}
	return FH_OK;
}
