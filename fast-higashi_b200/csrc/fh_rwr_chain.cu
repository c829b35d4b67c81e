// Fused tail of the RWR imputation of one bin block (reference: partial_rwr.py:99-126 and :138):
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
constexpr int SLOTS = 6;
constexpr int TILE_BYTES = BK * BN * 4;         // 16 KB: 32 k-rows x 128 n (four 32-wide TMA boxes)
constexpr int SLOT_BYTES = 2 * TILE_BYTES;      // hi (raw fp32 as landed), lo
constexpr int EPI_BYTES = 8 * 32 * 32 * 4;      // 8 drain warps x (32 x 32 floats, XOR-swizzled)
constexpr int SMEM_BYTES = SLOTS * SLOT_BYTES + EPI_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int NTHREADS = 512;
constexpr uint32_t TM_QHI = 0, TM_QLO = 128, TM_ACC = 256;

struct ChainP {
	int nb, w, ldw, ldp, k, ncell;
	long long p_cell_stride, out_cell_stride;
	const float* P;
	float* out;
	int vec_ok;
};

__global__ void __launch_bounds__(NTHREADS, 1)
rwr_chain_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmA, ChainP p) {
	extern __shared__ uint8_t smem_raw[];
	uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
	uint8_t* stagebuf = smem + SLOTS * SLOT_BYTES;
	uint64_t* bars = (uint64_t*)(stagebuf + EPI_BYTES);
	uint64_t* raw_full = bars;                    // TMA landed              (count 1 + tx)
	uint64_t* split_full = bars + SLOTS;          // lo half written         (count 4: splitter warps)
	uint64_t* empty = bars + 2 * SLOTS;           // MMAs reading the slot done (tcgen05.commit)
	uint64_t* acc_full = bars + 3 * SLOTS;        // [2] accumulator complete (tcgen05.commit)
	uint64_t* acc_empty = bars + 3 * SLOTS + 2;   // [2] accumulator drained  (count 8: drain warps)
	uint64_t* q_ready = bars + 3 * SLOTS + 4;     // Q hi/lo stored in TMEM   (count 8: drain warps)
	uint32_t* tmem_holder = (uint32_t*)(bars + 3 * SLOTS + 5);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int nkb = (p.nb + BK - 1) / BK;         // k-blocks of the bin dimension (K of every product)
	const int NT = (p.ldw + BN - 1) / BN;         // 128-column tiles of the window
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
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 0 && lane == 0) {
		asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmP) : "memory");
		asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
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
			long long it = 0;
			for (int cell = blockIdx.x; cell < p.ncell; cell += gridDim.x) {
				const int nslots = (chain ? nkb : 0) + NT * nkb;
				for (int j = 0; j < nslots; ++j, ++it) {
					const int s = (int)(it % SLOTS);
					mbar_wait(&empty[s], (uint32_t)(((it / SLOTS) & 1) ^ 1));
					uint8_t* dst = smem + s * SLOT_BYTES;
					mbar_expect_tx(&raw_full[s], TILE_BYTES);
					const CUtensorMap* tm;
					int n0, k0;
					if (chain && j < nkb) { tm = &tmP; n0 = 0; k0 = j * BK; }
					else {
						const int t = j - (chain ? nkb : 0);
						tm = &tmA; n0 = (t / nkb) * BN; k0 = (t % nkb) * BK;
					}
#pragma unroll
					for (int b = 0; b < BN / 32; ++b) tma_load_3d(dst + b * (BK * 128), tm, &raw_full[s], n0 + 32 * b, k0, cell);
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
			auto issue_kb = [&](uint32_t acc, int kb, uint32_t slot_addr) {
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
			for (int cell = blockIdx.x; cell < p.ncell; cell += gridDim.x) {
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
							issue_kb(acc, kb, smem_u32(smem + s * SLOT_BYTES));
						}
						umma_commit(&acc_full[cb]);
						++ch;
					}
					for (int kb = 0; kb < nkb; ++kb) umma_commit(&empty[(int)((p_it + kb) % SLOTS)]);  // P no longer needed
					it += nkb;
				}
				mbar_wait(q_ready, (uint32_t)(qn & 1)); ++qn;   // Q_k
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				for (int nt = 0; nt < NT; ++nt) {
					const int cb = (int)(ch & 1);
					mbar_wait(&acc_empty[cb], (uint32_t)(((ch >> 1) & 1) ^ 1));
					asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
					const uint32_t acc = tmem + TM_ACC + (uint32_t)(cb * BN);
					for (int kb = 0; kb < nkb; ++kb, ++it) {
						const int s = (int)(it % SLOTS);
						mbar_wait(&split_full[s], (uint32_t)((it / SLOTS) & 1));
						asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
						issue_kb(acc, kb, smem_u32(smem + s * SLOT_BYTES));
						umma_commit(&empty[s]);
					}
					umma_commit(&acc_full[cb]);
					++ch;
				}
			}
		}
	} else if (warp >= 4 && warp < 8) {
		// ------------------------------------------------------------------ splitters
		const int t = threadIdx.x - 128;  // 0..127
		long long it = 0;
		for (int cell = blockIdx.x; cell < p.ncell; cell += gridDim.x) {
			const int nslots = (chain ? nkb : 0) + NT * nkb;
			for (int j = 0; j < nslots; ++j, ++it) {
				const int s = (int)(it % SLOTS);
				mbar_wait(&raw_full[s], (uint32_t)((it / SLOTS) & 1));
				const float4* hi = (const float4*)(smem + s * SLOT_BYTES);
				uint4* lo = (uint4*)(smem + s * SLOT_BYTES + TILE_BYTES);
#pragma unroll
				for (int i = 0; i < TILE_BYTES / 16 / 128; ++i) {
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
		// ------------------------------------------------------------------ drain: accumulator -> Q / X
		const int q = warp & 3;             // TMEM lane quarter of this warp (rows 32q .. 32q+31)
		const int h = (warp - 8) >> 2;      // column half (64 columns)
		const int m = q * 32 + lane;        // this thread's row
		const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
		float* tile_s = (float*)(stagebuf + (warp - 8) * (32 * 32 * 4));
		// Q (this thread's row, this warp's 64 columns) -> TMEM hi / lo, then publish
		auto store_q = [&](const float (&qv)[64]) {
#pragma unroll
			for (int c = 0; c < 2; ++c) {
				uint32_t hi[32], lo[32];
#pragma unroll
				for (int j = 0; j < 32; ++j) { hi[j] = __float_as_uint(qv[c * 32 + j]); lo[j] = tf32_lo(qv[c * 32 + j]); }
				tmem_st32(tmem + lane_addr + TM_QHI + (uint32_t)(h * 64 + c * 32), hi);
				tmem_st32(tmem + lane_addr + TM_QLO + (uint32_t)(h * 64 + c * 32), lo);
			}
			tmem_st_wait();
			asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
			__syncwarp();
			if (lane == 0) mbar_arrive(q_ready);
		};
		// Q_1 = 0.5 P + 0.5 I from the transition matrix in global memory (first RWR step, Q_0 = I)
		auto first_q = [&](int cell) {
			float qv[64];
			const float* prow = p.P + (long long)cell * p.p_cell_stride + (long long)m * p.ldp + h * 64;
#pragma unroll
			for (int g = 0; g < 16; ++g) {
				const int col = h * 64 + 4 * g;
				float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
				if (m < p.nb && col < p.ldp) v = *reinterpret_cast<const float4*>(prow + 4 * g);
				const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
				for (int x = 0; x < 4; ++x) {
					float r = 0.f;
					if (m < p.nb && col + x < p.nb) r = 0.5f * e[x] + ((m == col + x) ? 0.5f : 0.f);
					qv[4 * g + x] = r;
				}
			}
			store_q(qv);
		};
		long long ch = 0;
		if ((int)blockIdx.x < p.ncell) first_q(blockIdx.x);
		for (int cell = blockIdx.x; cell < p.ncell; cell += gridDim.x) {
			for (int step = 1; step < p.k; ++step) {
				const int cb = (int)(ch & 1);
				mbar_wait(&acc_full[cb], (uint32_t)((ch >> 1) & 1));
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				float qv[64];
#pragma unroll
				for (int c = 0; c < 2; ++c) {
					uint32_t v[32];
					tmem_ld32(tmem + lane_addr + TM_ACC + (uint32_t)(cb * BN + h * 64 + c * 32), v);
#pragma unroll
					for (int j = 0; j < 32; ++j) {
						float r = 0.5f * __uint_as_float(v[j]);
						if (m == h * 64 + c * 32 + j && m < p.nb) r += 0.5f;
						qv[c * 32 + j] = r;
					}
				}
				asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
				__syncwarp();
				if (lane == 0) mbar_arrive(&acc_empty[cb]);
				++ch;
				store_q(qv);  // every MMA that read the old Q completed before acc_full fired
			}
			float* ob = p.out + (long long)cell * p.out_cell_stride;
			for (int nt = 0; nt < NT; ++nt) {
				const int cb = (int)(ch & 1);
				mbar_wait(&acc_full[cb], (uint32_t)((ch >> 1) & 1));
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				// the last tile's completion frees Q: hand the next cell's Q_1 to the MMA warp first (its
				// first accumulator is the other buffer), then drain this tile
				if (nt == NT - 1 && cell + (int)gridDim.x < p.ncell) first_q(cell + gridDim.x);
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
				// 32 x 32 transposes through an XOR-swizzled private buffer (word (r, j) at r*32 + (j ^ r)), then
				// each lane owns 4 consecutive columns of 8 rows: 128-bit stores, four 128-byte rows per instruction
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
		}
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 2) {
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
	}
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

// P: (ncell, nb, ldp) column-stochastic transition matrices; A: (ncell, nb, ldw) conv'd panels;
// out: cell c at out + c * out_cell_stride, rows of ldw floats. Returns FH_ERR_UNSUPPORTED (nothing
// launched) when the shape is outside the fused kernel's range - the caller runs the GEMM chain.
int fh_rwr_chain(const float* P, const float* A, float* out, int nb, int w, int ldw, int ldp, int k, int ncell,
                 long long p_cell_stride, long long a_cell_stride, long long out_cell_stride, void* stream) {
	if (ncell <= 0) return FH_OK;
	if (nb > BM || k < 1 || (ldp & 3) || (ldw & 3) || (p_cell_stride & 3) || (a_cell_stride & 3) || !aligned16(P) ||
	    !aligned16(A) || ncell > 65535) {
		fh_set_error("fh_rwr_chain: shape outside the fused kernel (nb <= 128, k >= 1, 16-byte aligned rows)");
		return FH_ERR_UNSUPPORTED;
	}
	CUtensorMap tp, ta;
	// (k rows x n contiguous) MN-major operands; the contiguous extent is the LOGICAL width, so pad columns and
	// rows beyond nb read as zeros whatever the buffers hold
	if (!make_map(&tp, P, nb, nb, ldp, ncell, p_cell_stride, 32, BK, true) ||
	    !make_map(&ta, A, w, nb, ldw, ncell, a_cell_stride, 32, BK, true)) {
		fh_set_error("fh_rwr_chain: cuTensorMapEncodeTiled failed");
		return FH_ERR_UNSUPPORTED;
	}
	ChainP p;
	p.nb = nb; p.w = w; p.ldw = ldw; p.ldp = ldp; p.k = k; p.ncell = ncell;
	p.p_cell_stride = p_cell_stride; p.out_cell_stride = out_cell_stride;
	p.P = P; p.out = out;
	p.vec_ok = aligned16(out) && (out_cell_stride % 4 == 0);
	static int num_sms = 0;
	if (!num_sms) {
		int dev = 0;
		FH_CUDA(cudaGetDevice(&dev));
		FH_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
	}
	FH_CUDA(cudaFuncSetAttribute(rwr_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
	rwr_chain_kernel<<<ncell < num_sms ? ncell : num_sms, NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(tp, ta, p);
	FH_LAUNCH_CHECK();
	return FH_OK;
}
