// fp32-in / fp32-out batched GEMM on the 5th-generation tensor cores (tcgen05, sm_100a) with the
// 3xTF32 split:  a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi,  a_hi = tf32(a), a_lo = tf32(a - a_hi),
// fp32 accumulation in TMEM. Used for every large contraction of the hot path (RWR: A A^T, Q P, Q A;
// sweep: X^T C, X W, X^T V) - the reference's torch.bmm / matmul call sites
// (partial_rwr.py:85,116,138; parafac2_intergrative.py:381-383,429-430,522-524).
//
// Persistent: one CTA per SM loops over 128 x 128 output tiles (tile t, t + #SMs, ...); the shared-memory
// ring and the two TMEM accumulators keep rolling across tiles, so the epilogue of one tile overlaps
// the loads / MMAs of the next. Roles inside a CTA:
//   warp 0      TMA producer: raw fp32 operand tiles (K-major or MN-major, 128-byte swizzle) from
//               HBM/L2 into a 4-stage shared-memory ring (cp.async.bulk.tensor, mbarrier tx-count)
//   warps 4-7   splitters. A: each thread owns one row of the landed tile (= one TMEM lane), un-swizzles it
//               and writes the hi (raw fp32: the tensor core truncates) and lo TF32 halves straight into
//               TENSOR MEMORY (tcgen05.st) - A is the TMEM operand of the MMAs, which takes its lo copy and
//               its three reads per k-block off shared memory (the kernel was shared-memory-bandwidth
//               bound: contractions 55 -> 48 ms per sweep). B: lo half written next to the raw tile.
//   warp 1      MMA issuer: one thread issues 3 x 4 tcgen05.mma.kind::tf32 (M=128, N=128, K=8, A from
//               TMEM) per stage; tcgen05.commit releases the stage / publishes the accumulator
//   warps 8-15  drain + epilogue. The tensor core accumulates with truncation, so a long K run drifts
//               by ~7e-9*K (measured). The accumulator is therefore double-buffered in TMEM and
//               drained every CHUNK_KB k-blocks (K=128) into fp32 registers with round-to-nearest adds
//               (tcgen05.ld), which keeps the error at the fp32 level for any K; the final sum goes
//               through alpha/beta/diag/column-scale and coalesced stores
//   warp 2      TMEM allocation (512 columns: two 128-column accumulators + 4 stages x (A_hi | A_lo))
#include "fh_tc.cuh"
#include "../../include/fh_b200.h"

namespace {
using namespace fh_tc;

constexpr int BM = 128, BN = 128, BK = 32;       // tile (BK fp32 = one 128-byte swizzle row)
constexpr int STAGES = 4;
constexpr int TILE_BYTES = BM * BK * 4;          // 16 KB per operand tile
constexpr int STAGE_BYTES = 3 * TILE_BYTES;      // A (raw fp32 as landed), B_hi (raw), B_lo
constexpr int EPI_BYTES = 8 * 32 * 32 * 4;       // epilogue transposes: 8 drain warps x (32 x 32 floats, XOR-swizzled)
constexpr uint32_t TM_A = 2 * BN;                // TMEM: accumulators [0, 2 BN), then per stage A_hi (32 cols) | A_lo (32 cols)
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int NTHREADS = 512;
constexpr int TURN_TILES = 8192, TURN_SLOTS = 64;  // split-K turn counters: tiles per launch, launches in flight
#ifndef FH_GEMM_CHUNK_KB
#define FH_GEMM_CHUNK_KB 4                       // compile-time knob for A/B builds (FH_NVCC_EXTRA, scripts/gpu_session.sh)
#endif
constexpr int CHUNK_KB = FH_GEMM_CHUNK_KB;       // k-blocks accumulated in TMEM before a drain (K = 128 at the default 4)

struct TcP {
	int M, N, K, batch;
	long long ldc, batch_c;
	float alpha, beta, diag;
	int epilogue;
	const float* cscale; long long cscale_batch; int cscale_recip;
	int a_bcast, b_bcast;  // operand shared by all batch items (batch stride 0)
	int vec_ok;            // C rows are 16-byte aligned: 128-bit epilogue accesses allowed
	int ksplit, kb_per_split;  // split-K: work item = (tile, k range); partial sums are added to C with fp32 atomics, or
	int* turn;             // (FH_GEMM_SPLITK_ORDERED=1) in the order of the ranges: one turn counter per output tile, zero between launches
	int dbg;               // FH_TC_DEBUG bits (timing experiments only, results wrong): 1 no C stores, 2 no split, 4 no MMA
};

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcP p,
               float* __restrict__ C) {
	extern __shared__ uint8_t smem_raw[];
	uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // pointer arithmetic keeps the shared address space (LDS / STS, not generic LD / ST)
	uint8_t* stagebuf = smem + STAGES * STAGE_BYTES;           // 8 warps x 32 x 32 floats (epilogue transposes)
	uint64_t* bars = (uint64_t*)(stagebuf + EPI_BYTES);
	uint64_t* raw_full = bars;                 // TMA landed            (count 1 + tx)
	uint64_t* split_full = bars + STAGES;      // hi/lo written         (count 4: one per splitter warp)
	uint64_t* empty = bars + 2 * STAGES;       // MMAs of the stage done (count 1, tcgen05.commit)
	uint64_t* acc_full = bars + 3 * STAGES;    // [2] accumulator chunk complete (count 1, tcgen05.commit)
	uint64_t* acc_empty = bars + 3 * STAGES + 2;  // [2] accumulator drained   (count 8: one per drain warp)
	uint32_t* tmem_holder = (uint32_t*)(bars + 3 * STAGES + 4);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int nkb = (p.K + BK - 1) / BK;
	// persistent: this CTA owns tiles blockIdx.x, blockIdx.x + gridDim.x, ... (n fastest, then m, then batch)
	const int tiles_n = (p.N + BN - 1) / BN, tiles_m = (p.M + BM - 1) / BM;
	const long long total_tiles = (long long)tiles_n * tiles_m * p.batch * p.ksplit;  // work items

	if (threadIdx.x == 0) {
		for (int s = 0; s < STAGES; ++s) {
			mbar_init(&raw_full[s], 1);
			mbar_init(&split_full[s], 4);
			mbar_init(&empty[s], 1);
		}
		mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
		mbar_init(&acc_empty[0], 8); mbar_init(&acc_empty[1], 8);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 0 && lane == 0) {
		asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
		asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
	}
	if (warp == 2) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(512) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem_acc = *tmem_holder;

	if (warp == 0) {
		// ------------------------------------------------------------------ TMA producer
		if (lane == 0) {
			long long it = 0;  // k-blocks issued by this CTA so far (ring position continues across tiles)
			for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
				const long long tl = tile / p.ksplit;
				const int ks = (int)(tile - tl * p.ksplit);
				const int bz = (int)(tl / ((long long)tiles_n * tiles_m));
				const int rem = (int)(tl - (long long)bz * tiles_n * tiles_m);
				const int m0 = (rem / tiles_n) * BM, n0 = (rem % tiles_n) * BN;
				const int za = p.a_bcast ? 0 : bz, zb = p.b_bcast ? 0 : bz;
				const int kb0 = ks * p.kb_per_split, kb1 = min(nkb, kb0 + p.kb_per_split);
				for (int kb = kb0; kb < kb1; ++kb, ++it) {
					const int s = (int)(it % STAGES);
					const uint32_t ph = (uint32_t)((it / STAGES) & 1);
					mbar_wait(&empty[s], ph ^ 1);
					uint8_t* st = smem + s * STAGE_BYTES;
					mbar_expect_tx(&raw_full[s], 2 * TILE_BYTES);
					const int k0 = kb * BK;
					if (A_MN) {  // global (K rows x M contiguous): four 32-wide boxes of BK rows
#pragma unroll
						for (int j = 0; j < BM / 32; ++j) tma_load_3d(st + j * (BK * 128), &tmA, &raw_full[s], m0 + 32 * j, k0, za);
					} else {     // global (M rows x K contiguous): one box of 32 x 128
						tma_load_3d(st, &tmA, &raw_full[s], k0, m0, za);
					}
					uint8_t* sb = st + TILE_BYTES;
					if (B_MN) {
#pragma unroll
						for (int j = 0; j < BN / 32; ++j) tma_load_3d(sb + j * (BK * 128), &tmB, &raw_full[s], n0 + 32 * j, k0, zb);
					} else {
						tma_load_3d(sb, &tmB, &raw_full[s], k0, n0, zb);
					}
				}
			}
		}
	} else if (warp == 1) {
		// ------------------------------------------------------------------ MMA issuer
		if (lane == 0) {
			// cute::UMMA::InstrDescriptor: c_format F32 [4,6)=1, a/b_format TF32 [7,10)/[10,13)=2,
			// a_major [15], b_major [16], N>>3 [17,23), M>>4 [24,29)
			// A comes from TMEM (always "K-major"). N of the instruction = the tile's useful columns rounded up to 32 (a whole
			// 32-column group of the MN-major B layout): the last column tile of a contraction whose N is just above a multiple
			// of 128 (P1: N = r = 137..149 -> tiles of 128 + 32) no longer pays 128 columns of tensor-pipe time. The columns
			// beyond it keep stale accumulator contents; the epilogue never stores columns >= N.
			const uint32_t idesc_full = make_idesc_tf32(false, B_MN, BN, BM);
			// K-major B tile (SWIZZLE_128B): rows of 128 B, 8-row groups 1024 B apart (SBO); a K=8 step = +32 B.
			// MN-major B tile (SWIZZLE_128B_BASE32B): k rows of 128 B (32 N elements), 4-row atoms 512 B apart
			// (SBO), 32-element N groups BK*128 B apart (LBO); a K=8 step = 8 rows = +1024 B.
			const uint32_t b_lbo = B_MN ? BK * 128 : 16, b_sbo = B_MN ? 512 : 1024, b_step = B_MN ? 1024 : 32, b_lt = B_MN ? 1 : 2;
			long long it = 0, ch = 0;  // k-blocks / accumulator chunks consumed by this CTA so far
			for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
				const int ksm = (int)(tile % p.ksplit);
				const int nkb_t = min(nkb, (ksm + 1) * p.kb_per_split) - ksm * p.kb_per_split;
				const int n0t = (int)((tile / p.ksplit) % tiles_n) * BN;
				const int n_eff = min(BN, (p.N - n0t + 31) & ~31);
				const uint32_t idesc = n_eff == BN ? idesc_full : make_idesc_tf32(false, B_MN, n_eff, BM);
				for (int kb = 0; kb < nkb_t; ++kb, ++it) {
					const int s = (int)(it % STAGES);
					const uint32_t ph = (uint32_t)((it / STAGES) & 1);
					const int cb = (int)(ch & 1);
					if (kb % CHUNK_KB == 0) {  // new chunk: its accumulator must have been drained
						mbar_wait(&acc_empty[cb], (uint32_t)(((ch >> 1) & 1) ^ 1));
						asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
					}
					const uint32_t acc = tmem_acc + (uint32_t)(cb * BN);
					mbar_wait(&split_full[s], ph);
					asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
					const uint32_t b_hi = smem_u32(smem + s * STAGE_BYTES) + TILE_BYTES, b_lo = b_hi + TILE_BYTES;
					const uint32_t ta = tmem_acc + TM_A + (uint32_t)(s * 64);
#pragma unroll
					for (int k = 0; k < BK / 8; ++k) {
						if (p.dbg & 4) break;
						const uint32_t a_hi = ta + (uint32_t)(k * 8), a_lo = a_hi + 32;
						const uint64_t dbh = make_desc(b_hi + k * b_step, b_lbo, b_sbo, b_lt), dbl = make_desc(b_lo + k * b_step, b_lbo, b_sbo, b_lt);
						umma_tf32_ts(acc, a_lo, dbh, idesc, ((kb % CHUNK_KB) | k) ? 1u : 0u);  // small terms first
						umma_tf32_ts(acc, a_hi, dbl, idesc, 1u);
						umma_tf32_ts(acc, a_hi, dbh, idesc, 1u);
					}
					umma_commit(&empty[s]);
					if (kb % CHUNK_KB == CHUNK_KB - 1 || kb == nkb_t - 1) {
						umma_commit(&acc_full[cb]);
						++ch;
					}
				}
			}
		}
	} else if (warp >= 4 && warp < 8) {
		// ------------------------------------------------------------------ splitters
		const int t = threadIdx.x - 128;  // 0..127
		long long it = 0;
		for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
			const int ksm = (int)(tile % p.ksplit);
			const int nkb_t = min(nkb, (ksm + 1) * p.kb_per_split) - ksm * p.kb_per_split;
			for (int kb = 0; kb < nkb_t; ++kb, ++it) {
				const int s = (int)(it % STAGES);
				const uint32_t ph = (uint32_t)((it / STAGES) & 1);
				mbar_wait(&raw_full[s], ph);
				uint8_t* st = smem + s * STAGE_BYTES;
				if (!(p.dbg & 2)) {
					// A: thread t owns row m = t of the tile (= TMEM lane t). The raw fp32 row is the hi operand (the
					// tensor core truncates to tf32), lo = tf32(a - trunc(a)); both go straight to TENSOR MEMORY
					// (tcgen05.st) and are the A operand of the MMAs - the kernel is shared-memory-bandwidth bound and
					// this removes the A_lo write and the three A reads per k-block from shared memory.
					uint32_t hi[32], lo[32];
					if (A_MN) {
						// MN-major landing (SWIZZLE_128B_ATOM_32B): 32-row m groups BK*128 B apart; k row = 128 B holding
						// 32 m elements, its 32-byte chunks XOR-ed with (k & 3)
						const uint8_t* g = st + (t >> 5) * (BK * 128);
						const int ml = t & 31;
#pragma unroll
						for (int k = 0; k < 32; ++k)
							hi[k] = *reinterpret_cast<const uint32_t*>(g + k * 128 + ((((ml >> 3) ^ (k & 3)) << 5) | ((ml & 7) << 2)));
					} else {
						// K-major landing (SWIZZLE_128B): row m = 128 B holding 32 k elements, 16-byte chunks XOR-ed with (m & 7)
						const uint4* g = reinterpret_cast<const uint4*>(st + t * 128);
#pragma unroll
						for (int c = 0; c < 8; ++c) {
							const uint4 v = g[c ^ (t & 7)];
							hi[4 * c] = v.x; hi[4 * c + 1] = v.y; hi[4 * c + 2] = v.z; hi[4 * c + 3] = v.w;
						}
					}
#pragma unroll
					for (int k = 0; k < 32; ++k) lo[k] = tf32_lo(__uint_as_float(hi[k]));
					const uint32_t ta = tmem_acc + ((uint32_t)((warp & 3) * 32) << 16) + TM_A + (uint32_t)(s * 64);
					tmem_st32(ta, hi);
					tmem_st32(ta + 32, lo);
					// B: lo half next to the raw tile in shared memory
					const float4* bh = (const float4*)(st + TILE_BYTES);
					uint4* bl = (uint4*)(st + 2 * TILE_BYTES);
#pragma unroll
					for (int i = 0; i < TILE_BYTES / 16 / 128; ++i) {
						const float4 v = bh[t + 128 * i];
						uint4 l;
						l.x = tf32_lo(v.x); l.y = tf32_lo(v.y); l.z = tf32_lo(v.z); l.w = tf32_lo(v.w);
						bl[t + 128 * i] = l;
					}
					tmem_st_wait();
					asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
				}
				asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (UMMA)
				__syncwarp();
				if (lane == 0) mbar_arrive(&split_full[s]);
			}
		}
	} else if (warp >= 8) {
		// ------------------------------------------------------------------ drain + epilogue
		const int q = warp & 3;             // TMEM lane quarter of this warp (rows 32q .. 32q+31)
		const int h = (warp - 8) >> 2;      // column half (64 columns)
		float* tile_s = (float*)(stagebuf + (warp - 8) * (32 * 32 * 4));
		long long ch = 0;
		for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
			const long long tl = tile / p.ksplit;
			const int ksm = (int)(tile - tl * p.ksplit);
			const int nkb_t = min(nkb, (ksm + 1) * p.kb_per_split) - ksm * p.kb_per_split;
			const int nchunk_t = (nkb_t + CHUNK_KB - 1) / CHUNK_KB;
			const int bz = (int)(tl / ((long long)tiles_n * tiles_m));
			const int rem = (int)(tl - (long long)bz * tiles_n * tiles_m);
			const int m0 = (rem / tiles_n) * BM, n0 = (rem % tiles_n) * BN;
			float sum[64];
#pragma unroll
			for (int j = 0; j < 64; ++j) sum[j] = 0.f;
			for (int chunk = 0; chunk < nchunk_t; ++chunk, ++ch) {
				const int cb = (int)(ch & 1);
				mbar_wait(&acc_full[cb], (uint32_t)((ch >> 1) & 1));
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
				for (int c = 0; c < 2; ++c) {
					uint32_t v[32];
					const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(cb * BN + h * 64 + c * 32);
					asm volatile(
						"tcgen05.ld.sync.aligned.32x32b.x32.b32 "
						"{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
						"%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
						: "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
						  "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
						  "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
						  "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
						: "r"(taddr));
					asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
					for (int j = 0; j < 32; ++j) sum[c * 32 + j] += __uint_as_float(v[j]);
				}
				asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
				__syncwarp();
				if (lane == 0) mbar_arrive(&acc_empty[cb]);
			}
			// epilogue of this tile; the MMA warp is already free to start the next tile's chunks.
			// A 32 x 32 transpose through a private staging buffer, then each lane owns 4 consecutive
			// columns of 8 rows: 128-bit stores, every warp store instruction covers four 128-byte rows.
			float* Cb = C + (long long)bz * p.batch_c;
			const float* cs = p.cscale ? p.cscale + (long long)bz * p.cscale_batch : nullptr;
			const int cg = lane & 7, ro = lane >> 3;
			// split-K: the k ranges of a tile add their partial sums to C in the order of the ranges (range 0 stores
			// beta C + x, range s waits for range s - 1): a fixed summation order, bit-reproducible, no atomics and no
			// zero-fill. The ranges of a tile are consecutive work items, i.e. they run on neighbouring CTAs of the same
			// wave (all CTAs of the persistent grid are resident, and a range only ever waits for a lower work item, so the
			// wait cannot deadlock); only the drain warps wait - the CTA's loads and MMAs for its next item go on.
			float beta_t = p.beta;
			const bool ordered = p.ksplit > 1 && p.turn != nullptr;  // else FH_GEMM_SPLITK_ORDERED=0: fp32 atomics onto beta C
			if (ordered) {
				if (ksm > 0) {
					beta_t = 1.f;
					if (lane == 0) {
						const volatile int* tp = p.turn + tl;
						while (*tp != ksm) __nanosleep(64);
					}
					__syncwarp();
					__threadfence();  // acquire: the previous range's stores to C are visible
				}
			}
#pragma unroll
			for (int c = 0; c < 2; ++c) {
				__syncwarp();
#pragma unroll
				for (int j = 0; j < 32; ++j) tile_s[lane * 32 + (j ^ lane)] = sum[c * 32 + j];  // word (r, j) at r*32 + (j ^ r)
				__syncwarp();
				const int n = n0 + h * 64 + c * 32 + 4 * cg;  // first of this lane's 4 columns
				if (n < p.N && !(p.dbg & 1)) {
					float csv[4] = {1.f, 1.f, 1.f, 1.f};
					if (cs) {
#pragma unroll
						for (int e = 0; e < 4; ++e)
							if (n + e < p.N) csv[e] = cs[n + e];
					}
					const bool vec = p.vec_ok && (n + 3 < p.N);
					// rolled on purpose: ncu showed the drain warps stalled on instruction fetch (stall_no_inst)
					// when this was fully unrolled - the kernel must stay inside the instruction cache
#pragma unroll 1
					for (int itr = 0; itr < 8; ++itr) {
						const int rr = itr * 4 + ro;
						const int r = m0 + q * 32 + rr;
						if (r >= p.M) continue;
						float x[4];
#pragma unroll
						for (int e = 0; e < 4; ++e) {
							x[e] = p.alpha * tile_s[rr * 32 + ((4 * cg + e) ^ rr)];
							if (p.epilogue == FH_EPI_DIAG_ADD && r == n + e) x[e] += p.diag;
							if (cs) x[e] = p.cscale_recip ? x[e] / csv[e] : x[e] * csv[e];
						}
						float* cp = Cb + (long long)r * p.ldc + n;
						if (p.ksplit > 1 && !ordered) {
#pragma unroll
							for (int e = 0; e < 4; ++e)
								if (n + e < p.N) atomicAdd(cp + e, x[e]);
						} else if (vec) {
							if (beta_t != 0.f) {
								const float4 o = *reinterpret_cast<const float4*>(cp);
								x[0] += beta_t * o.x; x[1] += beta_t * o.y; x[2] += beta_t * o.z; x[3] += beta_t * o.w;
							}
							*reinterpret_cast<float4*>(cp) = make_float4(x[0], x[1], x[2], x[3]);
						} else {
#pragma unroll
							for (int e = 0; e < 4; ++e)
								if (n + e < p.N) {
									if (beta_t != 0.f) x[e] += beta_t * cp[e];
									cp[e] = x[e];
								}
						}
					}
				}
			}
			if (ordered) {
				__threadfence();  // release: this range's stores to C before the turn moves on
				asm volatile("bar.sync 2, 256;" ::: "memory");  // all eight drain warps have stored their part
				if (threadIdx.x == 256) p.turn[tl] = (ksm + 1 == p.ksplit) ? 0 : ksm + 1;  // the last range leaves the counter at zero
			}
		}
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 2) {
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(512) : "memory");
	}
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

// returns FH_OK, or FH_ERR_UNSUPPORTED when the operands cannot be described to TMA (caller decides)
int fh_gemm_tc(const fh_gemm_desc* d, const float* A, const float* B, float* C, void* stream) {
	if (d->M <= 0 || d->N <= 0 || d->batch <= 0) return FH_OK;
	const bool a_mn = d->sa_m == 1 && d->sa_k != 1, b_mn = d->sb_n == 1 && d->sb_k != 1;
	const long long a_row = a_mn ? d->sa_k : d->sa_m, b_row = b_mn ? d->sb_k : d->sb_n;
	const bool a_bc = d->batch > 1 && d->batch_a == 0, b_bc = d->batch > 1 && d->batch_b == 0;
	if (d->kscale || d->K <= 0 || (a_row & 3) || (b_row & 3) || !aligned16(A) || !aligned16(B) ||
	    (d->batch > 1 && !a_bc && (d->batch_a & 3)) || (d->batch > 1 && !b_bc && (d->batch_b & 3)) ||
	    (d->sa_m == 1 && d->sa_k == 1) || (d->sb_k == 1 && d->sb_n == 1) || d->batch > 65535) {
		fh_set_error("fh_gemm_tc: operands not TMA-describable (strides must be multiples of 4 floats, bases 16-byte aligned)");
		return FH_ERR_UNSUPPORTED;
	}
	CUtensorMap ta, tb;
	bool ok;
	if (a_mn) ok = make_map(&ta, A, d->M, d->K, a_row, a_bc ? 1 : d->batch, d->batch_a, 32, BK, true);
	else ok = make_map(&ta, A, d->K, d->M, a_row, a_bc ? 1 : d->batch, d->batch_a, BK, BM, false);
	if (ok) {
		if (b_mn) ok = make_map(&tb, B, d->N, d->K, b_row, b_bc ? 1 : d->batch, d->batch_b, 32, BK, true);
		else ok = make_map(&tb, B, d->K, d->N, b_row, b_bc ? 1 : d->batch, d->batch_b, BK, BN, false);
	}
	if (!ok) {
		fh_set_error("fh_gemm_tc: cuTensorMapEncodeTiled failed");
		return FH_ERR_UNSUPPORTED;
	}
	TcP p;
	p.M = d->M; p.N = d->N; p.K = d->K; p.batch = d->batch;
	p.ldc = d->ldc; p.batch_c = d->batch_c;
	p.alpha = (float)d->alpha; p.beta = (float)d->beta; p.diag = (float)d->diag; p.epilogue = d->epilogue;
	p.cscale = d->cscale; p.cscale_batch = d->cscale_batch; p.cscale_recip = d->cscale_recip;
	p.a_bcast = a_bc; p.b_bcast = b_bc;
	p.vec_ok = aligned16(C) && (d->ldc % 4 == 0) && (d->batch <= 1 || d->batch_c % 4 == 0);
	const long long total_tiles = (long long)fh_cdiv(d->N, BN) * fh_cdiv(d->M, BM) * d->batch;
	static int num_sms = 0;
	if (!num_sms) {
		int dev = 0;
		FH_CUDA(cudaGetDevice(&dev));
		FH_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
	}
	// long-K problems with few output tiles (P3: cells x 256 outputs, K = nb*ldw ~ 36k): split K so every SM
	// has work; partial sums are added with fp32 atomics (C pre-scaled by beta here) or, on request, in the order of the k ranges
	static int dbg = -1;
	if (dbg < 0) { const char* e = getenv("FH_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
	p.dbg = dbg;
	static int ordered_splitk = -1;
	// FH_GEMM_SPLITK_ORDERED=1: bit-reproducible split-K (the k ranges of a tile take turns on C); default 0: fp32 atomics
	// (measured at config 2: the ordered mode costs 2.5 ms of a 134 ms sweep - P5 +1.4, P3 +0.7, CP-ALS +0.4)
	if (ordered_splitk < 0) { const char* e = getenv("FH_GEMM_SPLITK_ORDERED"); ordered_splitk = e ? atoi(e) : 0; }
	const int nkb_h = fh_cdiv(d->K, BK);
	p.ksplit = 1; p.kb_per_split = nkb_h; p.turn = nullptr;
	// ... and, for any long-K problem, so that the work items fill whole waves of the persistent grid: 196 output tiles
	// on 148 SMs (P3 at 12,500 cells) are two waves of which the second is a third full; split 3 ways they are four full ones
	if (nkb_h >= 64 && !d->cscale && d->epilogue == FH_EPI_NONE && (d->beta == 0.0 || d->beta == 1.0)) {
		int ks = 1;
		double best = (double)fh_cdiv(total_tiles, num_sms);  // waves, in units of one unsplit tile
		for (int c = 2; c <= 8 && nkb_h / c >= 32; ++c) {
			const double cost = (double)fh_cdiv(total_tiles * c, num_sms) / c;
			if (cost < best * 0.93) { best = cost; ks = c; }  // the ordered read-modify-write of C must be paid for
		}
		if (total_tiles * 2 <= num_sms + 12) {  // few tiles: as before, one work item per SM at least
			int k2 = (int)(num_sms / total_tiles);
			if (k2 > nkb_h / 32) k2 = nkb_h / 32;
			if (ordered_splitk && k2 > 8) k2 = 8;  // the ranges of a tile take turns on C: many short ranges would queue behind
			                                       // each other (the CP-ALS products of this shape run on up to 11 streams anyway)
			if (k2 > ks) ks = k2;
		}
		if (ks > 1 && !ordered_splitk) {  // FH_GEMM_SPLITK_ORDERED=0: zero-fill (beta = 0) + fp32 atomics, any summation order
			p.kb_per_split = (fh_cdiv(nkb_h, ks) + CHUNK_KB - 1) / CHUNK_KB * CHUNK_KB;
			p.ksplit = fh_cdiv(nkb_h, p.kb_per_split);
			if (d->beta == 0.0)
				for (int b = 0; b < d->batch; ++b)
					FH_CUDA(cudaMemset2DAsync(C + (long long)b * d->batch_c, (size_t)d->ldc * 4, 0, (size_t)d->N * 4, (size_t)d->M, (cudaStream_t)stream));
			p.beta = 0.f;
		} else if (ks > 1 && total_tiles <= TURN_TILES) {
			p.kb_per_split = (fh_cdiv(nkb_h, ks) + CHUNK_KB - 1) / CHUNK_KB * CHUNK_KB;
			p.ksplit = fh_cdiv(nkb_h, p.kb_per_split);
			// turn counters: TURN_SLOTS regions handed out round robin, so that split-K launches in flight on different
			// streams (CP-ALS runs on up to 11) never share one; every launch leaves its counters at zero
			static int* turn_base = nullptr;
			static unsigned turn_seq = 0;
			if (!turn_base) {
				FH_CUDA(cudaMalloc(&turn_base, (size_t)TURN_SLOTS * TURN_TILES * sizeof(int)));
				FH_CUDA(cudaMemset(turn_base, 0, (size_t)TURN_SLOTS * TURN_TILES * sizeof(int)));
			}
			p.turn = turn_base + (size_t)(turn_seq++ % TURN_SLOTS) * TURN_TILES;
		}
	}
	const long long work = total_tiles * p.ksplit;
	dim3 grid((unsigned)(work < num_sms ? work : num_sms));  // persistent: one CTA per SM
	cudaStream_t st = (cudaStream_t)stream;
#define FH_TC_LAUNCH(AM, BMN)                                                                                         \
	do {                                                                                                              \
		FH_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<AM, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES)); \
		gemm_tc_kernel<AM, BMN><<<grid, NTHREADS, SMEM_BYTES, st>>>(ta, tb, p, C);                                    \
	} while (0)
	const int tmr = fh_time_begin(FH_TIME_GEMM_TC, st);
	if (a_mn && b_mn) FH_TC_LAUNCH(true, true);
	else if (a_mn) FH_TC_LAUNCH(true, false);
	else if (b_mn) FH_TC_LAUNCH(false, true);
	else FH_TC_LAUNCH(false, false);
#undef FH_TC_LAUNCH
	fh_time_end(tmr, st);
	FH_LAUNCH_CHECK();
	return FH_OK;
}
