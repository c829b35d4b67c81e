// tcgen05 / TMEM / TMA building blocks shared by the tensor-core kernels (fh_gemm_tc.cu,
// fh_rwr_chain.cu): mbarrier and TMA wrappers, UMMA shared-memory / instruction descriptors for TF32
// operands, TMEM load/store, and the host-side tensor-map encoder. sm_100a only.
#pragma once
#include <cuda.h>
#include "fh_common.cuh"

namespace fh_tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"LAB_WAIT:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra LAB_DONE;\n"
		"bra LAB_WAIT;\n"
		"LAB_DONE:\n"
		"}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
	asm volatile(
		"cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
			smem_u32(dst)),
		"l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
		: "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
	// cute::UMMA::SmemDescriptor: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) |
	// layout type [61,64): SWIZZLE_128B = 2 (K-major), SWIZZLE_128B_BASE32B = 1 (the only MN-major
	// layout tf32 operands have: 32-byte chunks swizzled inside 128-byte rows, 4-row atoms)
	uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFF);
	d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
	d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
	d |= 1ull << 46;
	d |= (uint64_t)layout_type << 61;
	return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"setp.ne.b32 p, %4, 0;\n"
		"tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
		"}\n" ::"r"(tmem_d),
		"l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
		: "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t tf32_rn(uint32_t bits) { return (bits + 0x1000u) & 0xFFFFE000u; }

// shared -> global tile store (bulk async group), out-of-range rows / columns clipped by the tensor map
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
	asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"((uint64_t)map),
	             "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
	             : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// pull a tile into L2 ahead of its TMA load
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
	asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"((uint64_t)map), "r"(c0), "r"(c1), "r"(c2)
	             : "memory");
}
// UMMA instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 [4,6)=1, a/b_format TF32
// [7,10)/[10,13)=2, a_major [15], b_major [16] (1 = MN-major), N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t make_idesc_tf32(bool a_mn, bool b_mn, int n, int m) {
	return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
	       ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M=128 rows = TMEM lanes, one tf32 per column) is read
// from tensor memory (cute SM100_MMA_TF32_TS)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"setp.ne.b32 p, %4, 0;\n"
		"tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
		"}\n" ::"r"(tmem_d),
		"r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
		: "memory");
}
// 32 lanes x 32 columns of TMEM <-> registers (thread = lane/row, v[j] = column j); the caller's warp
// may only touch lanes 32*(warp%4) .. +31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
	asm volatile(
		"tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
		"{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
		"%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
		"r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
		"r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
		"r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
		"r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
		: "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// lo half of the 3xTF32 split: a - trunc_tf32(a) is exact in fp32; rounded to tf32 (ties away)
__device__ __forceinline__ uint32_t tf32_lo(float v) {
	return tf32_rn(__float_as_uint(v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u)));
}


// ---- kind::f16 (binary16 operands, fp32 accumulate), layouts confirmed on a B200 by scripts/umma_f16_probe.cu:
// K-major SWIZZLE_128B rows of 64 halves (K = 16 step = +32 B) / SWIZZLE_64B rows of 32 halves; MN-major SWIZZLE_128B
// (64-wide n groups LBO apart, 8-row k groups SBO = 1024 B apart, K = 16 step = +2048 B); A from TMEM = two halves per
// 32-bit column, even k in the low half (K = 16 step = +8 columns).
// instruction descriptor: c_format F32 [4,6) = 1, a/b_format F16 = 0
__device__ __forceinline__ uint32_t make_idesc_f16(bool a_mn, bool b_mn, int n, int m) {
	return (1u << 4) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"setp.ne.b32 p, %4, 0;\n"
		"tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
		"}\n" ::"r"(tmem_d),
		"l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
		: "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"setp.ne.b32 p, %4, 0;\n"
		"tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
		"}\n" ::"r"(tmem_d),
		"r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
		: "memory");
}
// 32 lanes x 16 columns of registers -> TMEM
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
	asm volatile(
		"tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
		"{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
		"r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
		"r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
		: "memory");
}
// 3xFP16 split of a (pre-scaled) pair: hi = rn16(x), lo = rn16(x - hi); returns the two packed pairs (first value in the low half)
__device__ __forceinline__ void f16_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
	asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
	float fa, fb;
	asm("{\n.reg .b16 l, h;\nmov.b32 {l, h}, %2;\ncvt.f32.f16 %0, l;\ncvt.f32.f16 %1, h;\n}\n" : "=f"(fa), "=f"(fb) : "r"(hi));
	asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - fb), "f"(a - fa));
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeFn get_encode() {
	static EncodeFn fn = nullptr;
	if (!fn) {
		void* p = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
			fn = (EncodeFn)p;
	}
	return fn;
}

// operand viewed as (rows x contiguous) with an optional batch dimension
inline bool make_map_sw(CUtensorMap* m, const float* base, long long contig_extent, long long rows, long long row_stride,
              int batch, long long batch_stride, int box_contig, int box_rows, CUtensorMapSwizzle sw) {
	EncodeFn enc = get_encode();
	if (!enc) return false;
	cuuint64_t dims[3] = {(cuuint64_t)contig_extent, (cuuint64_t)rows, (cuuint64_t)(batch > 0 ? batch : 1)};
	long long bs = batch_stride > 0 ? batch_stride : row_stride * rows;
	cuuint64_t strides[2] = {(cuuint64_t)row_stride * 4, (cuuint64_t)bs * 4};
	cuuint32_t box[3] = {(cuuint32_t)box_contig, (cuuint32_t)box_rows, 1};
	cuuint32_t es[3] = {1, 1, 1};
	CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
	                 sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	return r == CUDA_SUCCESS;
}
inline bool make_map(CUtensorMap* m, const float* base, long long contig_extent, long long rows, long long row_stride,
              int batch, long long batch_stride, int box_contig, int box_rows, bool mn_major) {
	EncodeFn enc = get_encode();
	if (!enc) return false;
	cuuint64_t dims[3] = {(cuuint64_t)contig_extent, (cuuint64_t)rows, (cuuint64_t)(batch > 0 ? batch : 1)};
	long long bs = batch_stride > 0 ? batch_stride : row_stride * rows;
	cuuint64_t strides[2] = {(cuuint64_t)row_stride * 4, (cuuint64_t)bs * 4};
	cuuint32_t box[3] = {(cuuint32_t)box_contig, (cuuint32_t)box_rows, 1};
	cuuint32_t es[3] = {1, 1, 1};
	CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
	                 mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	return r == CUDA_SUCCESS;
}

// binary16 tensor (two planes folded into the batch dimension by the caller): box_contig <= 64 halves (128-byte swizzle)
inline bool make_map16(CUtensorMap* m, const void* base, long long contig_extent, long long rows, long long row_stride,
              long long batch, long long batch_stride, int box_contig, int box_rows) {
	EncodeFn enc = get_encode();
	if (!enc) return false;
	cuuint64_t dims[3] = {(cuuint64_t)contig_extent, (cuuint64_t)rows, (cuuint64_t)(batch > 0 ? batch : 1)};
	cuuint64_t strides[2] = {(cuuint64_t)row_stride * 2, (cuuint64_t)batch_stride * 2};
	cuuint32_t box[3] = {(cuuint32_t)box_contig, (cuuint32_t)box_rows, 1};
	cuuint32_t es[3] = {1, 1, 1};
	CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
	                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	return r == CUDA_SUCCESS;
}

}  // namespace fh_tc
