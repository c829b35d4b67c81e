// tcgen05 3xTF32 GEMM (placeholder until the tensor-core kernel lands; never silently falls back).
#include "fh_common.cuh"
#include "../../include/fh_b200.h"
int fh_gemm_tc(const fh_gemm_desc* d, const float* A, const float* B, float* C, void* stream) {
	fh_set_error("FH_GEMM_TF32X3 not built into this library");
	return FH_ERR_UNSUPPORTED;
}
