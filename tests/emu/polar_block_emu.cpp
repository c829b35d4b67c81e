// Host build of fast-higashi_b200/csrc/fh_polar_block.cuh (FH_EMU): the block Jacobi phases run as loops
// over threads, forward or reverse. TEST INFRASTRUCTURE ONLY (tests/test_polar_block_emulation.py).
#define FH_EMU 1
#define __host__
#define __device__
#include <cstddef>
#include <cstring>
int fh_emu_reverse = 0;
int fh_emu_cross_only = 1;
int fh_emu_fp32_angle = 1;
long long fh_emu_count[4] = {0, 0, 0, 0};
#include "../../fast-higashi_b200/csrc/fh_polar_block.cuh"

extern "C" {
void fh_emu_set_policy(int cross_only) { fh_emu_cross_only = cross_only; }
void fh_emu_set_fp32_angle(int on) { fh_emu_fp32_angle = on; }
void fh_emu_counters(long long* out, int reset) {
	for (int i = 0; i < 4; ++i) { out[i] = fh_emu_count[i]; if (reset) fh_emu_count[i] = 0; }
}
int fh_emu_bj_rows(int n) { return bj_rows(n); }
int fh_emu_bj_ld(int n) { return bj_ld(n); }
long long fh_emu_bj_scratch_doubles(int n) { return (long long)bj_scratch_doubles(n); }
int fh_emu_bj_max_side(void) { return kBJMaxSide; }
// R: bj_rows(n) x bj_ld(n) doubles (padding zero), orthogonalised in place; returns the sweep count
int fh_emu_block_jacobi(double* R, int n, int nthreads, int max_sweeps, double skip_tol, int reverse, double* scratch) {
	fh_emu_reverse = reverse;
	return block_jacobi_sweeps(R, n, bj_ld(n), scratch, nthreads, max_sweeps, skip_tol);
}
}
