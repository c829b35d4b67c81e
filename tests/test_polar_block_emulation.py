"""Host emulation of the block one-sided Jacobi (csrc/fh_polar_block.cuh, opt-in FH_POLAR_BLOCK=1 on the device):
the SAME source compiled with FH_EMU runs its barrier-separated phases as loops over threads. Checks
  * numerics against an fp64 SVD and against the numpy restatement of the scalar kernel (graded spectra to kappa 3e6),
  * forward vs reverse thread order bit-identical (no dependence between threads inside a phase = no missing barrier),
  * independence of the CTA size, edge sizes (1, 8, 9, 16, 17, the 152 limit), rank-deficient input."""
import ctypes as C
import os
import subprocess
import sys
import numpy as np
import pytest
from conftest import ROOT

EMU_DIR = os.path.join(ROOT, "tests", "emu")
sys.path.insert(0, EMU_DIR)
from jacobi_numpy import pivoted_cholesky_upper, jacobi_rows, polar_from_rows  # noqa: E402


@pytest.fixture(scope="module")
def emu():
	so = os.path.join(EMU_DIR, "libpolar_block_emu.so")
	src = [os.path.join(EMU_DIR, "polar_block_emu.cpp"), os.path.join(ROOT, "fast-higashi_b200", "csrc", "fh_polar_block.cuh")]
	if not os.path.exists(so) or max(os.path.getmtime(s) for s in src) > os.path.getmtime(so):
		subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so, src[0]], check=True)
	L = C.CDLL(so)
	L.fh_emu_bj_scratch_doubles.restype = C.c_longlong
	L.fh_emu_block_jacobi.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p]
	return L


def run_block(L, Rn, nthreads=1024, reverse=0, skip=1e-17, max_sweeps=30):
	n = Rn.shape[0]
	rows, ld = L.fh_emu_bj_rows(n), L.fh_emu_bj_ld(n)
	R = np.zeros((rows, ld))
	R[:n, :n] = Rn
	scratch = np.full(int(L.fh_emu_bj_scratch_doubles(n)) + 2, np.nan)   # uninitialised shared memory
	sweeps = L.fh_emu_block_jacobi(R.ctypes.data, n, nthreads, max_sweeps, skip, reverse, scratch.ctypes.data)
	assert not R[n:].any() and not R[:, n:].any()                          # padding stays zero
	return R[:n, :n].copy(), sweeps


def graded(rows, cols, logk, seed):
	rng = np.random.default_rng(seed)
	n = min(rows, cols)
	Uq, _ = np.linalg.qr(rng.standard_normal((max(rows, cols), n)))
	Vq, _ = np.linalg.qr(rng.standard_normal((n, n)))
	T = ((Uq * np.logspace(0, -logk, n)) @ Vq.T).astype(np.float32).astype(np.float64)
	return T


@pytest.mark.parametrize("rows,cols,logk", [(316, 137, 2.0), (316, 144, 5.5), (215, 129, 6.5), (152, 150, 5.5), (72, 21, 6.0)])
def test_block_jacobi_polar_accuracy(emu, rows, cols, logk):
	T = graded(rows, cols, logk, 1)
	Ud, Sd, Vhd = np.linalg.svd(T, full_matrices=False)
	truth = Ud @ Vhd
	R0, perm = pivoted_cholesky_upper(T.T @ T)
	Rs = R0.copy()
	sw_scalar = jacobi_rows(Rs)
	Rb, sw_block = run_block(emu, R0)
	err_s = np.linalg.norm(polar_from_rows(T, Rs, perm) - truth) / np.linalg.norm(truth)
	Ub = polar_from_rows(T, Rb, perm)
	err_b = np.linalg.norm(Ub - truth) / np.linalg.norm(truth)
	assert sw_block <= sw_scalar + 2 and sw_block <= 12, (sw_block, sw_scalar)
	assert err_b <= 3 * err_s + 1e-9, (err_b, err_s)
	assert err_b < (2e-6 if logk <= 5.5 else 1e-3)                         # the GPU test's limits (test_gpu_parity.py)
	# rows orthogonal to the stopping rule: cos^2 <= 1e-11 wherever the skip rule would still rotate
	Gm = Rb @ Rb.T
	d = np.diag(Gm)
	cos2 = Gm * Gm / np.outer(d, d)
	np.fill_diagonal(cos2, 0)
	act = Gm * Gm > 1e-17 * np.minimum.outer(d, d) ** 2
	np.fill_diagonal(act, False)
	assert (cos2[act].max() if act.any() else 0.0) <= 1e-10
	# singular values: sqrt of the row norms
	np.testing.assert_allclose(np.sort(np.sqrt(d))[::-1], Sd, rtol=0, atol=1e-9 * Sd[0])


@pytest.mark.parametrize("n", [1, 2, 7, 8, 9, 16, 17, 24, 31, 40, 152])
def test_block_jacobi_edge_sizes_and_thread_order(emu, n):
	assert n <= emu.fh_emu_bj_max_side()
	rng = np.random.default_rng(n)
	A = rng.standard_normal((n + 5, n)) * np.logspace(0, -4, n)
	R0, perm = pivoted_cholesky_upper(A.T @ A)
	out = [run_block(emu, R0, nthreads=nt, reverse=rev) for nt, rev in ((1024, 0), (1024, 1), (256, 0), (512, 1))]
	for R, sw in out[1:]:
		assert sw == out[0][1]
		assert np.array_equal(R, out[0][0])                                # bit-identical: no intra-phase dependence, no CTA-size dependence
	R, sw = out[0]
	Gm = R @ R.T
	d = np.diag(Gm)
	off = np.abs(Gm - np.diag(d)) / np.sqrt(np.outer(d, d))
	assert off.max() < 1e-5 and sw <= 12
	np.testing.assert_allclose(np.sort(d)[::-1], np.sort(np.linalg.eigvalsh(A.T @ A))[::-1], rtol=1e-9, atol=1e-13 * d.max())


def test_block_jacobi_rank_deficient_and_diagonal(emu):
	R0 = np.diag(np.arange(1.0, 21.0))                                     # already orthogonal: one sweep, untouched
	R, sw = run_block(emu, R0)
	assert sw == 1 and np.array_equal(R, R0)
	rng = np.random.default_rng(3)
	B = rng.standard_normal((30, 6)) @ rng.standard_normal((6, 20))       # rank 6: pivoted Cholesky replaces the tail by eps * I
	R0, perm = pivoted_cholesky_upper(B.T @ B)
	R, sw = run_block(emu, R0)
	assert sw <= 12 and np.isfinite(R).all()
	lam = np.sort((R * R).sum(1))[::-1]
	np.testing.assert_allclose(lam[:6], np.sort(np.linalg.eigvalsh(B.T @ B))[::-1][:6], rtol=1e-9)
