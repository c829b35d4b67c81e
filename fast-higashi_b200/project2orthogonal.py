"""Polar factor on the B200 (mirror of fasthigashi/project2orthogonal.py:6-55).

`project2orthogonal(matrix, rank, compute_device)` keeps the reference signature and return value
(U Vh of the thin SVD, leading singular values). Two device paths, both via the C ABI:
  * batched small matrices (one per bin, parafac2_intergrative.py:396): fp64 Gram + fp64 Jacobi
    eigensolver (`fh_polar_batched`);
  * one tall matrix (cells x R, parafac2_intergrative.py:483,831): fp64 Gram, all-reduced across
    ranks when cell-sharded, inverse square root by coupled Newton-Schulz (`fh_inv_sqrt_spd`).
"""
import ctypes as C
import os
import torch
from . import _lib

JACOBI_MAX_SIDE = 160  # Gram side that still fits shared memory (fh_polar.cu)


def polar_batched(T, rows, cols, ld, out=None, want_sigma=False, want_sweeps=False):
	"""T: (batch, rows, ld) fp32 device (columns >= cols are ignored/kept). Returns (U, sigma_sum
	(batch,) fp64 device, sigma (batch, min(rows, cols)) fp64 or None[, max Jacobi sweeps])."""
	batch = T.shape[0]
	dev = T.device
	U = torch.zeros_like(T) if out is None else out
	n = min(rows, cols)
	ssum = torch.empty(batch, dtype=torch.float64, device=dev)
	sig = torch.empty(batch, n, dtype=torch.float64, device=dev) if want_sigma else None
	lib = _lib.lib()
	ws = _lib.workspace(lib.fh_polar_workspace_bytes(batch, rows, cols), dev, "polar")
	nsw = C.c_int(0)
	_lib.check(lib.fh_polar_batched(T.data_ptr(), U.data_ptr(), batch, rows, cols, ld, T.stride(0), ssum.data_ptr(),
	                                None if sig is None else sig.data_ptr(), int(os.environ.get("FH_POLAR_SWEEPS", "0")),
	                                ws.data_ptr(), ws.numel(), C.byref(nsw) if want_sweeps else None, _lib.stream_ptr()))
	if want_sweeps:
		return U, ssum, sig, nsw.value
	return U, ssum, sig


def inv_sqrt_spd(G):
	"""G^{-1/2} of one SPD fp64 matrix on device."""
	n = G.shape[0]
	out = torch.empty_like(G)
	ws = _lib.workspace((5 * n * n + 512) * 8, G.device, "ns")
	it = C.c_int(0)
	_lib.check(_lib.lib().fh_inv_sqrt_spd(G.data_ptr(), out.data_ptr(), n, ws.data_ptr(), ws.numel(), C.byref(it),
	                                      _lib.stream_ptr()))
	return out


def polar_tall(M, group=None):
	"""Polar factor of one tall (rows x R) fp32 matrix whose rows may be sharded over `group`
	(torch.distributed): V = M (M^T M)^{-1/2}, Gram in fp64 and all-reduced."""
	rows, R = M.shape
	if group is None and rows < R:
		# fewer rows than columns (fewer cells than the rank; the reference's SVD route handles it, project2orthogonal.py:6-29):
		# the R x R Gram is singular, the polar factor is (M M^T)^{-1/2} M through the rows x rows Gram
		G = torch.empty(rows, rows, dtype=torch.float64, device=M.device)
		_lib.gemm(M, M, G, rows, rows, R, (M.stride(0), 1), (1, M.stride(0)), rows, dtype=_lib.GEMM_F32_ACC64)
		Gi = inv_sqrt_spd(G)
		V = torch.empty(rows, R, dtype=torch.float32, device=M.device)
		_lib.gemm(Gi, M, V, rows, R, rows, (rows, 1), (M.stride(0), 1), R, dtype=_lib.GEMM_F64xF32_F32)
		return V
	G = torch.empty(R, R, dtype=torch.float64, device=M.device)
	_lib.gemm(M, M, G, R, R, rows, (1, M.stride(0)), (M.stride(0), 1), R, dtype=_lib.GEMM_F32_ACC64)
	if group is not None:
		import torch.distributed as dist
		dist.all_reduce(G, group=group)
	Gi = inv_sqrt_spd(G)
	V = torch.empty(rows, R, dtype=torch.float32, device=M.device)
	_lib.gemm(M, Gi, V, rows, R, R, (M.stride(0), 1), (R, 1), R, dtype=_lib.GEMM_F32xF64_F32)
	return V


@torch.no_grad()
def project2orthogonal(matrix, rank, compute_device=None):
	"""Reference-compatible entry (project2orthogonal.py:6): matrix (..., d1, d2) CUDA fp32 ->
	(U Vh (..., d1, d2), S (..., min(d1, d2)) descending). `rank` must equal min(d1, d2) (the only
	way the reference calls it)."""
	if not matrix.is_cuda:
		raise _lib.FHError("project2orthogonal: CUDA tensor required (no CPU path in fasthigashi_b200)")
	d1, d2 = matrix.shape[-2:]
	n = min(d1, d2)
	if rank is not None and rank < n:
		raise NotImplementedError("truncated polar (rank < min(shape)) is not used on the hot path")
	if n > JACOBI_MAX_SIDE:
		if matrix.dim() != 2 or d1 < d2:
			raise NotImplementedError("batched polar with Gram side > %d" % JACOBI_MAX_SIDE)
		V = polar_tall(matrix.contiguous().float())
		S = torch.linalg.svdvals((V.T.double() @ matrix.double()))  # small R x R, API completeness only
		return V, S.float()
	lead = matrix.shape[:-2]
	T = matrix.reshape(-1, d1, d2).contiguous().float()
	U, _, sig = polar_batched(T, d1, d2, d2, want_sigma=True)
	S = torch.sort(sig, dim=-1, descending=True).values.float()
	return U.reshape(*lead, d1, d2), S.reshape(*lead, n)
