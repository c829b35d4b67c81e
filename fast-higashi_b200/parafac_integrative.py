"""Inner CP-ALS on the projected tensor (mirror of fasthigashi/parafac_integrative.py:28-112)."""
import ctypes as C
import torch
from . import _lib


def cp_als_(Y, A, B, D, n_iter_max, tag="cp", want_norms=False):
	"""In-place device CP-ALS: Y (n, r, R) fp32 contiguous; A (n,r), B (r,r), D (R,r) fp32 contiguous. The reference's early
	stop (relative loss change < 1e-5) is decided on the device; nothing is read back unless `want_norms`, which returns
	(||Xhat||^2, <Xhat, Y>) of the last committed iteration (zeros unless n_iter_max > 1) and synchronises the stream."""
	n, r, R = Y.shape
	lib = _lib.lib()
	ws = _lib.workspace(lib.fh_cp_als_workspace_bytes(n, r, R), Y.device, tag)
	out = (C.c_double * 2)() if want_norms else None
	_lib.check(lib.fh_cp_als(Y.data_ptr(), n, r, R, A.data_ptr(), B.data_ptr(), D.data_ptr(), int(n_iter_max),
	                         ws.data_ptr(), ws.numel(), out, _lib.stream_ptr()))
	return (out[0], out[1]) if want_norms else (0.0, 0.0)


def core_sqnorm_accum(A, B, D, acc, tag="core"):
	"""acc (device fp64 scalar tensor) += ||[[A,B,D]]||^2 (parafac2_intergrative.py:623-632)."""
	n, r = A.shape
	R = D.shape[0]
	ws = _lib.workspace(3 * r * r * 8, A.device, tag + "_core")
	_lib.check(_lib.lib().fh_cp_core_sqnorm(A.data_ptr(), n, B.data_ptr(), D.data_ptr(), R, r, ws.data_ptr(),
	                                        acc.data_ptr(), _lib.stream_ptr()))


@torch.no_grad()
def parafac(X, rank, n_iter_max=100, init=None, verbose=False, common_factor=None):
	"""Reference-compatible entry: returns (factors, ||Xhat||^2, <Xhat, X>)."""
	if not X.is_cuda:
		raise _lib.FHError("parafac: CUDA tensor required (no CPU path in fasthigashi_b200)")
	Y = X.contiguous().float()
	A, B, D = [f.to(Y.device, torch.float32).contiguous().clone() for f in init]
	norm_hat, inner = cp_als_(Y, A, B, D, n_iter_max, want_norms=True)
	return [A, B, D], norm_hat, inner
