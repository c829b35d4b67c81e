"""ctypes binding of libfh_b200.so (the C ABI declared in include/fh_b200.h).

PyTorch is used only for device memory and streams: every wrapper passes raw `data_ptr()`s and
the current CUDA stream. There is NO fallback: if the shared library is missing the import of any
compute entry point raises, and every call on a non-CUDA tensor raises.
"""
import ctypes as C
import os
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfh_b200.so")

GEMM_F32, GEMM_F32_ACC64, GEMM_F64, GEMM_F32xF64_F32, GEMM_TF32X3, GEMM_F64xF32_F32 = 0, 1, 2, 3, 4, 5
EPI_NONE, EPI_DIAG_ADD, EPI_SYMMETRIC = 0, 1, 2


class GemmDesc(C.Structure):
	_fields_ = [("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("batch", C.c_int),
	            ("sa_m", C.c_longlong), ("sa_k", C.c_longlong), ("sb_k", C.c_longlong), ("sb_n", C.c_longlong),
	            ("ldc", C.c_longlong), ("batch_a", C.c_longlong), ("batch_b", C.c_longlong),
	            ("batch_c", C.c_longlong), ("alpha", C.c_double), ("beta", C.c_double), ("dtype", C.c_int),
	            ("epilogue", C.c_int), ("diag", C.c_double), ("kscale", C.c_void_p), ("kscale_batch", C.c_longlong),
	            ("cscale", C.c_void_p), ("cscale_batch", C.c_longlong), ("cscale_recip", C.c_int)]


class RwrDesc(C.Structure):
	_fields_ = [("nb", C.c_int), ("w", C.c_int), ("ldw", C.c_int), ("s", C.c_int), ("k", C.c_int),
	            ("do_conv", C.c_int), ("do_rwr", C.c_int), ("do_col", C.c_int), ("cell0", C.c_int),
	            ("ncell", C.c_int), ("use_tensor_cores", C.c_int), ("nnz", C.c_longlong)]


EXPORTS = ["fh_last_error", "fh_version", "fh_launch_count", "fh_tc_fallback_count", "fh_gemm_batched", "fh_rwr_workspace_bytes",
           "fh_rwr_batched", "fh_densify", "fh_rwr_dense", "fh_colsum_accum", "fh_avgpool", "fh_sqnorm_accum",
           "fh_dot_accum", "fh_polar_workspace_bytes", "fh_polar_batched", "fh_polar_isqrt_multi", "fh_inv_sqrt_spd",
           "fh_cp_als_workspace_bytes", "fh_cp_als", "fh_cp_core_sqnorm", "fh_scale_cols_batched", "fh_timing_enable", "fh_timing_read"]

_lib = None


class FHError(RuntimeError):
	pass


def lib():
	global _lib
	if _lib is None:
		if not os.path.exists(LIB_PATH):
			raise FHError("libfh_b200.so not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'`. "
			              "There is no CPU fallback." % LIB_PATH)
		L = C.CDLL(LIB_PATH)
		L.fh_last_error.restype = C.c_char_p
		L.fh_launch_count.restype = C.c_longlong
		L.fh_tc_fallback_count.restype = C.c_longlong
		for n in ["fh_rwr_workspace_bytes", "fh_polar_workspace_bytes", "fh_cp_als_workspace_bytes"]:
			getattr(L, n).restype = C.c_size_t
		L.fh_rwr_workspace_bytes.argtypes = [C.POINTER(RwrDesc)]
		L.fh_polar_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
		L.fh_cp_als_workspace_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
		vp, ll, ci, sz = C.c_void_p, C.c_longlong, C.c_int, C.c_size_t
		L.fh_gemm_batched.argtypes = [C.POINTER(GemmDesc), vp, vp, vp, vp]
		L.fh_rwr_batched.argtypes = [C.POINTER(RwrDesc), vp, vp, vp, vp, ll, vp, ll, vp, sz, C.POINTER(ci), vp]
		L.fh_densify.argtypes = [C.POINTER(RwrDesc), vp, vp, vp, vp, ll, vp]
		L.fh_rwr_dense.argtypes = [C.POINTER(RwrDesc), vp, ll, vp, ll, vp, sz, C.POINTER(ci), vp]
		L.fh_colsum_accum.argtypes = [vp, ci, ci, ci, ci, ll, vp, ll, vp]
		L.fh_avgpool.argtypes = [vp, ci, ci, ci, ci, ll, ci, vp, ll, vp]
		L.fh_sqnorm_accum.argtypes = [vp, ll, ll, ll, vp, vp]
		L.fh_dot_accum.argtypes = [vp, vp, ll, ll, ll, ll, vp, vp]
		L.fh_polar_batched.argtypes = [vp, vp, ci, ci, ci, ll, ll, vp, vp, ci, vp, sz, C.POINTER(ci), vp]
		L.fh_polar_isqrt_multi.argtypes = [vp, vp, vp, vp, vp, vp, ci, vp, ci, vp, vp]
		L.fh_inv_sqrt_spd.argtypes = [vp, vp, ci, vp, sz, C.POINTER(ci), vp]
		L.fh_cp_als.argtypes = [vp, ci, ci, ci, vp, vp, vp, ci, vp, sz, C.POINTER(C.c_double), vp]
		L.fh_scale_cols_batched.argtypes = [vp, ci, ci, ll, vp, ci, ci, vp, vp]
		L.fh_cp_core_sqnorm.argtypes = [vp, ci, vp, vp, ci, ci, vp, vp, vp]
		L.fh_timing_enable.argtypes = [ci]
		L.fh_timing_enable.restype = None
		L.fh_timing_read.argtypes = [ci, C.POINTER(C.c_double), C.POINTER(ll)]
		_lib = L
	return _lib


def check(rc):
	if rc != 0:
		raise FHError("libfh_b200: %s (code %d)" % (lib().fh_last_error().decode(), rc))


def require_cuda(device, what="fasthigashi_b200"):
	"""The one place that enforces "no CPU path" for the core objects."""
	if torch.device(device).type != "cuda":
		raise FHError("%s runs on CUDA devices only (got %s); there is no CPU path" % (what, device))


def stream_ptr():
	return torch.cuda.current_stream().cuda_stream


def launch_count():
	return int(lib().fh_launch_count())


TIME_DENSIFY, TIME_RWR_CHAIN, TIME_GEMM_TC, TIME_POLAR_JACOBI = 0, 1, 2, 3


def kernel_timing(on):
	"""Per-kernel CUDA-event timing inside the library (bench.py's roofline denominators): on / off, clears the records."""
	lib().fh_timing_enable(1 if on else 0)


def kernel_time(kind):
	"""(total ms, launches) of one kernel kind since kernel_timing(True); synchronises on the recorded events."""
	t, n = C.c_double(0.0), C.c_longlong(0)
	check(lib().fh_timing_read(int(kind), C.byref(t), C.byref(n)))
	return t.value, n.value


def _ptr(t):
	if t is None:
		return None
	if not t.is_cuda:
		raise FHError("fasthigashi_b200 computes on CUDA tensors only (got %s); there is no CPU path" % t.device)
	return t.data_ptr()


_ws_cache = {}


def workspace(nbytes, device, tag="ws"):
	"""A grow-only scratch buffer per (device, tag)."""
	key = (str(device), tag)
	buf = _ws_cache.get(key)
	if buf is None or buf.numel() < nbytes:
		_ws_cache[key] = buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
	return buf


def free_workspaces():
	_ws_cache.clear()


# ---------------------------------------------------------------------------------------------
def gemm(A, B, C_out, M, N, K, sa, sb, ldc, batch=1, batch_strides=(0, 0, 0), alpha=1.0, beta=0.0,
         dtype=GEMM_F32, epilogue=EPI_NONE, diag=0.0, kscale=None, kscale_batch=0, cscale=None,
         cscale_batch=0, cscale_recip=False):
	"""C[b](m,n) = alpha * sum_k A[b](m,k) kscale[b][k] B[b](k,n) (+...). sa = (sa_m, sa_k), sb = (sb_k, sb_n).
	Operands are torch tensors used as raw storage (data_ptr + element strides)."""
	d = GemmDesc()
	d.M, d.N, d.K, d.batch = int(M), int(N), int(K), int(batch)
	d.sa_m, d.sa_k = int(sa[0]), int(sa[1])
	d.sb_k, d.sb_n = int(sb[0]), int(sb[1])
	d.ldc = int(ldc)
	d.batch_a, d.batch_b, d.batch_c = [int(x) for x in batch_strides]
	d.alpha, d.beta, d.dtype, d.epilogue, d.diag = float(alpha), float(beta), int(dtype), int(epilogue), float(diag)
	d.kscale, d.kscale_batch = _ptr(kscale), int(kscale_batch)
	d.cscale, d.cscale_batch, d.cscale_recip = _ptr(cscale), int(cscale_batch), int(bool(cscale_recip))
	check(lib().fh_gemm_batched(C.byref(d), _ptr(A), _ptr(B), _ptr(C_out), stream_ptr()))
	return C_out


def rwr_desc(nb, w, ldw, s, k, do_conv, do_rwr, do_col, cell0, ncell, nnz, use_tc=False):
	d = RwrDesc()
	d.nb, d.w, d.ldw, d.s, d.k = int(nb), int(w), int(ldw), int(s), int(k)
	d.do_conv, d.do_rwr, d.do_col = int(bool(do_conv)), int(bool(do_rwr)), int(bool(do_col))
	d.cell0, d.ncell, d.use_tensor_cores, d.nnz = int(cell0), int(ncell), int(bool(use_tc)), int(nnz)
	return d


def scale_cols_batched(F, rows, r, ldf, Arows, nb, ldo, out):
	check(lib().fh_scale_cols_batched(_ptr(F), int(rows), int(r), int(ldf), _ptr(Arows), int(nb), int(ldo), _ptr(out), stream_ptr()))
	return out
