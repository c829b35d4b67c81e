"""TEST INFRASTRUCTURE ONLY: a host-memory stand-in for libfh_b200.so (include/fh_b200.h), so that the Python
ORCHESTRATION of the product (`Fast_Higashi_core`: what is computed when, which operand layouts and strides cross the C
ABI, what is all-reduced between ranks) can run in the CPU test suite - including world_size-2 `gloo` runs of the whole
cell-sharded sweep. Every entry point takes the same arguments as the C function (raw addresses, the ctypes descriptor
structs) and evaluates its documented contract with numpy / torch-CPU in fp64; the RWR and CP-ALS entries call the pinned
oracle. It says nothing about the CUDA kernels: those are checked on the GPU by tests/test_gpu_parity.py.

`install()` swaps it in by patching `_lib.lib`, `_lib._ptr`, `_lib.stream_ptr`, `_lib.require_cuda` and the handful of
`torch.cuda` stream objects the core touches; it returns an undo function. Nothing in the product imports this module.
"""
import contextlib
import ctypes as C

import numpy as np
import torch

from oracle import fh_oracle as O

_NP = {"f32": np.float32, "f64": np.float64, "i32": np.int32, "i16": np.int16, "i64": np.int64}
# FH_GEMM_*: (A, B, C) element types
_GEMM_TYPES = {0: ("f32", "f32", "f32"), 1: ("f32", "f32", "f64"), 2: ("f64", "f64", "f64"), 3: ("f32", "f64", "f32"),
               4: ("f32", "f32", "f32"), 5: ("f64", "f32", "f32")}


def arr(addr, kind, count):
	"""Writable numpy view of `count` elements at the raw host address `addr`."""
	dt = np.dtype(_NP[kind])
	if count <= 0:
		return np.empty(0, dt)
	if not addr:
		raise ValueError("NULL pointer for %d %s elements" % (count, kind))
	buf = (C.c_char * (int(count) * dt.itemsize)).from_address(int(addr))
	return np.frombuffer(buf, dtype=dt, count=int(count))


def strided(addr, kind, shape, strides):
	"""View with element strides (a stride may be 0: broadcast operand)."""
	extent = 1 + sum((n - 1) * s for n, s in zip(shape, strides))
	base = arr(addr, kind, extent)
	item = base.dtype.itemsize
	return np.lib.stride_tricks.as_strided(base, shape=tuple(int(n) for n in shape), strides=tuple(int(s) * item for s in strides))


def _obj(ref):
	"""ctypes.byref(x) -> x ; None -> None."""
	return None if ref is None else getattr(ref, "_obj", ref)


class FakeLib:
	def __init__(self):
		self.calls = {}
		self.error = b""

	def _count(self, name):
		self.calls[name] = self.calls.get(name, 0) + 1

	# -- bookkeeping ------------------------------------------------------------------------------
	def fh_last_error(self):
		return self.error

	def fh_version(self):
		return 100

	def fh_launch_count(self):
		return sum(self.calls.values())

	def fh_tc_fallback_count(self):
		return 0

	def fh_rwr_workspace_bytes(self, d):
		return 256

	def fh_polar_workspace_bytes(self, batch, rows, cols):
		return 256

	def fh_cp_als_workspace_bytes(self, n, r, R):
		return 256

	# -- GEMM -------------------------------------------------------------------------------------
	def fh_gemm_batched(self, dref, A, B, Cp, stream):
		self._count("gemm")
		d = _obj(dref)
		ka, kb, kc = _GEMM_TYPES[d.dtype]
		if 1 not in (d.sa_m, d.sa_k) or 1 not in (d.sb_k, d.sb_n):
			self.error = b"one stride of each operand must be 1"
			return -1
		a = strided(A, ka, (d.batch, d.M, d.K), (d.batch_a, d.sa_m, d.sa_k)).astype(np.float64)
		b = strided(B, kb, (d.batch, d.K, d.N), (d.batch_b, d.sb_k, d.sb_n)).astype(np.float64)
		if d.kscale:
			a = a * strided(d.kscale, "f32", (d.batch, 1, d.K), (d.kscale_batch, 0, 1)).astype(np.float64)
		v = d.alpha * np.matmul(a, b)
		if d.epilogue == 1:
			i = np.arange(min(d.M, d.N))
			v[:, i, i] += d.diag
		if d.cscale:
			s = strided(d.cscale, "f32", (d.batch, 1, d.N), (d.cscale_batch, 0, 1)).astype(np.float64)
			v = v / s if d.cscale_recip else v * s
		c = strided(Cp, kc, (d.batch, d.M, d.N), (d.batch_c, d.ldc, 1))
		if d.beta != 0.0:
			v = v + d.beta * c.astype(np.float64)
		c[...] = v.astype(c.dtype)
		return 0

	def fh_scale_cols_batched(self, F, rows, r, ldf, Arows, nb, ldo, out, stream):
		self._count("scale_cols")
		f = strided(F, "f32", (rows, r), (ldf, 1))
		a = strided(Arows, "f32", (nb, r), (r, 1))
		o = strided(out, "f32", (nb, rows, ldo), (rows * ldo, ldo, 1))
		o[...] = 0
		o[:, :, :r] = f[None, :, :] * a[:, None, :]
		return 0

	# -- RWR --------------------------------------------------------------------------------------
	@staticmethod
	def _densify(d, rowptr, col, val):
		"""(ncell, nb, w) fp32 with the 1e-8 floor, from the block CSR rows of cells [cell0, cell0 + ncell)."""
		rows = d.ncell * d.nb
		rp = arr(rowptr + 4 * d.cell0 * d.nb, "i32", rows + 1).astype(np.int64)
		lo, hi = int(rp[0]), int(rp[-1])
		cc = arr(col + 2 * lo, "i16", hi - lo).astype(np.int64)
		vv = arr(val + 4 * lo, "f32", hi - lo)
		x = np.zeros((rows, d.w), np.float32)
		x[np.repeat(np.arange(rows), np.diff(rp)), cc] = vv
		return torch.from_numpy(np.maximum(x, np.float32(1e-8)).reshape(d.ncell, d.nb, d.w))

	@staticmethod
	def _store_panel(d, x, out, cell_stride):
		o = strided(out, "f32", (d.ncell, d.nb, d.ldw), (cell_stride, d.ldw, 1))
		o[...] = 0
		o[:, :, :d.w] = x.numpy() if torch.is_tensor(x) else x

	def _rwr(self, d, x, bin_cov, bin_cov_ld, host_n_iter):
		cov = None
		if d.do_col and d.do_rwr:
			cov = torch.from_numpy(strided(bin_cov, "f32", (d.ncell, d.w), (bin_cov_ld, 1)).copy())
		y, n_it = O.partial_rwr(x, d.s, d.s + d.nb, bool(d.do_conv), bool(d.do_rwr), bool(d.do_col), cov, d.k)
		n = _obj(host_n_iter)
		if n is not None:
			n.value = int(n_it)
		return y

	def fh_rwr_batched(self, dref, rowptr, col, val, bin_cov, bin_cov_ld, out, out_cell_stride, ws, ws_bytes, host_n_iter, stream):
		self._count("rwr_batched")
		d = _obj(dref)
		y = self._rwr(d, self._densify(d, rowptr, col, val), bin_cov, bin_cov_ld, host_n_iter)
		self._store_panel(d, y, out, out_cell_stride)
		return 0

	def fh_densify(self, dref, rowptr, col, val, out, out_cell_stride, stream):
		self._count("densify")
		d = _obj(dref)
		self._store_panel(d, self._densify(d, rowptr, col, val), out, out_cell_stride)
		return 0

	def fh_rwr_dense(self, dref, x, cell_stride, bin_cov, bin_cov_ld, ws, ws_bytes, host_n_iter, stream):
		self._count("rwr_dense")
		d = _obj(dref)
		xin = torch.from_numpy(strided(x, "f32", (d.ncell, d.nb, d.w), (cell_stride, d.ldw, 1)).copy())
		self._store_panel(d, self._rwr(d, xin, bin_cov, bin_cov_ld, host_n_iter), x, cell_stride)
		return 0

	def fh_colsum_accum(self, x, ncell, nb, w, ldw, cell_stride, cov, cov_ld, stream):
		self._count("colsum")
		xv = strided(x, "f32", (ncell, nb, w), (cell_stride, ldw, 1))
		cv = strided(cov, "f32", (ncell, w), (cov_ld, 1))
		cv += xv.sum(1, dtype=np.float64).astype(np.float32)
		return 0

	def fh_avgpool(self, x, ncell, nb, w, ldw, cell_stride, ll, out, out_cell_stride, stream):
		self._count("avgpool")
		orow, ocol = nb // ll, w // ll
		xv = strided(x, "f32", (ncell, nb, w), (cell_stride, ldw, 1))[:, :orow * ll, :ocol * ll]
		pooled = xv.reshape(ncell, orow, ll, ocol, ll).mean(axis=(2, 4), dtype=np.float64)
		strided(out, "f32", (ncell, orow * ocol), (out_cell_stride, 1))[...] = pooled.reshape(ncell, -1).astype(np.float32)
		return 0

	def fh_sqnorm_accum(self, x, rows, cols, ld, acc, stream):
		self._count("sqnorm")
		arr(acc, "f64", 1)[0] += float(np.square(strided(x, "f32", (rows, cols), (ld, 1)).astype(np.float64)).sum())
		return 0

	def fh_dot_accum(self, x, y, rows, cols, ldx, ldy, acc, stream):
		self._count("dot")
		a = strided(x, "f32", (rows, cols), (ldx, 1)).astype(np.float64)
		b = strided(y, "f32", (rows, cols), (ldy, 1)).astype(np.float64)
		arr(acc, "f64", 1)[0] += float((a * b).sum())
		return 0

	# -- polar ------------------------------------------------------------------------------------
	@staticmethod
	def _eig_factor(G):
		"""Rows e_j * lambda_j^(-1/4) (so that WT^T WT = G^(-1/2)) and sum sqrt(lambda); directions below the
		rank-revealing threshold n * eps * lambda_max are dropped, as the pivoted Cholesky of the kernel does."""
		lam, E = np.linalg.eigh((G + G.T) / 2)
		keep = lam > lam.max() * G.shape[0] * np.finfo(np.float64).eps
		scale = np.where(keep, np.maximum(lam, 1e-300) ** -0.25, 0.0)
		return (E * scale[None, :]).T, float(np.sqrt(lam[keep]).sum())

	def fh_polar_isqrt_multi(self, G_all, WT_all, dev_n, dev_off, dev_slot, host_n, count, sigma_sum, max_sweeps, dev_nsweep, stream):
		self._count("polar_isqrt_multi")
		n = arr(dev_n, "i32", count)
		if not np.array_equal(n, arr(host_n, "i32", count)) or (count > 1 and np.any(np.diff(n) > 0)):
			self.error = b"problem table: host/device sizes differ or not sorted by decreasing size"
			return -1
		off, slot = arr(dev_off, "i64", count), arr(dev_slot, "i32", count)
		for i in range(count):
			ni = int(n[i])
			G = arr(G_all + 8 * int(off[i]), "f64", ni * ni).reshape(ni, ni)
			WT, ssum = self._eig_factor(G)
			arr(WT_all + 8 * int(off[i]), "f64", ni * ni)[:] = WT.reshape(-1)
			arr(sigma_sum + 8 * int(slot[i]), "f64", 1)[0] = ssum
			if dev_nsweep:
				arr(dev_nsweep + 4 * int(slot[i]), "i32", 1)[0] = 1
		return 0

	def fh_inv_sqrt_spd(self, G, out, n, ws, ws_bytes, host_iters, stream):
		self._count("inv_sqrt_spd")
		g = arr(G, "f64", n * n).reshape(n, n)
		lam, E = np.linalg.eigh((g + g.T) / 2)
		arr(out, "f64", n * n)[:] = ((E / np.sqrt(lam)[None, :]) @ E.T).reshape(-1)
		it = _obj(host_iters)
		if it is not None:
			it.value = 1
		return 0

	def fh_polar_batched(self, T, U, batch, rows, cols, ld, batch_stride, sigma_sum, sigma, max_sweeps, ws, ws_bytes, host_max_sweeps, stream):
		self._count("polar_batched")
		t = strided(T, "f32", (batch, rows, cols), (batch_stride, ld, 1)).astype(np.float64)
		u, s, vh = np.linalg.svd(t, full_matrices=False)
		strided(U, "f32", (batch, rows, cols), (batch_stride, ld, 1))[...] = (u @ vh).astype(np.float32)
		if sigma_sum:
			arr(sigma_sum, "f64", batch)[:] = s.sum(-1)
		if sigma:
			arr(sigma, "f64", batch * s.shape[-1])[:] = s.reshape(-1)
		m = _obj(host_max_sweeps)
		if m is not None:
			m.value = 1
		return 0

	# -- CP-ALS -----------------------------------------------------------------------------------
	def fh_cp_als(self, Y, n, r, R, A, B, D, n_iter_max, ws, ws_bytes, host_out, stream):
		self._count("cp_als")
		y = torch.from_numpy(arr(Y, "f32", n * r * R).reshape(n, r, R).copy())
		a, b, dd = arr(A, "f32", n * r).reshape(n, r), arr(B, "f32", r * r).reshape(r, r), arr(D, "f32", R * r).reshape(R, r)
		fac, norm_hat, inner = O.cp_als(y, [torch.from_numpy(x.copy()) for x in (a, b, dd)], int(n_iter_max))
		a[...], b[...], dd[...] = fac[0].numpy(), fac[1].numpy(), fac[2].numpy()
		if host_out is not None:
			host_out[0], host_out[1] = (norm_hat, inner) if n_iter_max > 1 else (0.0, 0.0)
		return 0

	def fh_cp_core_sqnorm(self, A, n, B, D, R, r, ws, acc, stream):
		self._count("core_sqnorm")
		a = arr(A, "f32", n * r).reshape(n, r).astype(np.float64)
		b = arr(B, "f32", r * r).reshape(r, r).astype(np.float64)
		d = arr(D, "f32", R * r).reshape(R, r).astype(np.float64)
		arr(acc, "f64", 1)[0] += float(((a.T @ a) * (b.T @ b) * (d.T @ d)).sum())
		return 0


class _NoStream:
	cuda_stream = 0

	def wait_event(self, e): pass
	def wait_stream(self, s): pass
	def synchronize(self): pass


class _NoEvent:
	def __init__(self, *a, **k): pass
	def record(self, *a): pass
	def elapsed_time(self, other): return 0.0


def install():
	"""Route the package's C-ABI calls to a FakeLib working on host memory. Returns (fake, undo)."""
	import fasthigashi_b200  # noqa: F401
	from fasthigashi_b200 import _lib
	fake = FakeLib()
	saved = [(_lib, n, getattr(_lib, n)) for n in ("lib", "_ptr", "stream_ptr", "require_cuda")]
	saved += [(torch.cuda, n, getattr(torch.cuda, n)) for n in ("current_stream", "Stream", "Event", "stream", "synchronize")]
	_lib.lib = lambda: fake
	_lib._ptr = lambda t: None if t is None else t.data_ptr()
	_lib.stream_ptr = lambda: 0
	_lib.require_cuda = lambda device, what="": None
	torch.cuda.current_stream = lambda *a, **k: _NoStream()
	torch.cuda.Stream = lambda *a, **k: _NoStream()
	torch.cuda.Event = _NoEvent
	torch.cuda.stream = lambda s: contextlib.nullcontext()
	torch.cuda.synchronize = lambda *a, **k: None
	_lib.free_workspaces()

	def undo():
		for obj, name, val in saved:
			setattr(obj, name, val)
		_lib.free_workspaces()
	return fake, undo
