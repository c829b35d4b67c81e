"""Partial RWR imputation on the B200 (mirror of fasthigashi/partial_rwr.py).

`partial_rwr` keeps the reference signature (partial_rwr.py:47-60) for dense `(cells, nb, w)` CUDA
input; `rwr_block_csr` is the native entry the PARAFAC2 driver uses: it imputes straight from the
device-resident block-CSR (densify + conv + RWR in one C-ABI call, fh_rwr_batched).
"""
import ctypes as C
import os
import torch
from . import _lib

RWR_SCRATCH_BYTES = int(os.environ.get("FH_RWR_SCRATCH_MB", "640")) << 20  # per-chunk intermediates (A, A A^T, P, Q); measured: 640 MB chunks beat L2-sized 80 MB ones (launch count, wave quantisation)


def pad4(w):
	return (int(w) + 3) // 4 * 4


def cells_per_chunk(nb, ldw, limit=RWR_SCRATCH_BYTES):
	ldp = pad4(nb)
	per_cell = 4 * (nb * ldw + 3 * nb * ldp) + 4
	return int(max(8, min(32768, limit // per_cell)))


def rwr_block_csr(ds, b, cell0, ncell, out, out_cell_stride, k, do_conv, do_rwr, do_col, bin_cov=None,
                  use_tc=False, chunk=None):
	"""Impute cells [cell0, cell0+ncell) of bin-block `b` of the block-CSR dataset `ds` into
	`out` (device fp32; cell c, row r, col j at out[c*out_cell_stride + r*ldw + j]).
	k >= 0: forced step count (every call of the ALS sweep); k < 0: the reference's auto-stop over
	the WHOLE cell range of this call (partial_rwr.py:119-123). bin_cov: (>=cell0+ncell, n_bins)
	device fp32, the chromosome-wide coverage (only read when do_col).
	Returns the step count the reference would report."""
	g = ds.geoms[b]
	ldw = pad4(g.w)
	dev = out.device
	rowptr, col, val = ds.rowptr[b], ds.col[b], ds.val[b]
	lib = _lib.lib()
	n_iter = C.c_int(0)
	if k < 0 or chunk is None:
		chunk = ncell if k < 0 else cells_per_chunk(g.nb, ldw)
	worst = 0
	for c0 in range(0, ncell, chunk):
		nc = min(chunk, ncell - c0)
		d = _lib.rwr_desc(g.nb, g.w, ldw, g.s, k, do_conv, do_rwr, do_col, cell0 + c0, nc, val.numel(), use_tc)
		nbytes = lib.fh_rwr_workspace_bytes(C.byref(d))
		ws = _lib.workspace(nbytes, dev, "rwr")
		cov_ptr, cov_ld = None, 0
		if do_col and do_rwr:
			cov_ld = bin_cov.stride(0)
			cov_ptr = bin_cov.data_ptr() + 4 * ((cell0 + c0) * cov_ld + g.col0)
		_lib.check(lib.fh_rwr_batched(C.byref(d), rowptr.data_ptr(), col.data_ptr(), val.data_ptr(), cov_ptr, cov_ld,
		                              out.data_ptr() + 4 * c0 * out_cell_stride, out_cell_stride,
		                              ws.data_ptr(), ws.numel(), C.byref(n_iter), _lib.stream_ptr()))
		worst = max(worst, n_iter.value)
	return worst


def densify_block(ds, b, cell0, ncell, out=None):
	"""Device densify of a block (sparse_for_schic.py:279-320): (ncell, nb, ldw) with floor 1e-8."""
	g = ds.geoms[b]
	ldw = pad4(g.w)
	if out is None:
		out = torch.empty(ncell, g.nb, ldw, dtype=torch.float32, device=ds.val[b].device)
	d = _lib.rwr_desc(g.nb, g.w, ldw, g.s, 0, False, False, False, cell0, ncell, ds.val[b].numel())
	_lib.check(_lib.lib().fh_densify(C.byref(d), ds.rowptr[b].data_ptr(), ds.col[b].data_ptr(), ds.val[b].data_ptr(),
	                                 out.data_ptr(), g.nb * ldw, _lib.stream_ptr()))
	return out


@torch.no_grad()
def partial_rwr(x, slice_start, slice_end, do_conv, do_rwr, do_col, bin_cov=torch.ones(1),
                bin_cov_row=torch.ones(1), return_rwr_iter=False, force_rwr_epochs=-1, final_transpose=True,
                slice_arrange=False, slice_arrange_size=100, use_tc=False, **kw):
	"""Reference-compatible dense entry (partial_rwr.py:45-175): x (cells, nb, w) CUDA fp32.
	Returns (imputed, n_iter); imputed is the (nb, w, cells) permuted view when final_transpose."""
	if not x.is_cuda:
		raise _lib.FHError("partial_rwr: CUDA tensor required (no CPU path in fasthigashi_b200)")
	n_iter = 0
	if do_conv or do_rwr:
		c, nb, w = x.shape
		ldw = pad4(w)
		buf = torch.zeros(c, nb, ldw, dtype=torch.float32, device=x.device)
		buf[:, :, :w] = x
		cov = None
		if do_col and do_rwr:
			cov = bin_cov.to(x.device, torch.float32).contiguous()
		it = C.c_int(0)
		d = _lib.rwr_desc(nb, w, ldw, slice_start, force_rwr_epochs, do_conv, do_rwr, do_col, 0, c, 0, use_tc)
		lib = _lib.lib()
		ws = _lib.workspace(lib.fh_rwr_workspace_bytes(C.byref(d)), x.device, "rwr")
		_lib.check(lib.fh_rwr_dense(C.byref(d), buf.data_ptr(), nb * ldw, None if cov is None else cov.data_ptr(),
		                            0 if cov is None else cov.stride(0), ws.data_ptr(), ws.numel(), C.byref(it),
		                            _lib.stream_ptr()))
		n_iter = it.value
		x = buf[:, :, :w]
		if final_transpose:
			x = x.permute(1, 2, 0)
	if return_rwr_iter:
		return x, n_iter
	return x, 0
