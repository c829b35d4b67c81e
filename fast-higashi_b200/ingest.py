"""Raw per-cell contact maps -> the normalised chromosome COO tensors the decomposition consumes
(the caller side of the hot path). Thin ctypes binding of libfh_host.so
(include/fh_host.h, csrc/fh_host.cpp: C++/OpenMP over cells).

Reference entry points covered (FastHigashi_Wrapper.py):
  * `get_qc`                           :428-458
  * `pack_training_data_one_process`   :221-366  (with preprocessing.filter_bin :474-489,
    normalize_per_batch :232-292, norm2 :195-215, normalize_by_coverage :137-142)
  * `preprocess_contact_map`           :368-419  (per-resolution cache file)
Only what the wrapper itself calls is implemented: `fac_size=1`, `is_sym=True`, `force_shift=False`
(:484-494). The scipy matrices are handed over as per-cell base pointers - nothing is copied or
concatenated in Python. There is one implementation: if the library is missing, calls raise.
"""
import ctypes as C
import os
import pickle

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfh_host.so")
EXPORTS = ["fh_host_last_error", "fh_host_version", "fh_host_qc_chrom", "fh_host_pack_chrom", "fh_host_pack_fetch",
           "fh_host_pack_free", "fh_host_block_csr_count", "fh_host_block_csr_fill"]
_I32, _I64, _F32, _F64 = 0, 1, 2, 3
_DTYPES = {np.dtype(np.int32): _I32, np.dtype(np.int64): _I64, np.dtype(np.float32): _F32, np.dtype(np.float64): _F64}


class IngestError(RuntimeError):
	pass


class _Cells(C.Structure):
	_fields_ = [("num_cell", C.c_int64), ("n_row", C.c_int32), ("n_col", C.c_int32), ("indptr", C.c_void_p),
	            ("indices", C.c_void_p), ("data", C.c_void_p), ("index_type", C.c_int32), ("data_type", C.c_int32)]


class _BlockGeom(C.Structure):
	_fields_ = [("num_bin", C.c_int32), ("bs_bin", C.c_int32), ("num_block", C.c_int32), ("nb", C.c_void_p), ("col0", C.c_void_p),
	            ("w", C.c_void_p), ("num_cell", C.c_int64)]


class _PackOpts(C.Structure):
	_fields_ = [("off_diag", C.c_int32), ("merge_row", C.c_int32), ("merge_col", C.c_int32), ("dead_bin", C.c_void_p),
	            ("batch_of_cell", C.c_void_p), ("num_batch", C.c_int32), ("batch_norm", C.c_int32), ("num_threads", C.c_int32)]


_lib = None


def lib():
	global _lib
	if _lib is None:
		if not os.path.exists(LIB_PATH):
			raise IngestError("libfh_host.so not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
		L = C.CDLL(LIB_PATH)
		L.fh_host_last_error.restype = C.c_char_p
		L.fh_host_qc_chrom.argtypes = [C.POINTER(_Cells), C.c_int32, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int32]
		L.fh_host_pack_chrom.argtypes = [C.POINTER(_Cells), C.POINTER(_PackOpts), C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
		                                 C.POINTER(C.c_int32)]
		L.fh_host_pack_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
		L.fh_host_pack_free.argtypes = [C.c_void_p]
		L.fh_host_pack_free.restype = None
		L.fh_host_block_csr_count.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.POINTER(_BlockGeom),
		                                      C.c_void_p, C.c_void_p, C.c_int32]
		L.fh_host_block_csr_fill.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64,
		                                     C.POINTER(_BlockGeom), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
		_lib = L
	return _lib


def _check(rc):
	if rc != 0:
		e = IngestError("libfh_host: %s (code %d)" % (lib().fh_host_last_error().decode(), rc))
		e.code = rc
		raise e


def num_threads():
	return int(os.environ.get("FH_HOST_THREADS", "0"))


class CellMatrices:
	"""Pointer tables over a sequence of per-cell scipy CSR matrices (kept alive by this object)."""

	def __init__(self, mats):
		if len(mats) == 0:
			raise IngestError("no cells")
		keep = []
		first = mats[0]
		self.shape = tuple(int(s) for s in first.shape)
		itype = np.dtype(first.indices.dtype) if getattr(first, "format", "") == "csr" else np.dtype(np.int32)
		dtype = np.dtype(first.dtype)
		if itype not in (np.dtype(np.int32), np.dtype(np.int64)):
			itype = np.dtype(np.int64)
		if dtype not in _DTYPES:
			dtype = np.dtype(np.float64)
		n = len(mats)
		tab = np.empty((3, n), dtype=np.uintp)
		for i, m in enumerate(mats):
			if m.format != "csr":
				m = m.tocsr()
			if tuple(m.shape) != self.shape:
				raise IngestError("cell %d has shape %s, expected %s" % (i, m.shape, self.shape))
			ip, ix, dv = m.indptr, m.indices, m.data
			# scipy arrays are 1-d and contiguous; normalise the rare odd dtype so one type code describes every cell
			if ip.dtype != itype or not ip.flags.c_contiguous: ip = np.ascontiguousarray(ip, dtype=itype)
			if ix.dtype != itype or not ix.flags.c_contiguous: ix = np.ascontiguousarray(ix, dtype=itype)
			if dv.dtype != dtype or not dv.flags.c_contiguous: dv = np.ascontiguousarray(dv, dtype=dtype)
			keep.append((ip, ix, dv))
			tab[0, i] = ip.__array_interface__["data"][0]
			tab[1, i] = ix.__array_interface__["data"][0]
			tab[2, i] = dv.__array_interface__["data"][0]
		self._keep, self._tab = keep, tab
		self.num_cell, self.dtype = n, dtype
		self.desc = _Cells(n, self.shape[0], self.shape[1], tab[0].ctypes.data, tab[1].ctypes.data, tab[2].ctypes.data,
		                   _DTYPES[itype], _DTYPES[dtype])


def load_raw_chrom(raw_dir, chrom, reorder=None, filename_pattern="%s_sparse_adj.npy"):
	a = np.load(os.path.join(raw_dir, filename_pattern % chrom), allow_pickle=True)
	if reorder is not None:
		a = a[reorder]
	return CellMatrices(a)


# ------------------------------------------------------------------------------------------------
def qc_partial(raw_dir, chrom_list, resolution, filename_pattern="%s_sparse_adj.npy"):
	"""The per-chromosome part of get_qc for a SUBSET of the chromosomes: (number of chromosomes on which each cell
	passes (cells,) float64, summed read counts (cells,) float64, dtype of the files). Partial results of disjoint
	subsets add up (the counts are integers, so the sums are exact in any order)."""
	scale = int(1000000 / resolution)
	passed, read_all, dtype = 0, 0, np.float64
	L = lib()
	for chrom in chrom_list:
		cm = load_raw_chrom(raw_dir, chrom, None, filename_pattern)
		contacts = np.empty(cm.num_cell, dtype=np.float64)
		reads = np.empty(cm.num_cell, dtype=np.float64)
		n_bin = C.c_int64(0)
		_check(L.fh_host_qc_chrom(C.byref(cm.desc), scale, contacts.ctypes.data, reads.ctypes.data, C.byref(n_bin), num_threads()))
		n_bin = n_bin.value
		if np.sum(contacts > n_bin) > 0.5 * len(contacts):
			mask = contacts > n_bin
		else:
			mask = contacts > np.quantile(contacts, 0.5)
		passed = passed + mask.astype(np.float64)
		read_all = read_all + reads
		dtype = cm.dtype if np.issubdtype(cm.dtype, np.floating) else np.float64
	return passed, read_all, dtype


def qc_combine(passed, read_all, num_chrom, dtype=np.float64):
	kept = (np.asarray(passed) >= num_chrom).astype("float32")
	return kept, np.log1p(np.asarray(read_all).astype(dtype))  # the reference sums in the files' dtype (fp32 files -> fp32 log1p)


def get_qc(raw_dir, chrom_list, resolution, filename_pattern="%s_sparse_adj.npy"):
	"""FastHigashi_Wrapper.py:428-458 -> (kept float32 (cells,), log1p(total read count) (cells,)).
	A cell is kept when, on every chromosome, its number of distinct contacts (nnz + positive
	diagonal)/2 exceeds the number of well-covered bins (or the median, when fewer than half pass)."""
	passed, read_all, dtype = qc_partial(raw_dir, chrom_list, resolution, filename_pattern)
	return qc_combine(passed, read_all, len(chrom_list), dtype)


def pack_training_data_one_process(raw_dir, chrom, reorder, off_diag, merge_fac_row=1, merge_fac_col=1,
                                   batch_id=None, batch_norm=True, blacklist=None,
                                   filename_pattern="%s_sparse_adj.npy", raw=None):
	"""FastHigashi_Wrapper.py:221-366 -> (indices int32 (3, nnz) [row, col, cell], values fp32 (nnz,),
	shape (n_valid, n_valid, cells)). `batch_id` (cells,) in `reorder` order switches the bulk to the
	sum of per-batch bulks (:276-279) and, with batch_norm, applies the per-batch normalisation.
	`blacklist`: {chrom: bin ids} (the reference reads raw/blacklist.npy, :236-252). `raw`: a
	CellMatrices (or a sequence of scipy matrices) instead of the file.
	One deliberate difference: contacts that collide when bins are coarsened are summed (the
	reference's `sum_duplicates()` at :260 is a no-op on the COO objects it builds)."""
	cm = raw if isinstance(raw, CellMatrices) else CellMatrices(raw) if raw is not None else \
		load_raw_chrom(raw_dir, chrom, reorder, filename_pattern)
	n0, n1 = cm.shape
	opts = _PackOpts(int(off_diag), int(merge_fac_row), int(merge_fac_col), None, None, 0, int(bool(batch_norm)), num_threads())
	keep = []
	if blacklist is not None:
		bl = np.asarray(blacklist[chrom]).astype(np.int64)
		bl = bl[bl < n0]
		dead = np.zeros(max(n0, n1), dtype=np.uint8)
		dead[bl] = 1
		keep.append(dead)
		opts.dead_bin = dead.ctypes.data
	if batch_id is not None:
		names, codes = np.unique(np.asarray(batch_id), return_inverse=True)  # sorted, as the reference iterates them
		if len(codes) != cm.num_cell:
			raise IngestError("batch_id has %d entries for %d cells" % (len(codes), cm.num_cell))
		codes = np.ascontiguousarray(codes, dtype=np.int32)
		keep.append(codes)
		opts.batch_of_cell, opts.num_batch = codes.ctypes.data, len(names)
	L = lib()
	handle, nnz, num_bins = C.c_void_p(None), C.c_int64(0), C.c_int32(0)
	_check(L.fh_host_pack_chrom(C.byref(cm.desc), C.byref(opts), C.byref(handle), C.byref(nnz), C.byref(num_bins)))
	try:
		indices = np.empty((3, nnz.value), dtype=np.int32)
		values = np.empty(nnz.value, dtype=np.float32)
		_check(L.fh_host_pack_fetch(handle, indices.ctypes.data, values.ctypes.data))
	finally:
		L.fh_host_pack_free(handle)
	return indices, values, (int(num_bins.value), int(num_bins.value), cm.num_cell)


def load_blacklist(temp_dir):
	try:
		return np.load(os.path.join(temp_dir, "raw", "blacklist.npy"), allow_pickle=True).item()
	except Exception:
		return None


def preprocess_contact_map(config, reorder, path2input_cache, off_diag, res, batch_id=None, batch_norm=True):
	"""FastHigashi_Wrapper.py:368-419: all chromosomes of one resolution, cached next to the
	reference's cache (own file: plain arrays, no class pickles). Returns [(indices, values, shape)]."""
	if path2input_cache is not None and os.path.exists(path2input_cache):
		with open(path2input_cache, "rb") as f:
			return pickle.load(f)
	raw_dir = os.path.join(config["temp_dir"], "raw")
	fac = int(res / config["resolution"])
	bl = load_blacklist(config["temp_dir"])
	out = [pack_training_data_one_process(raw_dir, chrom, reorder, off_diag, fac, fac, batch_id, batch_norm, bl)
	       for chrom in config["chrom_list"]]
	if path2input_cache is not None:
		with open(path2input_cache, "wb") as f:
			pickle.dump(out, f, protocol=4)
	return out


# ------------------------------------------------------------------------------------------------
def block_csr(indices, values, num_bin, bs_bin, num_cell, nb, col0, w):
	"""COO chromosome tensor -> the per-bin-block CSR arrays of `sparse_for_schic.Chrom_Dataset`
	(replaces the reference's bin-block x cell-batch split, sparse_for_schic.py:440-499).
	`indices`: (3, nnz) int32/int64 numpy [row, col, cell], any order; `values`: (nnz,) float32;
	`nb`, `col0`, `w`: per-block rows, first window column, window width.
	Returns ([rowptr int32], [col int16], [val float32]) as numpy arrays, one per block.
	Raises IngestError; `.code` is the library's (FH_HOST_EWINDOW = -4, FH_HOST_EDUP = -5)."""
	indices = np.asarray(indices)
	if indices.ndim != 2 or indices.shape[0] != 3:
		raise IngestError("indices must be (3, nnz)")
	if indices.dtype not in (np.dtype(np.int32), np.dtype(np.int64)):
		indices = indices.astype(np.int64)
	rows = [np.ascontiguousarray(indices[i]) for i in range(3)]  # views when `indices` is C-contiguous
	values = np.ascontiguousarray(values, dtype=np.float32)
	nnz = int(values.shape[0])
	if indices.shape[1] != nnz:
		raise IngestError("indices must be (3, nnz)")
	nb = np.ascontiguousarray(nb, dtype=np.int32)
	col0 = np.ascontiguousarray(col0, dtype=np.int32)
	w = np.ascontiguousarray(w, dtype=np.int32)
	B = len(nb)
	geom = _BlockGeom(int(num_bin), int(bs_bin), B, nb.ctypes.data, col0.ctypes.data, w.ctypes.data, int(num_cell))
	itype = _DTYPES[indices.dtype]
	L = lib()
	rowptr = [np.empty(int(num_cell) * int(n) + 1, dtype=np.int32) for n in nb]
	rp_tab = np.array([a.ctypes.data for a in rowptr], dtype=np.uintp)
	nnz_block = np.zeros(B, dtype=np.int64)
	ptrs = [a.ctypes.data if nnz else None for a in rows]
	_check(L.fh_host_block_csr_count(ptrs[0], ptrs[1], ptrs[2], itype, nnz, C.byref(geom), rp_tab.ctypes.data,
	                                 nnz_block.ctypes.data, num_threads()))
	col = [np.empty(int(n), dtype=np.int16) for n in nnz_block]
	val = [np.empty(int(n), dtype=np.float32) for n in nnz_block]
	c_tab = np.array([a.ctypes.data for a in col], dtype=np.uintp)
	v_tab = np.array([a.ctypes.data for a in val], dtype=np.uintp)
	_check(L.fh_host_block_csr_fill(ptrs[0], ptrs[1], ptrs[2], itype, values.ctypes.data if nnz else None, nnz, C.byref(geom),
	                                rp_tab.ctypes.data, c_tab.ctypes.data, v_tab.ctypes.data, num_threads()))
	return rowptr, col, val
