"""TEST INFRASTRUCTURE ONLY - numpy restatement of the reference's ingest stage, the checker for
fast-higashi_b200/ingest.py (libfh_host.so). Nothing in the product path may import this module.
PINNED: tests/test_ingest_golden.py checks it against tests/golden/ingest_cases.npz, which the
unmodified reference produced (tests/golden/make_golden_ingest.py).

Reference behaviour restated (FastHigashi_Wrapper.py):
  * `get_qc`                           :428-458   per-cell QC mask + log1p read counts
  * `pack_training_data_one_process`   :221-366   blacklist, bin merging to the decomposition
    resolution, |col-row| <= off_diag filter, bulk / valid-bin map (`preprocessing.filter_bin`
    :474-489), optional per-batch normalisation (`preprocessing.normalize_per_batch` :232-292 +
    `norm2` :195-215), per-cell coverage normalisation (`normalize_by_coverage` :137-142),
    log1p, clip at mean + 15 sigma
Every chromosome is flattened into four flat arrays (row, col, cell, value); each step is a
whole-array numpy operation. Entries keep the reference's order (cell-major, row-major inside a
cell). Only `fac_size=1`, `is_sym=True`, `force_shift=False` (:484-494).
"""
import os

import numpy as np


class RawChrom:
	"""One chromosome of `raw/{chrom}_sparse_adj.npy` flattened: row/col int64, cell int64 (index in
	the order given), data float64, shape (n_row, n_col), num_cell."""
	__slots__ = ("row", "col", "cell", "data", "shape", "num_cell", "dtype")

	def __init__(self, row, col, cell, data, shape, num_cell, dtype):
		self.row, self.col, self.cell, self.data = row, col, cell, data
		self.shape, self.num_cell, self.dtype = tuple(int(s) for s in shape), int(num_cell), dtype


def flatten_cells(mats):
	"""List/object-array of per-cell scipy sparse matrices -> RawChrom (entries stay in CSR order)."""
	indptr, indices, data = [], [], []
	shape = None
	for m in mats:
		if m.format != "csr":
			m = m.tocsr()
		if shape is None:
			shape = m.shape
		indptr.append(np.asarray(m.indptr, dtype=np.int64))
		indices.append(np.asarray(m.indices))
		data.append(np.asarray(m.data))
	num_cell = len(indptr)
	if num_cell == 0:
		raise ValueError("no cells")
	nnz_per_cell = np.array([p[-1] for p in indptr], dtype=np.int64)
	col = np.concatenate(indices).astype(np.int64)
	val = np.concatenate(data)
	cell = np.repeat(np.arange(num_cell, dtype=np.int64), nnz_per_cell)
	per_row = np.concatenate([np.diff(p) for p in indptr])
	row = np.repeat(np.tile(np.arange(shape[0], dtype=np.int64), num_cell), per_row)
	return RawChrom(row, col, cell, val.astype(np.float64), shape, num_cell, val.dtype)


def load_raw_chrom(raw_dir, chrom, reorder=None, filename_pattern="%s_sparse_adj.npy"):
	a = np.load(os.path.join(raw_dir, filename_pattern % chrom), allow_pickle=True)
	if reorder is not None:
		a = a[reorder]
	return flatten_cells(a)


def _bulk(rc, mask=None):
	n0, n1 = rc.shape
	key = rc.row * n1 + rc.col
	w = rc.data
	if mask is not None:
		key, w = key[mask], w[mask]
	return np.bincount(key, weights=w, minlength=n0 * n1).reshape(n0, n1)


# ------------------------------------------------------------------------------------------------
def get_qc(raw_dir, chrom_list, resolution, filename_pattern="%s_sparse_adj.npy"):
	"""FastHigashi_Wrapper.py:428-458 -> (kept float32 (cells,), log1p(total read count) (cells,)).
	A cell is kept when, on every chromosome, its number of distinct contacts
	(nnz + nonzero diagonal)/2 exceeds the number of well-covered bins (or the median, when fewer
	than half of the cells would pass)."""
	scale = int(1000000 / resolution)
	masks, read_all, dtype = [], 0, np.float64
	for chrom in chrom_list:
		rc = load_raw_chrom(raw_dir, chrom, None, filename_pattern)
		bulk = _bulk(rc)
		cov = np.sum(bulk > 0, axis=-1)
		n_bin = np.sum(cov > 0.1 * cov.shape[0] * scale)
		nnz = np.bincount(rc.cell, minlength=rc.num_cell).astype(np.float64)
		# m.diagonal() sums duplicate entries; canonical CSR has none, bincount covers both
		dkey = rc.cell[rc.row == rc.col] * rc.shape[0] + rc.row[rc.row == rc.col]
		dsum = np.bincount(dkey, weights=rc.data[rc.row == rc.col], minlength=rc.num_cell * rc.shape[0])
		diag_pos = (dsum.reshape(rc.num_cell, rc.shape[0]) > 0).sum(1)
		contacts = (nnz + diag_pos) / 2
		reads = np.bincount(rc.cell, weights=rc.data, minlength=rc.num_cell)
		if np.sum(contacts > n_bin) > 0.5 * len(contacts):
			masks.append(contacts > n_bin)
		else:
			masks.append(contacts > np.quantile(contacts, 0.5))
		read_all = read_all + reads
		dtype = rc.dtype if np.issubdtype(rc.dtype, np.floating) else np.float64
	kept = (np.sum(np.array(masks).astype("float"), axis=0) >= len(chrom_list)).astype("float32")
	return kept, np.log1p(read_all.astype(dtype))  # the reference sums in the raw dtype (fp32 files -> fp32 log1p)


def filter_bin(bulk):
	"""preprocessing.py:474-489 (is_sym): bins with any coverage are valid; old -> new index, -1 = dropped."""
	c = bulk.sum(1)
	v = c > min(0.0, 0.01 * bulk.shape[1])
	m = np.cumsum(v) - 1
	m[~v] = -1
	return m, int(v.sum()), v


def _diag_profile(m, length):
	"""sum of the k-th diagonal (both triangles for k > 0), k < length (preprocessing.py:246-251)."""
	out = np.zeros(length)
	for k in range(min(length, m.shape[0])):
		out[k] = np.diagonal(m, k).sum() * (1 if k == 0 else 2)
	return out


def normalize_per_batch(rc, bulk, batch_bulk, batch_of_cell, off_diag):
	"""preprocessing.py:232-292 + norm2 :195-215 on the flat arrays: every contact is rescaled by the
	ratio of the pooled to its batch's coverage (sqrt row x sqrt col) and divided by the batch's
	relative contact-vs-distance profile. `off_diag` = number of diagonals profiled."""
	bk_cov = bulk.sum(axis=-1)
	nb = bulk / (np.sqrt(bk_cov[None]) + 1e-15) / (np.sqrt(bk_cov[:, None]) + 1e-15)
	bulk_ratio = _diag_profile(nb, off_diag) / nb.sum()
	names = list(batch_bulk.keys())
	ratio = np.zeros((len(names), off_diag))
	cov = np.zeros((len(names), bulk.shape[0]))
	for i, b in enumerate(names):
		m = batch_bulk[b]
		m_cov = m.sum(axis=-1)
		mn = m / (np.sqrt(m_cov[None]) + 1e-15) / (np.sqrt(m_cov[:, None]) + 1e-15)
		ratio[i] = _diag_profile(mn, off_diag) / (mn.sum() + 1e-15) / (bulk_ratio + 1e-15)
		cov[i] = m_cov
	lut = {b: i for i, b in enumerate(names)}
	bidx = np.array([lut[b] for b in batch_of_cell], dtype=np.int64)[rc.cell]
	d = np.abs(rc.row - rc.col)
	data = rc.data / (np.sqrt(cov[bidx, rc.row]) * np.sqrt(cov[bidx, rc.col])) * (np.sqrt(bk_cov[rc.row]) * np.sqrt(bk_cov[rc.col]))
	rc.data = data / (ratio[bidx, d] + 1e-15)


def pack_training_data_one_process(raw_dir, chrom, reorder, off_diag, merge_fac_row=1, merge_fac_col=1,
                                   batch_id=None, batch_norm=True, blacklist=None,
                                   filename_pattern="%s_sparse_adj.npy", raw=None):
	"""FastHigashi_Wrapper.py:221-366 -> (indices int32 (3, nnz) [row, col, cell], values fp32 (nnz,),
	shape (n_valid, n_valid, cells)). `batch_id` (cells,) in `reorder` order switches the bulk to the
	sum of per-batch bulks (:276-279) and, with batch_norm, applies the per-batch normalisation.
	`blacklist`: {chrom: bin ids} (the reference reads raw/blacklist.npy, :236-252)."""
	rc = raw if raw is not None else load_raw_chrom(raw_dir, chrom, reorder, filename_pattern)
	n0, n1 = rc.shape
	if blacklist is not None:
		bl = np.asarray(blacklist[chrom])
		bl = bl[bl < n0]
		dead = np.zeros(max(n0, n1), dtype=bool)
		dead[bl] = True
		keep = ~(dead[rc.row] | dead[rc.col]) & (rc.data != 0)  # csr_matrix(dense) drops zeros (:251)
		rc.row, rc.col, rc.cell, rc.data = rc.row[keep], rc.col[keep], rc.cell[keep], rc.data[keep]
	if merge_fac_row > 1 or merge_fac_col > 1:
		# :256-261 - coarsen the bins and sum what collides inside a cell
		rc.row //= merge_fac_row
		rc.col //= merge_fac_col
		n0, n1 = int(np.ceil(n0 / merge_fac_col)), int(np.ceil(n1 / merge_fac_row))
		key = (rc.cell * n0 + rc.row) * n1 + rc.col
		uniq, inv = np.unique(key, return_inverse=True)
		rc.data = np.bincount(inv, weights=rc.data, minlength=len(uniq))
		rc.col = uniq % n1
		rc.row = (uniq // n1) % n0
		rc.cell = uniq // (n0 * n1)
		rc.shape = (n0, n1)
	keep = np.abs(rc.col - rc.row) <= off_diag  # :264-268
	rc.row, rc.col, rc.cell, rc.data = rc.row[keep], rc.col[keep], rc.cell[keep], rc.data[keep]

	if batch_id is not None:
		batch_id = np.asarray(batch_id)
		batch_bulk = {b: _bulk(rc, (batch_id == b)[rc.cell]) for b in np.unique(batch_id)}
		bulk = 0
		for b in batch_bulk:
			bulk = bulk + batch_bulk[b]
	else:
		batch_bulk = None
		bulk = _bulk(rc)
	mapping, num_bins, _ = filter_bin(bulk / rc.num_cell)
	if batch_bulk is not None and batch_norm:
		normalize_per_batch(rc, bulk, batch_bulk, batch_id, off_diag + 1)
	# per-cell coverage normalisation (preprocessing.py:137-142): data *= n_rows / (sum + 1e-15)
	total = np.bincount(rc.cell, weights=rc.data, minlength=rc.num_cell)
	rc.data = rc.data * (rc.shape[0] / (total + 1e-15))[rc.cell]
	if np.isnan(rc.data).any():
		raise AssertionError("NaN after normalisation (%s)" % chrom)
	row, col = mapping[rc.row], mapping[rc.col]
	keep = (row != -1) & (col != -1)
	indices = np.ascontiguousarray(np.stack([row[keep], col[keep], rc.cell[keep]]).astype(np.int32))
	values = np.log1p(rc.data[keep].astype(np.float32))
	shape = (num_bins, num_bins, rc.num_cell)
	if indices.shape[1]:
		assert indices.min() >= 0 and (indices.max(1) < np.asarray(shape)).all()
		mean_, std_ = np.mean(values), np.std(values)
		values = np.clip(values, a_min=None, a_max=mean_ + 15 * std_)
	return indices, np.ascontiguousarray(values), shape
