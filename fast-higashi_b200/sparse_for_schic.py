"""Host-side mirror of the reference's sparse staging layer, re-designed for a device-resident
block-CSR layout.

Reference interface mirrored (names, argument meaning, attribute names):
  * `Sparse`         sparse_for_schic.py:58-275   (N-d COO + indptr on dim 0; only what the hot
                                                   path's callers use: construction, sort_indices,
                                                   get_slice_idx_value)
  * `Chrom_Dataset`  sparse_for_schic.py:356-510  (bin-block x cell-batch split with compact
                                                   +-flank windows) and `.fetch` :588-613

Layout difference (DESIGN.md "data layout"): the reference keeps one pinned COO object
(`Fake_Sparse`, :322-353: int16 row, int16 col, int32 cell, fp32 value = 12 B/nnz) per
(bin-block, cell-batch) and re-uploads it on every fetch. Here each bin-block of a chromosome
is ONE CSR over (cell, local row): `rowptr` int32 [(cells*nb)+1], `col` int16 [nnz] (window
local), `val` fp32 [nnz] = 6 B/nnz, uploaded once and kept in HBM. Good-QC cells come first,
bad-QC cells follow (same order as the reference, :422-436).
"""
import math
from collections import namedtuple
import numpy as np
import torch

BlockGeom = namedtuple("BlockGeom", "row0 nb col0 w s e")


def block_geometry(num_bin, bs_bin, flank, compact=True):
	"""Window geometry of every bin-block; follows sparse_for_schic.py:457-492.
	row0/nb: global rows of the block; col0/w: global first column and width of its window;
	[s, e): where the diagonal (nb x nb) block sits inside the window."""
	geoms = []
	for i in range(0, num_bin, bs_bin):
		nb = min(bs_bin, num_bin - i)
		if not compact:
			geoms.append(BlockGeom(i, nb, 0, num_bin, i, i + nb))
			continue
		right = flank if num_bin - i - nb - flank > 0 else num_bin - i - nb
		if i > flank:
			col0, s = i - flank, flank
		else:
			col0, s = 0, i
		w = (i - col0) + nb + right
		geoms.append(BlockGeom(i, nb, col0, w, s, s + nb))
	return geoms


class Sparse:
	"""Minimal COO container with the reference's constructor (sparse_for_schic.py:59-70)."""

	def __init__(self, indices, values, shape, indptr=None, copy=True, verbose=False):
		self.indices = torch.as_tensor(np.asarray(indices) if not torch.is_tensor(indices) else indices)
		self.values = torch.as_tensor(np.asarray(values) if not torch.is_tensor(values) else values)
		if copy:
			self.indices = self.indices.clone()
			self.values = self.values.clone()
		self.shape = np.asarray(shape).astype(np.int64)
		self.ndim = len(self.shape)
		self.indptr = indptr
		if tuple(self.indices.shape) != (self.ndim, len(self.values)):
			raise AssertionError("indices must be (ndim, nnz)")
		if self.indices.numel():
			mx = self.indices.max(1).values.cpu().numpy()
			if (self.indices.min() < 0) or (mx >= self.shape).any():
				raise AssertionError((mx, self.shape))

	def sort_indices(self, dim=0, force=False):
		assert dim == 0
		if self.indptr is not None and not force:
			return
		order = torch.argsort(self.indices[0].long(), stable=True)
		self.indices = self.indices[:, order]
		self.values = self.values[order]
		cnt = torch.bincount(self.indices[0].long(), minlength=int(self.shape[0]))
		self.indptr = torch.cat([cnt.new_zeros(1), torch.cumsum(cnt, 0)])

	def get_slice_idx_value(self, idx, dim=0, device="cpu"):
		assert dim == 0 and isinstance(idx, slice)
		self.sort_indices()
		start = idx.start or 0
		stop = min(idx.stop if idx.stop is not None else int(self.shape[0]), int(self.shape[0]))
		lo, hi = int(self.indptr[start]), int(self.indptr[stop])
		return self.indices[:, lo:hi], self.values[lo:hi], (stop - start,) + tuple(self.shape[1:]), start

	def __len__(self):
		return int(self.shape[0])

	# -- the container operations of the reference class (sparse_for_schic.py:96-275), which its own `test()` (:634-661)
	#    exercises; none of them is on the decomposition's path, they are here so that code written against the
	#    reference's `Sparse` keeps working
	def numel(self):
		return int(np.prod(self.shape))

	def _flat(self):
		"""Row-major linear index of every entry."""
		mult = torch.as_tensor(np.concatenate([np.cumprod(self.shape[::-1])[::-1][1:], [1]]).astype(np.int64))
		return (self.indices.long() * mult[:, None]).sum(0)

	def permute(self, *dims, inplace=False):
		if tuple(sorted(dims)) != tuple(range(self.ndim)):
			raise AssertionError("dims must be a permutation of range(ndim)")
		d = list(dims)
		indices, shape = self.indices[d], self.shape[d]
		indptr = self.indptr if d[0] == 0 else None
		if inplace:
			self.indices, self.shape, self.indptr = indices, shape, indptr
			return self
		return Sparse(indices, self.values, shape, indptr=indptr)

	def reshape(self, *dims, inplace=False):
		dims = np.asarray(dims, dtype=np.int64).copy()
		total = int(np.prod(self.shape))
		if (dims == -1).sum() > 1:
			raise AssertionError(dims)
		if (dims == -1).any():
			known = int(-np.prod(dims))
			if known == 0 or total % known:
				raise AssertionError(dims)
			dims[dims == -1] = total // known
		if int(np.prod(dims)) != total:
			raise AssertionError((self.shape, dims))
		flat = self._flat()
		rows = []
		for n in dims[::-1]:
			rows.append(flat % int(n))
			flat = torch.div(flat, int(n), rounding_mode="floor")
		indices = torch.stack(rows[::-1]).to(self.indices.dtype if self.indices.dtype == torch.int64 else torch.int64)
		if inplace:
			self.indices, self.shape, self.ndim, self.indptr = indices, dims, len(dims), None
			return self
		return Sparse(indices, self.values, dims)

	def slicing(self, idx, dim=0):
		assert dim == 0 and isinstance(idx, slice) and idx.step in (None, 1)
		indices, values, shape, start = self.get_slice_idx_value(idx)
		indices = indices.clone()
		indices[0] -= start
		return Sparse(indices, values, shape)

	def indexing(self, idx, dim=0):
		assert dim == 0
		self.sort_indices()
		lo, hi = int(self.indptr[idx]), int(self.indptr[idx + 1])
		return Sparse(self.indices[1:, lo:hi], self.values[lo:hi], tuple(self.shape[1:]))

	def __getitem__(self, item):
		if isinstance(item, (int, np.integer)):
			return self.indexing(int(item))
		if isinstance(item, slice):
			return self.slicing(item)
		raise NotImplementedError(type(item))

	def filter_max_distance(self, max_distance=100):
		keep = (self.indices[1].long() - self.indices[0].long()).abs() <= max_distance
		self.indices, self.values, self.indptr = self.indices[:, keep], self.values[keep], None

	def to_dense(self):
		out = torch.zeros(tuple(int(x) for x in self.shape), dtype=self.values.dtype)
		out.view(-1).index_put_((self._flat(),), self.values, accumulate=False)
		return out

	def to_scipy(self):
		from scipy.sparse import coo_matrix
		assert self.ndim == 2
		return coo_matrix((self.values.numpy(), (self.indices[0].numpy(), self.indices[1].numpy())), tuple(int(x) for x in self.shape))

	def to_csr(self):
		return self.to_scipy().tocsr()


class Chrom_Dataset:
	"""One chromosome at one resolution as block-CSR (see module docstring).

	Constructor signature = reference's (sparse_for_schic.py:357-358). `tensor` is a `Sparse`
	(ours or the reference's: anything with `.indices (3,nnz)`, `.values`, `.shape`).
	`bs_cell` keeps the reference's meaning only where it changes results: the RWR auto-stop
	of `init_params` is a max over the cells of one cell batch (partial_rwr.py:120-123).
	"""

	def __init__(self, tensor, bs_bin, bs_cell, good_qc_num=-1, kind="hic", upper_sim=False,
	             compact=False, flank=0, chrom="chr1", resolution=10000, device=None):
		if kind != "hic":
			raise NotImplementedError("only kind='hic' is on the hot path")
		if upper_sim:
			raise NotImplementedError("upper_sim=True is never used by the wrapper (FastHigashi_Wrapper.py:531)")
		shape = [int(x) for x in tensor.shape]
		self.resolution = resolution
		self.chrom = chrom
		self.length = shape[0]
		self.num_bin = shape[0]
		self.total_cell_num = shape[-1]
		self.num_cell = self.total_cell_num if good_qc_num == -1 else int(good_qc_num)
		self.bs_bin = int(bs_bin)
		self.bs_cell = int(bs_cell)
		self.kind = kind
		self.upper_sim = upper_sim
		self.compact = compact
		self.flank = int(flank)
		self.geoms = block_geometry(self.num_bin, self.bs_bin, self.flank, compact)
		self.bin_slice_list = [slice(g.row0, g.row0 + g.nb) for g in self.geoms]
		self.local_bin_slice_list = [slice(g.s, g.e) for g in self.geoms]
		self.col_bin_slice_list = [slice(g.col0, g.col0 + g.w) for g in self.geoms]
		good = [slice(c, min(c + self.bs_cell, self.num_cell)) for c in range(0, self.num_cell, self.bs_cell)]
		bad = [slice(c, min(c + self.bs_cell, self.total_cell_num))
		       for c in range(self.num_cell, self.total_cell_num, self.bs_cell)]
		self.cell_slice_list = good + bad
		self.num_bin_batch = len(self.geoms)
		self.num_cell_batch = len(good)
		self.num_cell_batch_bad = len(bad)
		self.shape = [self.num_bin, self.bs_bin + 2 * self.flank, self.num_cell] if compact \
			else [self.num_bin, shape[1], self.num_cell]
		self.global_slice_bin = slice(0, self.num_bin)
		self._build(tensor, device)

	# -- construction -------------------------------------------------------------------------
	def _build(self, tensor, device):
		idx = tensor.indices
		val = tensor.values
		if not torch.is_tensor(idx): idx = torch.as_tensor(np.ascontiguousarray(idx))
		if not torch.is_tensor(val): val = torch.as_tensor(np.ascontiguousarray(val))
		dev = torch.device(device) if device is not None else idx.device
		if idx.device.type == "cpu":
			self._build_host(idx, val, dev)
		else:
			self._build_sort(idx, val, dev)

	def _build_sort(self, idx, val, dev):
		"""COO already on the device: one device sort of 64-bit keys."""
		idx = idx.to(dev)
		val = val.to(dev, torch.float32)
		C = self.total_cell_num
		nblk = len(self.geoms)
		col0 = torch.tensor([g.col0 for g in self.geoms], device=dev, dtype=torch.int64)
		wid = torch.tensor([g.w for g in self.geoms], device=dev, dtype=torch.int64)
		nbs = torch.tensor([g.nb for g in self.geoms], device=dev, dtype=torch.int64)
		wmax = int(wid.max()) if nblk else 1
		row, col, cell = idx[0].long(), idx[1].long(), idx[2].long()
		blk = torch.div(row, self.bs_bin, rounding_mode="floor")
		lrow = row - blk * self.bs_bin
		lcol = col - col0[blk]
		if lcol.numel() and (bool((lcol < 0).any()) or bool((lcol >= wid[blk]).any())):
			raise ValueError("%s: contact outside the +-flank window (|col-row| > flank=%d); filter "
			                 "with off_diag first (FastHigashi_Wrapper.py:265-269)" % (self.chrom, self.flank))
		key = ((blk * C + cell) * self.bs_bin + lrow) * wmax + lcol
		key, order = torch.sort(key)
		if key.numel() > 1 and bool((key[1:] == key[:-1]).any()):
			raise ValueError("%s: duplicate (row, col, cell) entries; sum duplicates first" % self.chrom)
		lrow, lcol, cell, blk, val = lrow[order], lcol[order], cell[order], blk[order], val[order]
		bounds = torch.searchsorted(blk, torch.arange(nblk + 1, device=dev))
		self.rowptr, self.col, self.val = [], [], []
		for b, g in enumerate(self.geoms):
			lo, hi = int(bounds[b]), int(bounds[b + 1])
			rid = cell[lo:hi] * g.nb + lrow[lo:hi]
			cnt = torch.bincount(rid, minlength=C * g.nb)
			rp = torch.zeros(C * g.nb + 1, device=dev, dtype=torch.int64)
			torch.cumsum(cnt, 0, out=rp[1:])
			if hi - lo >= 2 ** 31:
				raise ValueError("block nnz exceeds int32")
			self.rowptr.append(rp.int())
			self.col.append(lcol[lo:hi].short())
			self.val.append(val[lo:hi].clone())  # own allocation: the kernels need 16-byte aligned bases
		self.device = dev

	def _build_host(self, idx, val, dev):
		"""COO in host memory: two counting passes in libfh_host.so (include/fh_host.h `fh_host_block_csr_*`, OpenMP),
		then one upload per array. Same arrays, bit for bit, as the device-sort route of `_build`."""
		from . import ingest
		g = self.geoms
		try:
			rowptr, col, vals = ingest.block_csr(idx.numpy(), val.to(torch.float32).numpy(), self.num_bin, self.bs_bin,
			                                     self.total_cell_num, [x.nb for x in g], [x.col0 for x in g], [x.w for x in g])
		except ingest.IngestError as e:
			code = getattr(e, "code", 0)
			if code == -4:
				raise ValueError("%s: contact outside the +-flank window (|col-row| > flank=%d); filter "
				                 "with off_diag first (FastHigashi_Wrapper.py:265-269)" % (self.chrom, self.flank)) from e
			if code == -5:
				raise ValueError("%s: duplicate (row, col, cell) entries; sum duplicates first (%s)" % (self.chrom, e)) from e
			raise
		self.rowptr = [torch.from_numpy(a).to(dev) for a in rowptr]
		self.col = [torch.from_numpy(a).to(dev) for a in col]
		self.val = [torch.from_numpy(a).to(dev) for a in vals]
		self.device = dev

	@classmethod
	def from_reference(cls, ds, device=None):
		"""Re-stage a reference `Chrom_Dataset` (duck-typed: tensor_list/bad_tensor_list of
		`Fake_Sparse` with +1-offset indices, sparse_for_schic.py:322-353,499) as block-CSR."""
		rows, cols, cells, vals = [], [], [], []
		for b in range(ds.num_bin_batch):
			r0 = ds.bin_slice_list[b].start
			c0 = ds.col_bin_slice_list[b].start or 0
			lists = list(ds.tensor_list[b]) + list(ds.bad_tensor_list[b])
			for fs, sl in zip(lists, ds.cell_slice_list):
				rows.append(fs.indices[0].long() - 1 + r0)
				cols.append(fs.indices[1].long() - 1 + c0)
				cells.append(fs.indices[2].long() + sl.start)
				vals.append(fs.values)
		idx = torch.stack([torch.cat(rows), torch.cat(cols), torch.cat(cells)])
		t = Sparse(idx, torch.cat(vals), (ds.num_bin, ds.num_bin, ds.total_cell_num), copy=False)
		return cls(t, ds.bs_bin, ds.bs_cell, good_qc_num=ds.num_cell, kind="hic", upper_sim=False,
		           compact=ds.compact, flank=ds.flank, chrom=ds.chrom, resolution=ds.resolution,
		           device=device)

	# -- access -------------------------------------------------------------------------------
	def __len__(self):
		return self.length

	def to(self, device):
		dev = torch.device(device)
		self.rowptr = [t.to(dev) for t in self.rowptr]
		self.col = [t.to(dev) for t in self.col]
		self.val = [t.to(dev) for t in self.val]
		self.device = dev
		return self

	def pin_memory(self):
		# reference API (sparse_for_schic.py:576); the block-CSR is uploaded once, pinning only
		# speeds that single copy
		if self.device.type == "cpu" and torch.cuda.is_available():
			self.rowptr = [t.pin_memory() for t in self.rowptr]
			self.col = [t.pin_memory() for t in self.col]
			self.val = [t.pin_memory() for t in self.val]
		return self

	def nnz(self):
		return int(sum(v.numel() for v in self.val))

	def to_payload(self):
		"""Everything needed to rebuild this dataset in another process: plain fields + the CSR arrays on the host."""
		d = {k: v for k, v in self.__dict__.items() if k not in ("rowptr", "col", "val", "device")}
		d["rowptr"] = [t.cpu() for t in self.rowptr]
		d["col"] = [t.cpu() for t in self.col]
		d["val"] = [t.cpu() for t in self.val]
		return d

	@classmethod
	def from_payload(cls, payload, device="cpu"):
		new = object.__new__(cls)
		new.__dict__.update(payload)
		new.device = torch.device("cpu")
		return new.to(device)

	def fetch(self, bin_id, cell_id, save_context=None, transpose=False, good_qc=True, **kwargs):
		"""Reference API (sparse_for_schic.py:588-613): the dense block of bin block `bin_id` and cell batch `cell_id`
		(index into the good-QC batches, or into the bad-QC batches with good_qc=False), floor 1e-8, on the device:
		(cells, nb, w) when `transpose` else the (nb, w, cells) view. Returns ((tensor, [seconds]), kind).
		The decomposition itself never densifies separately (fh_rwr_batched reads the block-CSR)."""
		import time
		from .partial_rwr import densify_block
		t = time.perf_counter()
		sl = self.cell_slice_list[cell_id if good_qc else self.num_cell_batch + cell_id]
		g = self.geoms[bin_id]
		x = densify_block(self, bin_id, sl.start, sl.stop - sl.start)[:, :, :g.w]
		if not transpose:
			x = x.permute(1, 2, 0)
		return (x, [time.perf_counter() - t]), self.kind

	def fetch_bad(self, bin_id, cell_id, **kwargs):
		return self.fetch(bin_id, cell_id, good_qc=False, **kwargs)

	def norm(self):
		"""Frobenius norm of all stored values (sparse_for_schic.py:615-620: good-QC cells only)."""
		total = 0.0
		for b, g in enumerate(self.geoms):
			hi = int(self.rowptr[b][self.num_cell * g.nb])
			total += float(self.val[b][:hi].double().square().sum())
		return math.sqrt(total)

	def cell_range_csr(self, b, cell_start, cell_stop):
		"""CSR of cells [cell_start, cell_stop) of block b, rowptr rebased to 0 (host helper)."""
		g = self.geoms[b]
		rp = self.rowptr[b][cell_start * g.nb: cell_stop * g.nb + 1].long()
		lo, hi = int(rp[0]), int(rp[-1])
		return (rp - lo).int(), self.col[b][lo:hi], self.val[b][lo:hi]

	def select_cells(self, cell_start, cell_stop, good_qc_num=None):
		"""A dataset holding only cells [cell_start, cell_stop) (cell-slab sharding, §8e)."""
		return self.select_cell_ranges([(cell_start, cell_stop)], good_qc_num)

	def select_cell_ranges(self, ranges, good_qc_num=None):
		"""A dataset holding the cells of the given [start, stop) ranges, concatenated in that order (a rank's slab of
		good-QC cells followed by its slab of bad-QC cells). `good_qc_num`: how many of them are good (default: all)."""
		new = object.__new__(Chrom_Dataset)
		new.__dict__.update(self.__dict__)
		n = int(sum(b - a for a, b in ranges))
		new.total_cell_num = n
		new.num_cell = n if good_qc_num is None else int(good_qc_num)
		new.rowptr, new.col, new.val = [], [], []
		for b in range(len(self.geoms)):
			parts = [self.cell_range_csr(b, lo, hi) for lo, hi in ranges if hi > lo]
			if len(parts) == 1:
				rp, c, v = parts[0]
				new.rowptr.append(rp.clone()); new.col.append(c.clone()); new.val.append(v.clone())
				continue
			if not parts:
				dev = self.rowptr[b].device
				new.rowptr.append(torch.zeros(1, dtype=torch.int32, device=dev))
				new.col.append(self.col[b][:0].clone()); new.val.append(self.val[b][:0].clone())
				continue
			rps, off = [parts[0][0][:1].long()], 0
			for rp, _, _ in parts:
				rps.append(rp[1:].long() + off)
				off += int(rp[-1])
			new.rowptr.append(torch.cat(rps).int())
			new.col.append(torch.cat([p[1] for p in parts]))
			new.val.append(torch.cat([p[2] for p in parts]))
		good = [slice(c, min(c + new.bs_cell, new.num_cell)) for c in range(0, new.num_cell, new.bs_cell)]
		bad = [slice(c, min(c + new.bs_cell, n)) for c in range(new.num_cell, n, new.bs_cell)]
		new.cell_slice_list = good + bad
		new.num_cell_batch, new.num_cell_batch_bad = len(good), len(bad)
		new.shape = list(self.shape[:2]) + [new.num_cell]
		return new
