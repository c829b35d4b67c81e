"""`FastHigashi` - the user-facing API of the reference (fasthigashi/FastHigashi_Wrapper.py:115-878),
kept signature-for-signature for the decomposition hot path:

    FastHigashi(config_path, path2input_cache, path2result_dir, off_diag, filter, do_conv, do_rwr,
                do_col, no_col)                                            (:116-124)
    .prep_dataset(meta_only=False, batch_norm=True)                        (:460)
    .run_model(dim1=.6, rank=256, n_iter_parafac=1, n_iter_max=None, tol=2e-5, extra="",
               run_init=True)                                              (:657-663)
    .fetch_cell_embedding(final_dim=None, restore_order=False)             (:750)
    .load_model(...), .correct_batch_linear(...), .only_partial_rwr()      (:704, :815, :569)
    .get_qc(), .pack_training_data_one_process(...)                        (:428, :221)

`run_model` drives the B200 core (parafac2_intergrative.Fast_Higashi_core of this package) instead
of the reference's torch path; it fills the same attributes and writes the same pickles.

Scope: the tensor stage of the reference's ingest (`get_qc`,
`pack_training_data_one_process`, the per-resolution cache) is provided by `ingest.py`; contact-pair
parsing (`Fast_process.py`) is not. `prep_dataset` takes the tensors from, in this order:
  1. `set_tensors(...)`: in-memory COO tensors per chromosome (what `pack_training_data_one_process`
     returns, FastHigashi_Wrapper.py:366);
  2. this package's cache `cache_intra_{res}_offdiag_{off_diag}_b200.pkl`, else the reference's own
     `cache_intra_{res}_offdiag_{off_diag}_.pkl` (:482-483; needs the reference importable to unpickle);
  3. `{temp_dir}/raw/{chrom}_sparse_adj.npy` (per-cell scipy CSR, what the reference's `extract_table`
     writes) through `ingest.preprocess_contact_map`;
  4. otherwise it raises.
"""
import json
import math
import os
import pickle
import time

import numpy as np
import torch

from . import ingest
from .parafac2_intergrative import Fast_Higashi_core
from .sparse_for_schic import Sparse, Chrom_Dataset


def get_config(config_path="./config.jSON"):
	with open(config_path, "r") as c:
		return json.load(c)


class FastHigashi:
	def __init__(self, config_path, path2input_cache, path2result_dir, off_diag, filter, do_conv, do_rwr, do_col, no_col):
		self.off_diag = off_diag
		self.filter = filter
		self.do_conv = do_conv
		self.do_rwr = do_rwr
		self.do_col = do_col
		self.no_col = no_col
		self.config_path = config_path
		self.config = get_config(config_path) if isinstance(config_path, (str, os.PathLike)) else dict(config_path)
		self.chrom_list = self.config["chrom_list"]
		self.temp_dir = self.config["temp_dir"]
		self.data_dir = self.config.get("data_dir", self.temp_dir)
		self.fh_resolutions = self.config["resolution_fh"]
		self.embedding_storage = None
		self.model = None
		self.path2input_cache = path2input_cache or self.temp_dir
		self.path2result_dir = path2result_dir or self.temp_dir
		for d in (self.path2input_cache, self.path2result_dir):
			if not os.path.exists(d):
				os.makedirs(d, exist_ok=True)
		if not torch.cuda.is_available():
			raise RuntimeError("fasthigashi_b200.FastHigashi needs a CUDA device; there is no CPU path")
		self.gpu_id = torch.cuda.current_device()
		self.device = "cuda:%d" % self.gpu_id
		self.avail_mem = torch.cuda.mem_get_info(self.gpu_id)[0]
		self._tensors = None
		self._meta = None
		self.group = None

	def distribute(self, group):
		"""Cell-sharded mode, one process per GPU (not in the reference, which is single-GPU): call on every
		rank of `group` (a torch.distributed process group) before `prep_dataset`. The ingest is partitioned by
		CHROMOSOME (rank k packs chromosomes k, k + N, ... for all cells - the pooled normalisation is per chromosome, so
		nothing is reduced), every chromosome's block-CSR is then scattered as cell slabs; `run_model` runs the sharded
		core and gathers the embedding rows, so `fetch_cell_embedding` works unchanged on every rank; rank 0 writes files."""
		self.group = group
		return self

	def _rank_world(self):
		if getattr(self, "group", None) is None:
			return 0, 1
		import torch.distributed as dist
		return dist.get_rank(self.group), dist.get_world_size(self.group)

	# ------------------------------------------------------------------------------------------
	def set_tensors(self, tensors, qc=None, readcount=None, label_info=None):
		"""tensors: {resolution: [(indices (3, nnz) [row, col, cell], values (nnz,), shape (n, n, cells)),
		...one per chromosome of config['chrom_list']]} - the output of the reference's
		`pack_training_data_one_process` (:221-366) for every chromosome. qc (cells,) > 0 marks good
		cells; readcount (cells,) is the log1p total read count (`get_qc`, :428-458)."""
		self._tensors = tensors
		self._meta = (qc, readcount, label_info)

	def preprocess_meta(self):
		"""FastHigashi_Wrapper.py:176-211: good-QC cells first; returns (label_info, reorder, readcount, qc)."""
		import pandas as pd
		qc, readcount, label_info = self._meta if self._meta is not None else (None, None, None)
		rank, world = self._rank_world()
		cached = qc is None and os.path.isfile(os.path.join(self.path2input_cache, "qc.npy"))
		if world > 1:  # one decision for all ranks: get_qc is a collective there, and rank 0 (re)writes the cache below
			import torch.distributed as dist
			flag = [cached]
			dist.broadcast_object_list(flag, src=dist.get_global_rank(self.group, 0) if hasattr(dist, "get_global_rank") else 0, group=self.group)
			cached = flag[0]
		if cached:
			qc = np.load(os.path.join(self.path2input_cache, "qc.npy"))
			readcount = np.load(os.path.join(self.path2input_cache, "read_count_all.npy"))
		if world > 1:
			dist.barrier(group=self.group)  # everybody has read the cache before rank 0 rewrites it
		if qc is None:
			qc, readcount = self.get_qc()
		qc = np.asarray(qc)
		readcount = np.asarray(readcount)
		good, bad = np.where(qc > 0)[0], np.where(qc <= 0)[0]
		reorder = np.concatenate([np.sort(good), np.sort(bad)], axis=0)
		if label_info is None:
			p = os.path.join(self.data_dir, "label_info.pickle")
			label_info = pickle.load(open(p, "rb")) if os.path.exists(p) else {}
		label_info = pd.DataFrame(label_info)
		if len(label_info) == 0:
			label_info = pd.DataFrame(np.ones(len(readcount)), columns=["placeholder"])
		label_info = label_info.iloc[reorder].reset_index()
		if "batch_id" in self.config:
			self.batch_id = np.asarray(label_info[self.config["batch_id"]])
		if rank == 0:
			np.save(os.path.join(self.path2input_cache, "reorder.npy"), reorder)
			np.save(os.path.join(self.path2input_cache, "qc.npy"), qc)  # :208-209
			np.save(os.path.join(self.path2input_cache, "read_count_all.npy"), readcount)
		return label_info, reorder, readcount, qc

	def get_qc(self):
		"""FastHigashi_Wrapper.py:428-458 over `{temp_dir}/raw/{chrom}_sparse_adj.npy`."""
		raw_dir = os.path.join(self.temp_dir, "raw")
		if not os.path.isdir(raw_dir):
			raise RuntimeError("no QC information: pass qc/readcount to set_tensors(), provide qc.npy / read_count_all.npy in "
			                   "path2input_cache, or the per-cell matrices under %s" % raw_dir)
		rank, world = self._rank_world()
		if world == 1:
			return ingest.get_qc(raw_dir, self.chrom_list, self.config["resolution"])
		# chromosome-partitioned: per-cell pass counts and read sums of disjoint chromosome subsets add up exactly
		import torch.distributed as dist
		mine = self.chrom_list[rank::world]
		passed, reads, dtype = ingest.qc_partial(raw_dir, mine, self.config["resolution"]) if mine else (0, 0, np.float64)
		ncell = [int(np.size(passed)) if mine else 0]
		sizes = [None] * world
		dist.all_gather_object(sizes, (ncell[0], np.dtype(dtype).str), group=self.group)
		n = max(x[0] for x in sizes)
		dtype = np.dtype(next(x[1] for x in sizes if x[0] > 0))
		buf = torch.zeros(2, n, dtype=torch.float64)
		if mine:
			buf[0], buf[1] = torch.from_numpy(np.asarray(passed, dtype=np.float64)), torch.from_numpy(np.asarray(reads, dtype=np.float64))
		buf = buf.to(self.device)
		dist.all_reduce(buf, group=self.group)
		buf = buf.cpu().numpy()
		return ingest.qc_combine(buf[0], buf[1], len(self.chrom_list), dtype)

	def pack_training_data_one_process(self, raw_dir, chrom, reorder, off_diag=None, fac_size=None, merge_fac_row=1,
	                                   merge_fac_col=1, is_sym=True, filename_pattern="%s_sparse_adj.npy", force_shift=None,
	                                   batch_norm=True, bar=None):
		"""Reference signature (FastHigashi_Wrapper.py:221-231); only the options the wrapper itself
		uses are supported (fac_size 1, is_sym, no forced shift)."""
		if (fac_size not in (None, 1)) or not is_sym or force_shift:
			raise NotImplementedError("fac_size != 1, is_sym=False and force_shift are never used by prep_dataset (:484-494)")
		return ingest.pack_training_data_one_process(
			raw_dir, chrom, reorder, self.off_diag if off_diag is None else off_diag, merge_fac_row, merge_fac_col,
			getattr(self, "batch_id", None) if "batch_id" in self.config else None, batch_norm,
			ingest.load_blacklist(self.temp_dir), filename_pattern)

	def preprocess_contact_map(self, config, reorder, path2input_cache, batch_norm, key_fn=lambda c: c, off_diag=None,
	                           merge_fac_row=1, merge_fac_col=1, fac_size=1, is_sym=True, force_shift=False,
	                           filename_pattern="%s_sparse_adj.npy", **kwargs):
		"""Reference signature (FastHigashi_Wrapper.py:368-412): every chromosome of `config['chrom_list']` packed at the
		coarsening `merge_fac_row` (= merge_fac_col), cached in `path2input_cache`, returned as a list of `Sparse` tensors
		sorted on dim 0. The cache is this package's own format (plain arrays), not class pickles."""
		if (fac_size not in (None, 1)) or not is_sym or force_shift or merge_fac_row != merge_fac_col:
			raise NotImplementedError("fac_size != 1, is_sym=False, force_shift and unequal merge factors are never used by prep_dataset (:484-494)")
		res = int(config["resolution"]) * int(merge_fac_row)
		packed = ingest.preprocess_contact_map(config, reorder, path2input_cache, self.off_diag if off_diag is None else off_diag, res,
		                                       getattr(self, "batch_id", None) if "batch_id" in self.config else None, batch_norm)
		out = []
		for idx, val, shape in packed:
			m = Sparse(torch.as_tensor(idx), torch.as_tensor(val), shape, copy=False)
			m.sort_indices()
			out.append(m)
		return out

	def _load_tensors(self, res, reorder):
		if self._tensors is not None:
			out = []
			inv = np.empty(len(reorder), dtype=np.int64)
			inv[reorder] = np.arange(len(reorder))
			for idx, val, shape in self._tensors[res]:
				idx = torch.as_tensor(np.asarray(idx)).long().clone()
				idx[2] = torch.as_tensor(inv)[idx[2]]  # cell ids follow `reorder` (:233)
				out.append(Sparse(idx, torch.as_tensor(np.asarray(val)).float(), shape, copy=False))
			return out
		ours = os.path.join(self.path2input_cache, "cache_intra_%d_offdiag_%d_b200.pkl" % (res, self.off_diag))
		path = os.path.join(self.path2input_cache, "cache_intra_%d_offdiag_%d_%s.pkl" % (res, self.off_diag, ""))
		if not os.path.exists(ours) and os.path.exists(path):
			out = []
			with open(path, "rb") as f:
				for _ in self.chrom_list:
					out.append(pickle.load(f))  # the reference's Sparse objects (reference must be importable)
			return out
		if os.path.exists(ours) or os.path.isdir(os.path.join(self.temp_dir, "raw")):
			packed = ingest.preprocess_contact_map(self.config, reorder, ours, self.off_diag, res,
			                                       getattr(self, "batch_id", None) if "batch_id" in self.config else None,
			                                       self._batch_norm)
			return [Sparse(torch.as_tensor(idx), torch.as_tensor(val), shape, copy=False) for idx, val, shape in packed]  # int32 ids: no copy
		raise RuntimeError("no input tensors: call set_tensors(), or provide %s/raw/{chrom}_sparse_adj.npy, or a cache file %s"
		                   % (self.temp_dir, ours))

	def prep_dataset(self, meta_only=False, batch_norm=True):
		"""FastHigashi_Wrapper.py:460-567 from the tensor stage on: batch sizes (:500-517), auto do_col
		(:545-551), one device-resident block-CSR `Chrom_Dataset` per (resolution, chromosome)."""
		self._batch_norm = batch_norm
		self.label_info, reorder, readcount, qc = self.preprocess_meta()
		self.reorder = reorder
		self.coverage_feats = readcount[reorder].reshape((-1, 1))
		if meta_only:
			return
		good_qc_num = int(np.sum(qc > 0))
		print("total number of cells that pass qc check", good_qc_num, "bad", len(qc) - good_qc_num, "total:", len(qc))
		if self._rank_world()[1] > 1:
			return self._prep_dataset_distributed(reorder, good_qc_num, len(qc))
		datasets = []
		for res in self.fh_resolutions:
			all_matrix = self._load_tensors(res, reorder)
			num_cell = int(all_matrix[-1].shape[-1])
			total_reads, total_possible = 0, 0
			for i, m in enumerate(all_matrix):
				bs_bin_local, bs_cell = self._batch_sizes(int(m.shape[0]), res, good_qc_num, num_cell)
				total_reads += len(m.values)
				total_possible += float(np.prod(np.asarray(m.shape, dtype=np.float64)))
				datasets.append(Chrom_Dataset(tensor=m, bs_bin=bs_bin_local, bs_cell=bs_cell,
				                              good_qc_num=good_qc_num if self.filter else -1, kind="hic", upper_sim=False,
				                              compact=True, flank=self.off_diag, chrom=self.chrom_list[i], resolution=res,
				                              device=self.device))
			sparsity = total_reads / total_possible
			print("sparsity", sparsity)
			do_col = sparsity * (500000 / res) ** 2 <= 0.03 or self.do_col
			if self.no_col:
				do_col = False
			print("do_conv", self.do_conv, "do_rwr", self.do_rwr, "do_col", do_col)
			self.final_do_col = do_col
			if self.no_col and self.do_col:
				print("choose one between do col or no col!")
				raise EOFError
		self.good_qc_num = good_qc_num
		self.all_matrix = datasets

	def _batch_sizes(self, size, res, good_qc_num, num_cell):
		"""bs_bin, bs_cell of one chromosome (FastHigashi_Wrapper.py:500-517)."""
		max_tensor_size = self.avail_mem / (4 * 12)
		recommend_bs_bin = min(max(int(15000000 / res), 128), 256)
		n_batch = max(math.ceil(size / recommend_bs_bin), 1)
		bs_bin_local = math.ceil(size / n_batch)
		bs_cell = int(max_tensor_size / (bs_bin_local * (bs_bin_local + 2 * self.off_diag)))
		# one fh_rwr call takes at most 65,535 cells and the auto-stop RWR of init_params / only_partial_rwr is one call per
		# cell batch (partial_rwr.py:119-123 decides per batch): larger batches would be rejected by the library, so the rule
		# is capped here (it only binds for > 65,535 cells on one GPU at low resolution, with 180 GB of HBM)
		bs_cell = min(bs_cell, 65535)
		ncell_eff = good_qc_num if self.filter else num_cell
		n_cb = int(math.ceil(ncell_eff / max(bs_cell, 1)))
		return bs_bin_local, min(int(math.ceil(ncell_eff / n_cb)), ncell_eff)

	def _load_tensor_one(self, res, reorder, i):
		"""The COO tensor of chromosome i at `res` (distributed mode: only the owner rank loads it)."""
		if self._tensors is not None:
			idx, val, shape = self._tensors[res][i]
			inv = np.empty(len(reorder), dtype=np.int64)
			inv[reorder] = np.arange(len(reorder))
			idx = torch.as_tensor(np.asarray(idx)).long().clone()
			idx[2] = torch.as_tensor(inv)[idx[2]]
			return Sparse(idx, torch.as_tensor(np.asarray(val)).float(), shape, copy=False)
		raw_dir = os.path.join(self.temp_dir, "raw")
		if not os.path.isdir(raw_dir):
			raise RuntimeError("distributed prep_dataset needs set_tensors() or the per-cell matrices under %s" % raw_dir)
		fac = int(res / self.config["resolution"])
		idx, val, shape = ingest.pack_training_data_one_process(
			raw_dir, self.chrom_list[i], reorder, self.off_diag, fac, fac,
			getattr(self, "batch_id", None) if "batch_id" in self.config else None, self._batch_norm, ingest.load_blacklist(self.temp_dir))
		return Sparse(torch.as_tensor(idx), torch.as_tensor(val), shape, copy=False)

	def _prep_dataset_distributed(self, reorder, good_qc_num, num_cell):
		"""prep_dataset when `distribute(group)` was called: chromosome i is packed and staged (on the host) by rank
		i mod N alone, then scattered as cell slabs; every rank ends with `shard_datasets(all cells, N, rank)`."""
		from .sharding import scatter_dataset
		rank, world = self._rank_world()
		datasets = []
		for res in self.fh_resolutions:
			total_reads, total_possible = 0, 0.0
			for i, chrom in enumerate(self.chrom_list):
				owner = i % world
				full = None
				if rank == owner:
					m = self._load_tensor_one(res, reorder, i)
					bs_bin_local, bs_cell = self._batch_sizes(int(m.shape[0]), res, good_qc_num, num_cell)
					full = Chrom_Dataset(tensor=m, bs_bin=bs_bin_local, bs_cell=bs_cell, good_qc_num=good_qc_num if self.filter else -1,
					                     kind="hic", upper_sim=False, compact=True, flank=self.off_diag, chrom=chrom, resolution=res,
					                     device="cpu")
					del m
				slab, meta = scatter_dataset(full, owner, self.group, self.device)
				del full
				total_reads += meta["nnz"]
				total_possible += float(np.prod(np.asarray(meta["shape"], dtype=np.float64)))
				datasets.append(slab)
			sparsity = total_reads / total_possible
			do_col = sparsity * (500000 / res) ** 2 <= 0.03 or self.do_col
			if self.no_col:
				do_col = False
			if rank == 0:
				print("sparsity", sparsity)
				print("do_conv", self.do_conv, "do_rwr", self.do_rwr, "do_col", do_col)
			self.final_do_col = do_col
			if self.no_col and self.do_col:
				print("choose one between do col or no col!")
				raise EOFError
		self.good_qc_num = good_qc_num
		self.all_matrix = datasets

	def only_partial_rwr(self, out_format=None):
		"""FastHigashi_Wrapper.py:569-655: impute every cell (good and bad QC) with conv + auto-stopped RWR
		(`force_rwr_epochs=-1`, `do_col=False`, one stop decision per (bin-block, cell batch) as in the
		reference), paste the block windows into the full (n, n) map, symmetrise (`m + m^T` with the
		diagonal halved) and write one fp32 dataset per cell named by its ORIGINAL cell id.
		Output: `{path2result_dir}/impute_prwr.hdf5` with the reference's layout (group per chromosome:
		"shape" + one dataset per cell) when h5py is importable or out_format="hdf5"; otherwise one
		`impute_prwr_{chrom}.npz` per chromosome with the same keys (h5py is not part of this image).
		Everything up to the final device->host copy runs on the GPU: fh_rwr_batched per block, the paste and the
		symmetrisation in fp32 (x + y and 2x/2 are exact in fp32, so the result equals the reference's
		float64-then-cast arithmetic bit for bit)."""
		from .partial_rwr import rwr_block_csr, pad4
		my_rank, world = self._rank_world()
		suffix = "" if world == 1 else "_rank%d" % my_rank  # distributed: every rank writes the maps of ITS cells
		if out_format is None:
			try:
				import h5py  # noqa: F401
				out_format = "hdf5"
			except ImportError:
				out_format = "npz"
		h5 = None
		if out_format == "hdf5":
			import h5py
			h5 = h5py.File(os.path.join(self.path2result_dir, "impute_prwr%s.hdf5" % suffix), "w")
		elif out_format != "npz":
			raise ValueError("out_format must be 'hdf5' or 'npz'")
		written = []
		for ds in self.all_matrix:
			n = ds.num_bin
			maps = {"shape": np.asarray([n, n])}
			# position of the dataset's cells in the unsharded (good first, then bad) order that `reorder` indexes
			if world == 1:
				cell_pos = np.arange(ds.total_cell_num)
			else:
				from .sharding import cell_slab
				glo, ghi = cell_slab(self.good_qc_num if self.filter else len(self.reorder), world, my_rank)
				nbad_all = len(self.reorder) - self.good_qc_num if self.filter else 0
				blo, bhi = cell_slab(nbad_all, world, my_rank)
				cell_pos = np.concatenate([np.arange(glo, ghi), self.good_qc_num + np.arange(blo, bhi)]) if self.filter else np.arange(glo, ghi)
				assert len(cell_pos) == ds.total_cell_num
			for sl in ds.cell_slice_list:
				c0, nc = sl.start, sl.stop - sl.start
				if nc <= 0:
					continue
				panels = []
				for b, g in enumerate(ds.geoms):
					ldw = pad4(g.w)
					X = torch.zeros(nc, g.nb * ldw, dtype=torch.float32, device=self.device)
					rwr_block_csr(ds, b, c0, nc, X, g.nb * ldw, -1, self.do_conv, self.do_rwr, False)
					panels.append(X.view(nc, g.nb, ldw))
				step = max(1, min(nc, (1 << 30) // (4 * n * n)))
				for s0 in range(0, nc, step):
					s1 = min(s0 + step, nc)
					full = torch.zeros(s1 - s0, n, n, dtype=torch.float32, device=self.device)
					for X, g in zip(panels, ds.geoms):
						full[:, g.row0:g.row0 + g.nb, g.col0:g.col0 + g.w] = X[s0:s1, :, :g.w]
					full = full + full.transpose(1, 2)
					d = torch.diagonal(full, dim1=1, dim2=2)
					d.sub_(d / 2)
					host = full.cpu().numpy()
					for i in range(s1 - s0):
						maps[str(self.reorder[cell_pos[c0 + s0 + i]])] = host[i]
				del panels
			if h5 is not None:
				group = h5.create_group(ds.chrom)
				for k, v in maps.items():
					group.create_dataset(k, data=v)
			else:
				path = os.path.join(self.path2result_dir, "impute_prwr_%s%s.npz" % (ds.chrom, suffix))
				np.savez(path, **maps)
				written.append(path)
		if h5 is not None:
			h5.close()
			return os.path.join(self.path2result_dir, "impute_prwr%s.hdf5" % suffix)
		return written

	def run_model(self, dim1=.6, rank=256, n_iter_parafac=1, n_iter_max=None, tol=2e-5, extra="", run_init=True, init_svd="auto"):
		"""FastHigashi_Wrapper.py:657-701. `init_svd` (not in the reference): "host" = the reference's sklearn SVD of the pooled
		features (reproduces its start from a shared seed), "device" = cell-sharded randomized SVD, "auto" = host for a single
		process up to 20,000 cells, device otherwise (Fast_Higashi_core)."""
		self.rank = rank
		save_str = "dim1_%.1f_rank_%d_niterp_%d_%s" % (dim1, rank, n_iter_parafac, extra)
		self.save_str = save_str
		print(save_str)
		start = time.time()
		my_rank, world = self._rank_world()
		if self.model is None:
			self.model = Fast_Higashi_core(rank=rank, off_diag=self.off_diag, res_list=self.fh_resolutions,
			                               group=self.group if world > 1 else None, init_svd=init_svd).to(self.device)
		if n_iter_max is None:
			n_iter_max = int(self.good_qc_num / 15)
		result = self.model.fit_transform(self.all_matrix, size_ratio=dim1, n_iter_max=n_iter_max, n_iter_parafac=n_iter_parafac,
		                                  do_conv=self.do_conv, do_rwr=self.do_rwr, do_col=self.final_do_col, tol=tol,
		                                  gpu_id=self.gpu_id, run_init=run_init, verbose=my_rank == 0)
		print("takes: %.2f s" % (time.time() - start))
		_, factors_all, p_list = result
		A_list, B_list, D_list, meta_embedding = factors_all
		if world > 1:  # rows of this rank's good then bad cells -> all cells in the unsharded order, on every rank
			from .sharding import gather_cell_rows
			meta_embedding = gather_cell_rows(meta_embedding, self.all_matrix[0].num_cell, self.group)
		self.meta_embedding = meta_embedding.detach().cpu().numpy()
		self.A_list = [A.detach().cpu().numpy() for A in A_list]
		self.B_list = [B.detach().cpu().numpy() for B in B_list]
		self.D_list = [D.detach().cpu().numpy() for D in D_list]
		self.p_list = [[p.detach().cpu().numpy() for p in temp] for temp in p_list]
		if my_rank != 0:
			return
		pickle.dump([self.A_list, self.B_list, self.D_list, self.meta_embedding, self.p_list],
		            open(os.path.join(self.path2result_dir, "results_all%s.pkl" % save_str), "wb"), protocol=4)
		pickle.dump([self.meta_embedding, self.D_list],
		            open(os.path.join(self.path2result_dir, "results%s.pkl" % save_str), "wb"), protocol=4)

	def load_model(self, dim1=.6, rank=256, n_iter_parafac=1, extra=""):
		save_str = "dim1_%.1f_rank_%d_niterp_%d_%s" % (dim1, rank, n_iter_parafac, extra)
		data = pickle.load(open(os.path.join(self.path2result_dir, "results_all%s.pkl" % save_str), "rb"))
		print("model loaded")
		self.A_list, self.B_list, self.D_list, self.meta_embedding, self.p_list = data
		self.rank = rank

	def restore_order_fun(self, x):
		new_x = np.zeros_like(x)
		new_x[self.reorder] = x
		return new_x

	def _reduce(self, embedding, dim, svd):
		"""TruncatedSVD(n_components=dim).fit_transform(embedding): sklearn on the host (reference behaviour, :767,:872) or
		the same randomized algorithm on the device (dist_svd.py; matters from ~100k cells on)."""
		if svd == "auto":  # the reference's host route (same numpy RNG stream) up to 20,000 cells, the device above - as init_svd
			svd = "host" if (embedding.shape[0] <= 20000 or not torch.cuda.is_available()) else "device"
		if svd == "host":
			from sklearn.decomposition import TruncatedSVD
			return TruncatedSVD(n_components=dim).fit_transform(embedding)
		if svd != "device":
			raise ValueError("svd must be 'auto', 'host' or 'device'")
		from .dist_svd import sharded_truncated_svd
		dev = getattr(self, "device", "cpu")
		emb, _, _ = sharded_truncated_svd(torch.as_tensor(embedding, dtype=torch.float64).to(dev), dim, n_iter=5,
		                                  seed=int(np.random.randint(0, 2 ** 31 - 1)))
		return emb.cpu().numpy()

	def fetch_cell_embedding(self, final_dim=None, restore_order=False, svd="auto"):
		"""FastHigashi_Wrapper.py:750-789. `svd`: "host" = numpy / sklearn post-processing exactly as the reference, "device" =
		the two truncated SVDs on the GPU (dist_svd.py), "auto" (default) = host up to 20,000 cells, device above."""
		print("fetching embedding")
		from sklearn.preprocessing import quantile_transform, normalize
		final_dim = self.rank if final_dim is None else final_dim
		self._embed_svd = svd
		embedding_list = []
		for p in self.D_list:
			p = np.asarray(p)
			p = p / np.linalg.norm(p, axis=0, keepdims=True)
			embedding_list.append(self.meta_embedding @ p)
		embedding = np.concatenate(embedding_list, axis=1)
		self.label_info["coverage_fh"] = quantile_transform(self.coverage_feats, n_quantiles=100)
		embed = self._reduce(embedding, final_dim, svd)
		if restore_order:
			embedding = self.restore_order_fun(embedding)
			embed = self.restore_order_fun(embed)
		store = {"embed_all": embedding, "embed_raw": embed, "embed_l2_norm": normalize(embed), "restore_order": restore_order}
		self.embedding_storage = store
		self.correct_batch_linear("coverage_fh", False)
		return store

	def correct_batch_linear(self, var_to_regress_name, add_intercept_back=False):
		"""FastHigashi_Wrapper.py:815-878."""
		from sklearn.linear_model import LinearRegression
		from sklearn.preprocessing import normalize
		if self.embedding_storage is None:
			print("Run fetch_cell_embedding() first!")
			return None
		names = [var_to_regress_name] if isinstance(var_to_regress_name, str) else list(var_to_regress_name)
		cols = []
		for name in names:
			if name not in self.label_info:
				print("var_to_regress %s not in label_info.pickle!" % name)
				return None
			v = np.array(self.label_info[name])
			if self.embedding_storage["restore_order"]:
				v = self.restore_order_fun(v)
			if v.dtype not in [np.dtype("float32"), np.dtype("float16"), np.dtype("float64")]:
				uniq = np.unique(v)
				v = np.stack([(v == u).astype(np.float64) for u in uniq], 1)
			cols.append(v.reshape(len(v), -1))
		var = np.concatenate(cols, axis=-1)
		key = "_".join(names)
		model = LinearRegression()
		embedding = self.embedding_storage["embed_all"]
		embedding = embedding - model.fit(var, embedding).predict(var)
		if add_intercept_back:
			embedding = embedding + model.intercept_[None]
		reduce = self._reduce(embedding, self.embedding_storage["embed_raw"].shape[-1], getattr(self, "_embed_svd", "host"))
		self.embedding_storage["embed_correct_%s" % key] = reduce
		self.embedding_storage["embed_l2_norm_correct_%s" % key] = normalize(reduce)
		return self.embedding_storage
