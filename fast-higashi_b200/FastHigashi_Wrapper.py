"""`FastHigashi` - the user-facing API of the reference (fasthigashi/FastHigashi_Wrapper.py:115-878),
kept signature-for-signature for the decomposition hot path:

    FastHigashi(config_path, path2input_cache, path2result_dir, off_diag, filter, do_conv, do_rwr,
                do_col, no_col)                                            (:116-124)
    .prep_dataset(meta_only=False, batch_norm=True)                        (:460)
    .run_model(dim1=.6, rank=256, n_iter_parafac=1, n_iter_max=None, tol=2e-5, extra="",
               run_init=True)                                              (:657-663)
    .fetch_cell_embedding(final_dim=None, restore_order=False)             (:750)
    .load_model(...), .correct_batch_linear(...)                           (:704, :815)

`run_model` drives the B200 core (parafac2_intergrative.Fast_Higashi_core of this package) instead
of the reference's torch path; it fills the same attributes and writes the same pickles.

Scope: raw-file ingest, QC and normalisation (`pack_training_data_one_process`,
preprocessing.py, Fast_process.py) are NOT part of the hot path. `prep_dataset` therefore accepts the
tensors in one of three ways, in this order:
  1. `set_tensors(...)`: in-memory COO tensors per chromosome (what `pack_training_data_one_process`
     returns, FastHigashi_Wrapper.py:366) - the native entry;
  2. the reference's own input cache `cache_intra_{res}_offdiag_{off_diag}_.pkl` (:482-483) holding its
     `Sparse` objects, if a reference installation wrote one (needs the reference importable to unpickle);
  3. otherwise it raises: run the reference's ingest first (documented in INTEGRATION.md).
"""
import json
import math
import os
import pickle
import sys
import time

import numpy as np
import torch

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
		if qc is None and os.path.isfile(os.path.join(self.path2input_cache, "qc.npy")):
			qc = np.load(os.path.join(self.path2input_cache, "qc.npy"))
			readcount = np.load(os.path.join(self.path2input_cache, "read_count_all.npy"))
		if qc is None:
			raise RuntimeError("no QC information: pass qc/readcount to set_tensors() or provide the reference's "
			                   "qc.npy / read_count_all.npy in path2input_cache (get_qc is offline ingest, out of scope)")
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
		np.save(os.path.join(self.path2input_cache, "reorder.npy"), reorder)
		if "batch_id" in self.config:
			self.batch_id = np.asarray(label_info[self.config["batch_id"]])
		return label_info, reorder, readcount, qc

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
		path = os.path.join(self.path2input_cache, "cache_intra_%d_offdiag_%d_%s.pkl" % (res, self.off_diag, ""))
		if os.path.exists(path):
			out = []
			with open(path, "rb") as f:
				for _ in self.chrom_list:
					out.append(pickle.load(f))  # the reference's Sparse objects (reference must be importable)
			return out
		raise RuntimeError("no input tensors: call set_tensors() or let the reference's prep_dataset write %s first "
		                   "(contact-pair ingest and normalisation are outside the hot path, INTEGRATION.md)" % path)

	def prep_dataset(self, meta_only=False, batch_norm=True):
		"""FastHigashi_Wrapper.py:460-567 from the tensor stage on: batch sizes (:500-517), auto do_col
		(:545-551), one device-resident block-CSR `Chrom_Dataset` per (resolution, chromosome)."""
		self.label_info, reorder, readcount, qc = self.preprocess_meta()
		self.reorder = reorder
		self.coverage_feats = readcount[reorder].reshape((-1, 1))
		if meta_only:
			return
		good_qc_num = int(np.sum(qc > 0))
		print("total number of cells that pass qc check", good_qc_num, "bad", len(qc) - good_qc_num, "total:", len(qc))
		datasets = []
		for res in self.fh_resolutions:
			all_matrix = self._load_tensors(res, reorder)
			num_cell = int(all_matrix[-1].shape[-1])
			max_tensor_size = self.avail_mem / (4 * 12)
			recommend_bs_bin = min(max(int(15000000 / res), 128), 256)
			total_reads, total_possible = 0, 0
			for i, m in enumerate(all_matrix):
				size = int(m.shape[0])
				n_batch = max(math.ceil(size / recommend_bs_bin), 1)
				bs_bin_local = math.ceil(size / n_batch)
				bs_cell = int(max_tensor_size / (bs_bin_local * (bs_bin_local + 2 * self.off_diag)))
				ncell_eff = good_qc_num if self.filter else num_cell
				n_cb = int(math.ceil(ncell_eff / max(bs_cell, 1)))
				bs_cell = min(int(math.ceil(ncell_eff / n_cb)), ncell_eff)
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

	def run_model(self, dim1=.6, rank=256, n_iter_parafac=1, n_iter_max=None, tol=2e-5, extra="", run_init=True):
		"""FastHigashi_Wrapper.py:657-701."""
		self.rank = rank
		save_str = "dim1_%.1f_rank_%d_niterp_%d_%s" % (dim1, rank, n_iter_parafac, extra)
		self.save_str = save_str
		print(save_str)
		start = time.time()
		if self.model is None:
			self.model = Fast_Higashi_core(rank=rank, off_diag=self.off_diag, res_list=self.fh_resolutions).to(self.device)
		if n_iter_max is None:
			n_iter_max = int(self.good_qc_num / 15)
		result = self.model.fit_transform(self.all_matrix, size_ratio=dim1, n_iter_max=n_iter_max, n_iter_parafac=n_iter_parafac,
		                                  do_conv=self.do_conv, do_rwr=self.do_rwr, do_col=self.final_do_col, tol=tol,
		                                  gpu_id=self.gpu_id, run_init=run_init)
		print("takes: %.2f s" % (time.time() - start))
		_, factors_all, p_list = result
		A_list, B_list, D_list, meta_embedding = factors_all
		self.meta_embedding = meta_embedding.detach().cpu().numpy()
		self.A_list = [A.detach().cpu().numpy() for A in A_list]
		self.B_list = [B.detach().cpu().numpy() for B in B_list]
		self.D_list = [D.detach().cpu().numpy() for D in D_list]
		self.p_list = [[p.detach().cpu().numpy() for p in temp] for temp in p_list]
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

	def fetch_cell_embedding(self, final_dim=None, restore_order=False):
		"""FastHigashi_Wrapper.py:750-789 (host numpy/sklearn post-processing, as in the reference)."""
		print("fetching embedding")
		from sklearn.preprocessing import quantile_transform, normalize
		from sklearn.decomposition import TruncatedSVD
		final_dim = self.rank if final_dim is None else final_dim
		embedding_list = []
		for p in self.D_list:
			p = np.asarray(p)
			p = p / np.linalg.norm(p, axis=0, keepdims=True)
			embedding_list.append(self.meta_embedding @ p)
		embedding = np.concatenate(embedding_list, axis=1)
		self.label_info["coverage_fh"] = quantile_transform(self.coverage_feats, n_quantiles=100)
		embed = TruncatedSVD(n_components=final_dim).fit_transform(embedding)
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
		from sklearn.decomposition import TruncatedSVD
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
		reduce = TruncatedSVD(n_components=self.embedding_storage["embed_raw"].shape[-1]).fit_transform(embedding)
		self.embedding_storage["embed_correct_%s" % key] = reduce
		self.embedding_storage["embed_l2_norm_correct_%s" % key] = normalize(reduce)
		return self.embedding_storage
