"""Integrative PARAFAC2 driver on the B200 (mirror of fasthigashi/parafac2_intergrative.py (sic)).

Same class name, methods and return values as the reference's `Fast_Higashi_core`
(parafac2_intergrative.py:45-850); the sweep is re-derived for a device-resident design:

  * the imputed tensor X of every (chromosome, bin-block) lives in HBM as a (cells, nb*ldw) fp32
    matrix (cell-major panels), produced by ONE RWR pass per sweep (`cache="sweep"`, default) or per
    run (`cache="run"`); the reference re-imputes 2-3x per sweep (:357,451,508);
  * with that layout the three cell-mode contractions are plain large GEMMs against X
        P1  T1  = X^T (V D)                       (:336,374-383)   temp_i = (T1_i diag(A_i)) B^T
        P3  M  += X W,  W_i = U_i B diag(A_i) D^T (:422-430)       M = SVD_term^T
        P5  Z   = X^T V, Y_i = U_i^T Z_i          (:522-529)
    and the per-bin pieces are small batched GEMMs with the A-row scaling folded into the load;
  * cells shard over ranks (one process per GPU): T1, Y, the R x R Gram of M and three scalars are
    all-reduced (NCCL), everything else is local.

PyTorch is used for memory, streams and torch.distributed only; all arithmetic on the timed path
goes through libfh_b200.so (fast-higashi_b200/_lib.py).
"""
import math
import os
import sys
import time
import numpy as np
import torch

from . import _lib
from .partial_rwr import rwr_block_csr, pad4, cells_per_chunk
from .project2orthogonal import polar_tall
from .parafac_integrative import cp_als_, core_sqnorm_accum
from .sparse_for_schic import Chrom_Dataset
from .sharding import polar_bin_range
from .dist_svd import sharded_truncated_svd, sharded_svd_gram


def _as_block_csr(ds, device):
	if isinstance(ds, Chrom_Dataset):
		return ds.to(device)
	if hasattr(ds, "tensor_list"):  # a reference Chrom_Dataset (duck-typed)
		return Chrom_Dataset.from_reference(ds, device=device)
	raise TypeError("schic entries must be Chrom_Dataset objects (ours or the reference's)")


MAX_POLAR_SIDE = 160  # fh_polar.cu kMaxGram: a Gram matrix of that side (fp64) fills the 227 KB of shared memory
MAX_CP_RANK = 169     # fh_cp.cu: (r (r | 1) + r) * 8 bytes of shared memory for the r x r SPD inverse


class Fast_Higashi_core:
	HOST_INIT_MAX_CELLS = 20000  # init_svd="auto": above this (or when cell-sharded) the init SVDs stay on the device

	def __init__(self, rank, off_diag, res_list, cache="sweep", use_tc=None, group=None, init_svd="auto"):
		self.rank = rank
		self.off_diag = off_diag
		self.res_list = res_list
		self.device = torch.device("cpu")
		self.cache = cache            # "sweep": one RWR pass per ALS sweep; "run": one per run
		self.use_tc = use_tc          # None -> decided in .to()
		self.group = group            # torch.distributed process group when cell-sharded
		if init_svd not in ("auto", "host", "device"):
			raise ValueError("init_svd must be 'auto', 'host' (the reference's sklearn SVD, features gathered to rank 0) or 'device'")
		# "host": the reference's own route (sklearn TruncatedSVD with the numpy global RNG: a shared seed reproduces the
		# reference's start). "device": cell-sharded randomized SVD (dist_svd.py), nothing gathered - measured 4.0 s for the
		# whole init of 100k cells on 8 GPUs, where the host route would gather ~50 GB of fp64 features per chromosome to
		# rank 0. "auto" (default): host for a single process with <= HOST_INIT_MAX_CELLS cells, device otherwise.
		self.init_svd = init_svd
		self.verbose = True
		self.n_rwr_passes = 0
		self._X = {}
		self._eig = {}
		self._scratch = {}
		self._Z_valid = set()

	def to(self, device):
		self.device = torch.device(device)
		_lib.require_cuda(self.device)
		_lib.lib()
		if self.use_tc is None:
			self.use_tc = True  # tcgen05 3xTF32 for the large contractions (csrc/fh_gemm_tc.cu)
		return self

	# ------------------------------------------------------------------------------------------
	def _dist(self):
		if self.group is None:
			return None
		import torch.distributed as dist
		return dist

	def _allreduce(self, t, op=None):
		d = self._dist()
		if d is not None:
			d.all_reduce(t, op=op or d.ReduceOp.SUM, group=self.group)
		return t

	def _allreduce_async(self, t):
		"""Start an all-reduce (SUM) of t on the communicator's own stream; returns a handle whose .wait() makes the current
		stream wait for it (None when not sharded). Lets the next block's RWR / GEMMs run under the transfer."""
		d = self._dist()
		if d is None:
			return None
		return d.all_reduce(t, op=d.ReduceOp.SUM, group=self.group, async_op=True)

	def _log(self, *a):
		if self.verbose and (self.group is None or self._dist().get_rank(self.group) == 0):
			print(*a)
			sys.stdout.flush()

	def _gemm_dtype(self):
		return _lib.GEMM_TF32X3 if self.use_tc else _lib.GEMM_F32

	# ------------------------------------------------------------------------------------------
	# sizes: parafac2_intergrative.py:558-590
	def _setup(self, schic, size_ratio, size_list):
		rank = self.rank
		self.schic = [_as_block_csr(ds, self.device) for ds in schic]
		if size_list is None:
			size_list = [min(int(ds.num_bin * size_ratio * ds.resolution / 1000000), rank) for ds in self.schic]
			chrom2size = {}
			for ds, size in zip(self.schic, size_list):
				chrom2size[ds.chrom] = min(chrom2size.get(ds.chrom, size), size)
		else:
			chrom2size = {}
			for ds, size in zip(self.schic, size_list):
				if ds.chrom in chrom2size and chrom2size[ds.chrom] != size:
					print("size of the same chromosome must be same!", size, chrom2size[ds.chrom], ds.chrom)
					raise EOFError
				chrom2size[ds.chrom] = size
		self.chrom2size = chrom2size
		self._check_limits()
		self.chrom2num_bin = {}
		self.chrom2id = {c: [] for c in chrom2size}
		for ci, ds in enumerate(self.schic):
			self.chrom2id[ds.chrom].append(ci)
			start = self.chrom2num_bin.get(ds.chrom, 0)
			ds.global_slice_bin = slice(start, start + ds.num_bin)
			self.chrom2num_bin[ds.chrom] = start + ds.num_bin
		for ds_in, ds in zip(schic, self.schic):
			try:
				ds_in.global_slice_bin = ds.global_slice_bin  # the reference writes this too (:586-589)
			except Exception:
				pass
		self.num_cell = self.schic[0].num_cell
		self.total_cell_num = self.schic[0].total_cell_num

	def _check_limits(self):
		"""The library's size limits (INTEGRATION.md "Limits"), checked BEFORE init_params spends its RWR passes: the per-bin
		polar step keeps a Gram matrix of side min(window, r) in shared memory (fh_polar.cu: <= 160) and the inner CP-ALS an
		r x r fp64 SPD inverse (fh_cp.cu: r <= 169). The reference has no such limit (rank 256 with dim1 > 0.64 on human chr1
		at 500 kb gives r > 160): fail here with the way out instead of at the first sweep."""
		for ds in self.schic:
			r = self.chrom2size[ds.chrom]
			side = max(min(pad4(g.w), r) for g in ds.geoms)
			if side > MAX_POLAR_SIDE or r > MAX_CP_RANK:
				raise ValueError("%s at %d bp: per-chromosome rank r = %d (Gram side %d of the per-bin polar step) exceeds this library's limits "
				                 "(Gram side <= %d, r <= %d); lower size_ratio / dim1 (r = int(bins * dim1 * resolution / 1e6)) or pass size_list"
				                 % (ds.chrom, ds.resolution, r, side, MAX_POLAR_SIDE, MAX_CP_RANK))

	# ------------------------------------------------------------------------------------------
	# I1: init_params, parafac2_intergrative.py:61-301
	@torch.no_grad()
	def init_params(self, schic, do_conv, do_rwr, do_col):
		dev, R = self.device, self.rank
		t0 = time.perf_counter()
		sizes = [self.chrom2size[ds.chrom] for ds in self.schic]
		# same CPU RNG call order as the reference (:71-80) so a shared seed gives the same start
		A_list = [torch.randn([ds.num_bin, r], dtype=torch.float32) * 1e-2 + 1 for ds, r in zip(self.schic, sizes)]
		B_dict = {c: torch.eye(r, dtype=torch.float32).add_(torch.randn(r, dtype=torch.float32), alpha=1e-2)
		          for c, r in self.chrom2size.items()}
		uniq = list(self.chrom2size.values()) * len(self.res_list)
		cum = np.concatenate([[0], np.cumsum(uniq)])
		dist = self._dist()
		rank0 = dist is None or dist.get_rank(self.group) == 0
		if self.init_svd == "auto":
			self.init_svd = "host" if (dist is None and self.schic[0].num_cell <= self.HOST_INIT_MAX_CELLS) else "device"
		C = None
		cstart = 0
		self.bin_cov_list, self.bad_bin_cov_list, n_i_all = [], [], []
		# FH_INIT_TIMING=1: where init_params spends its time (synchronises after every part; diagnostics only)
		import os as _os, time as _time
		_timing = _os.environ.get("FH_INIT_TIMING") == "1"
		_acc = {"features (RWR auto-stop + coverage + pooling)": 0.0, "per-chromosome SVD": 0.0, "joint SVD": 0.0}

		def _lap(key, t0):
			if _timing:
				torch.cuda.synchronize()
				_acc[key] += _time.perf_counter() - t0
			return _time.perf_counter()
		for ci, ds in enumerate(self.schic):
			nbad = ds.total_cell_num - ds.num_cell
			# one coverage table for good then bad cells (rows follow the dataset's cell order)
			cov = torch.full((ds.total_cell_num, ds.num_bin), 1e-4, dtype=torch.float32, device=dev)
			n1m = int(math.ceil(ds.num_bin * ds.resolution / 1000000))
			size1 = min(int(math.ceil(ds.num_bin / ds.num_bin_batch * ds.resolution / 1000000)) + 2 * self.off_diag + 1, n1m)
			feats_dim = int(math.ceil(n1m * size1))
			ll = int(math.ceil(1000000 / ds.resolution))
			feats = torch.zeros(ds.num_cell, feats_dim, dtype=torch.float32, device=dev)
			n_i_list = []

			def feature_pass(with_col, fstart):
				for b, g in enumerate(ds.geoms):
					ldw = pad4(g.w)
					width = 0
					for sl in ds.cell_slice_list[:ds.num_cell_batch]:
						nc = sl.stop - sl.start
						x = torch.empty(nc, g.nb * ldw, dtype=torch.float32, device=dev)
						n_i = rwr_block_csr(ds, b, sl.start, nc, x, g.nb * ldw, -1, do_conv, do_rwr, with_col,
						                    bin_cov=cov if with_col else None, use_tc=self.use_tc)
						if not with_col:
							n_i_list.append(n_i)
							_lib.check(_lib.lib().fh_colsum_accum(x.data_ptr(), nc, g.nb, g.w, ldw, g.nb * ldw,
							                                      cov.data_ptr() + 4 * (sl.start * cov.stride(0) + g.col0),
							                                      cov.stride(0), _lib.stream_ptr()))
						if with_col == do_col:
							orow, ocol = g.nb // ll, g.w // ll
							width = orow * ocol
							if width:
								pooled = torch.empty(nc, width, dtype=torch.float32, device=dev)
								_lib.check(_lib.lib().fh_avgpool(x.data_ptr(), nc, g.nb, g.w, ldw, g.nb * ldw, ll,
								                                 pooled.data_ptr(), width, _lib.stream_ptr()))
								feats[sl, fstart:fstart + width] = pooled
						del x
					fstart += width
					if not with_col:
						for sl in ds.cell_slice_list[ds.num_cell_batch:]:
							nc = sl.stop - sl.start
							x = torch.empty(nc, g.nb * ldw, dtype=torch.float32, device=dev)
							rwr_block_csr(ds, b, sl.start, nc, x, g.nb * ldw, -1, do_conv, do_rwr, False, use_tc=self.use_tc)
							_lib.check(_lib.lib().fh_colsum_accum(x.data_ptr(), nc, g.nb, g.w, ldw, g.nb * ldw,
							                                      cov.data_ptr() + 4 * (sl.start * cov.stride(0) + g.col0),
							                                      cov.stride(0), _lib.stream_ptr()))
							del x
				return fstart

			if _timing:
				torch.cuda.synchronize()
			_t = _time.perf_counter()
			fstart = feature_pass(False, 0)
			if do_col:
				fstart = feature_pass(True, fstart)
			_t = _lap("features (RWR auto-stop + coverage + pooling)", _t)
			r = self.chrom2size[ds.chrom]
			# host randomized SVD exactly as the reference (:257-258, numpy global RNG); with cell
			# sharding the features are gathered to rank 0 (SURVEY.md 8e "init")
			if self.init_svd == "device":
				# cell-sharded randomized SVD on the device: only (features x k) sketches and k x k Grams are
				# all-reduced, the embedding rows stay with their cells (dist_svd.py)
				emb, _, _ = sharded_truncated_svd(feats[:, :fstart], r, n_iter=2, group=self.group if dist is not None else None,
				                                  seed=1000 + ci)
				if C is None:
					C = torch.zeros(ds.num_cell, int(cum[-1]), dtype=torch.float64, device=dev)
				C[:, cstart:cstart + emb.shape[1]] = emb
			else:
				f_host = feats[:, :fstart].cpu().numpy().astype(np.float64)
				if dist is not None:
					gathered = [None] * dist.get_world_size(self.group)
					dist.all_gather_object(gathered, f_host, group=self.group)
					f_host = np.concatenate(gathered, 0)
				if rank0:
					from sklearn.decomposition import TruncatedSVD  # host route only: the import alone costs ~1.8 s
					emb = TruncatedSVD(n_components=r, n_iter=2).fit_transform(f_host)
					if C is None:
						C = np.empty((f_host.shape[0], cum[-1]))
					C[:, cstart:cstart + emb.shape[1]] = emb
			cstart += r
			_t = _lap("per-chromosome SVD", _t)
			ni = torch.tensor([max(n_i_list) if n_i_list else 0], device=dev)
			if dist is not None:
				self._allreduce(ni, dist.ReduceOp.MAX)
			n_i_all.append(int(ni.item()))
			cov[cov <= 1e-4] = float("inf")
			self.bin_cov_list.append(cov[:ds.num_cell])
			self.bad_bin_cov_list.append(cov[ds.num_cell:] if nbad > 0 else 0)
			self._cov_all = getattr(self, "_cov_all", {})
			self._cov_all[ci] = cov
			del feats
		self.n_i = np.array(n_i_all)
		self._log("rwr iters:", self.n_i)
		_t = _time.perf_counter()
		# joint SVD of the per-chromosome embeddings (:283-290); one-off, cuSOLVER through torch
		if self.init_svd == "device":
			meta, SVh = sharded_svd_gram(C, R, self.group if dist is not None else None)
			meta, SVh = meta.float().contiguous(), SVh.float().contiguous()
		elif rank0:
			Ct = torch.from_numpy(C).float().to(dev)
			U, S, Vh = torch.linalg.svd(Ct, full_matrices=False)
			meta_all = U[:, :R].contiguous()
			SVh = (Vh[:R] * S[:R, None]).contiguous()
		if self.init_svd == "device":
			pass  # meta rows are already the local cells', SVh is replicated
		elif dist is not None:
			sizes_all = [None] * dist.get_world_size(self.group)
			dist.all_gather_object(sizes_all, self.num_cell, group=self.group)
			if not rank0:
				meta_all = torch.empty(sum(sizes_all), R, device=dev)
				SVh = torch.empty(R, int(cum[-1]), device=dev)
			dist.broadcast(meta_all, src=dist.get_global_rank(self.group, 0) if hasattr(dist, "get_global_rank") else 0, group=self.group)
			dist.broadcast(SVh, src=dist.get_global_rank(self.group, 0) if hasattr(dist, "get_global_rank") else 0, group=self.group)
			off = sum(sizes_all[:dist.get_rank(self.group)])
			meta = meta_all[off:off + self.num_cell].contiguous()
			# A, B must be identical on every rank
		else:
			meta = meta_all
		self.A_dev = [a.to(dev) for a in A_list]
		self.B_dict = {c: b.to(dev) for c, b in B_dict.items()}
		self.D_dict = {c: SVh[:, a:b].clone().contiguous() for c, a, b in zip(self.chrom2size, cum[:-1], cum[1:])}
		if dist is not None:
			src = dist.get_global_rank(self.group, 0) if hasattr(dist, "get_global_rank") else 0
			for t in self.A_dev + list(self.B_dict.values()):
				dist.broadcast(t, src=src, group=self.group)
		self.meta_embedding = meta
		_lap("joint SVD", _t)
		if _timing and rank0:
			print("[init timing] total %.2f s: " % (time.perf_counter() - t0) + ", ".join("%s %.2f s" % kv for kv in _acc.items()), flush=True)
		self._log(f"time elapsed: {time.perf_counter() - t0:.2f}")
		self._log("finish init")

	def load_state(self, A_list, B_list, D_list, meta_embedding, bin_cov_list, bad_bin_cov_list, n_i):
		"""Start from a given state (e.g. the reference's init, for lock-step parity runs)."""
		dev = self.device
		f = lambda x: torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).to(dev, torch.float32).contiguous().clone()
		self.A_dev = [f(a) for a in A_list]
		self.B_dict = {c: f(b) for c, b in zip(self.chrom2size, B_list)}
		self.D_dict = {c: f(d) for c, d in zip(self.chrom2size, D_list)}
		self.meta_embedding = f(meta_embedding)
		self.bin_cov_list = [f(b) for b in bin_cov_list]
		self.bad_bin_cov_list = [f(b) if (torch.is_tensor(b) or isinstance(b, np.ndarray)) and np.size(b) else 0
		                         for b in bad_bin_cov_list]
		self._cov_all = {}
		for ci, ds in enumerate(self.schic):
			good = self.bin_cov_list[ci]
			bad = self.bad_bin_cov_list[ci]
			if not torch.is_tensor(bad) and ds.total_cell_num > ds.num_cell:
				# bad-QC cells without a coverage table (the reference stores 0 for "none"): rows of inf, i.e. a 1/bin_cov column
				# scale of 0, instead of reading past the end of the good cells' table in a do_col transform
				bad = torch.full((ds.total_cell_num - ds.num_cell, good.shape[1]), float("inf"), dtype=torch.float32, device=dev)
			self._cov_all[ci] = torch.cat([good, bad], 0).contiguous() if torch.is_tensor(bad) else good
			self.bin_cov_list[ci] = self._cov_all[ci][:ds.num_cell]
		self.n_i = np.asarray(n_i)
		self._Z_valid = set()

	# ------------------------------------------------------------------------------------------
	def _impute_good(self, ci, b, do_conv, do_rwr, do_col):
		"""The imputed (cells, nb*ldw) matrix of block b; cached per sweep or per run."""
		key = (ci, b)
		ds = self.schic[ci]
		g = ds.geoms[b]
		ldw = pad4(g.w)
		X = self._X.get(key)
		if X is None:
			X = torch.zeros(ds.num_cell, g.nb * ldw, dtype=torch.float32, device=self.device)
			self._X[key] = X
			self._X_valid = getattr(self, "_X_valid", set())
		flags = (bool(do_conv), bool(do_rwr), bool(do_col))
		if getattr(self, "_X_flags", flags) != flags:  # transform() called with other flags than fit(): never reuse the old maps
			self.invalidate_cache()
			self._invalidate_Z()
		self._X_flags = flags
		if key not in self._X_valid:
			ev = getattr(self, "input_events", None)
			if ev is not None and ev.get(ci) is not None:  # block-CSR of this chromosome still in flight (H2D on another stream)
				torch.cuda.current_stream().wait_event(ev[ci])
			rwr_block_csr(ds, b, 0, ds.num_cell, X, g.nb * ldw, int(self.n_i[ci]), do_conv, do_rwr, do_col,
			              bin_cov=self._cov_all[ci] if do_col else None, use_tc=self.use_tc)
			self._X_valid.add(key)
		return X

	def _polar_table(self):
		"""Problem table of the per-bin polar step (one Gram matrix per bin, all chromosomes): offsets into one fp64
		buffer, sorted by decreasing size for fh_polar_isqrt_multi. Built once. Multi-GPU: the bins are independent
		problems that depend only on the all-reduced T1, so rank k takes a contiguous share [lo, hi) of the bins of EVERY
		block (sharding.polar_bin_range) - its Gram matrices, eigen-problems and G^{-1/2} products - and the inverse
		square roots are exchanged with ONE all-reduce over the zero-initialised buffer (x + 0 is exact: every rank ends
		with identical bits)."""
		if getattr(self, "_ptab", None) is not None:
			return self._ptab
		dev = self.device
		d = self._dist()
		world = d.get_world_size(self.group) if d is not None else 1
		rank = d.get_rank(self.group) if d is not None else 0
		block, own, n_list, off_list, slot_list, lengths = {}, {}, [], [], [], []
		off = slot = 0
		for ci, ds in enumerate(self.schic):
			r = self.chrom2size[ds.chrom]
			cnt = 0
			for b, g in enumerate(ds.geoms):
				ns = min(pad4(g.w), r)
				block[(ci, b)] = (off, ns)
				lo, hi = polar_bin_range(g.nb, world, rank)
				own[(ci, b)] = (lo, hi)
				for i in range(lo, hi):
					n_list.append(ns); off_list.append(off + i * ns * ns); slot_list.append(slot + i)
				off += g.nb * ns * ns
				slot += g.nb
				cnt += g.nb
			lengths.append(cnt)
		n_arr = np.asarray(n_list, dtype=np.int32)
		order = np.argsort(-n_arr, kind="stable")
		self._ptab = dict(block=block, own=own, world=world, count=len(n_list), n_host=np.ascontiguousarray(n_arr[order]),
		                  n_dev=torch.from_numpy(np.ascontiguousarray(n_arr[order])).to(dev),
		                  off_dev=torch.from_numpy(np.asarray(off_list, dtype=np.int64)[order].copy()).to(dev),
		                  slot_dev=torch.from_numpy(np.asarray(slot_list, dtype=np.int32)[order].copy()).to(dev),
		                  ssum=torch.zeros(slot, dtype=torch.float64, device=dev),
		                  nsweep=torch.zeros(slot, dtype=torch.int32, device=dev),  # Jacobi sweeps per bin (diagnostics)
		                  chrom_lengths=torch.tensor(lengths, device=dev),
		                  G=torch.zeros(off, dtype=torch.float64, device=dev), WT=torch.empty(off, dtype=torch.float64, device=dev))
		return self._ptab

	def _padded_factors(self, chrom):
		"""B (r x r) and D (R x r) copied to the row pitch round_up(r, 4) (zero pad columns)."""
		B, D = self.B_dict[chrom], self.D_dict[chrom]
		r = B.shape[0]
		rp = pad4(r)
		if rp == r:
			return B, D
		Bp = self._buf(("Bp", chrom), r, rp); Bp[:, :r].copy_(B)
		Dp = self._buf(("Dp", chrom), D.shape[0], rp); Dp[:, :r].copy_(D)
		return Bp, Dp

	def _build_W(self, ci, b):
		"""W (nb*ldw, R) with W_i = U_i B diag(A_i) D^T = (U_i B) (D diag(A_i))^T  (lhs of :422-430, transposed)."""
		ds = self.schic[ci]
		g = ds.geoms[b]
		dev, R = self.device, self.rank
		gd = self._gemm_dtype()
		r = self.chrom2size[ds.chrom]
		rp, ldw = pad4(r), pad4(g.w)
		P = g.nb * ldw
		Bp, Dp = self._padded_factors(ds.chrom)
		U = self.projection_dev[ci][b]
		Arows = self.A_dev[ci][g.row0:g.row0 + g.nb]
		UB = self._buf("UB", P, rp)
		_lib.gemm(U, Bp, UB, P, r, r, (rp, 1), (rp, 1), rp, dtype=gd)
		Dsc = self._buf("Dsc", g.nb, R, rp)
		_lib.scale_cols_batched(Dp, R, r, rp, Arows, g.nb, rp, Dsc)
		W = self._buf("W", P, R)
		_lib.gemm(UB, Dsc, W, ldw, R, r, (rp, 1), (1, rp), R, batch=g.nb, batch_strides=(ldw * rp, R * rp, ldw * R), dtype=gd)
		return W

	def _buf(self, key, *shape, dtype=torch.float32):
		"""Scratch buffer of the sweep, allocated (and zero-filled) ONCE per (key, shape): the timed sweep makes no allocator
		calls and no per-block zero-fills. Pad columns (row pitch round_up(r, 4) > r) are never written by the GEMMs (their
		stores stop at N), so they keep the zeros of the first fill. Buffers of equal key and shape are shared by successive
		blocks - safe because every use is ordered on the caller's stream."""
		k = (key, shape, dtype)
		t = self._scratch.get(k)
		if t is None:
			t = torch.zeros(*shape, dtype=dtype, device=self.device)
			self._scratch[k] = t
		return t

	def invalidate_cache(self):
		self._X_valid = set()

	def _invalidate_Z(self):
		"""Z = X^T V of the last P5 is only reusable while X (flags, step counts) and V are the ones it was formed from."""
		self._Z_valid = set()

	def release(self):
		"""Drop the resident imputed tensor and the scratch buffers (the factors stay)."""
		self._X = {}
		self._scratch = {}
		self.invalidate_cache()
		self._invalidate_Z()
		_lib.free_workspaces()

	# P1-P5: parafac2_intergrative.py:304-540
	@torch.no_grad()
	def update_meta_embedding_interactions(self, schic=None, projection_list=None, projected_tensor_list=None,
	                                       do_conv=True, do_rwr=True, do_col=False, first_iter=False):
		dev, R = self.device, self.rank
		gd = self._gemm_dtype()
		nch = len(self.schic)
		V = self.meta_embedding
		Cn = self.num_cell
		if self.cache == "sweep" or not hasattr(self, "_X_valid"):
			self.invalidate_cache()
			self.n_rwr_passes += 1
		MT = self._buf("MT", Cn, R).zero_()  # SVD_term^T (accumulated block by block)
		stats = self._buf("stats", 2 * nch + 1, dtype=torch.float64).zero_()  # x_U | ||X||^2 | x_V
		tab = self._polar_table()
		G_all, WT_all = tab["G"], tab["WT"]
		if tab["world"] > 1:
			G_all.zero_()  # the other ranks' slots must be exact zeros for the exchange by all-reduce
		temps = {}
		# Every r-wide operand lives at a row pitch rp = round_up(r, 4) (zero pad columns) so that TMA can
		# describe it and all GEMMs below except the fp64 ones run on the tcgen05 kernel.
		# ---- phase A: impute, P1, Gram of every temp_i. Sharded: the all-reduce of a block's T1 runs on the communicator's
		# stream under the NEXT block's RWR and P1 GEMM (software pipeline of depth one, two T1 buffers)
		def finish_block(pend):
			work, ci, b, T1, Bp = pend
			ds = self.schic[ci]
			g = ds.geoms[b]
			r = self.chrom2size[ds.chrom]
			rp, ldw = pad4(r), pad4(g.w)
			if work is not None:
				work.wait()
			t = self._tic()
			Arows = self.A_dev[ci][g.row0:g.row0 + g.nb]
			Bsc = self._buf("Bsc", g.nb, r, rp)
			_lib.scale_cols_batched(Bp, r, r, rp, Arows, g.nb, rp, Bsc)
			temp = self._buf(("temp", ci, b), g.nb, ldw, rp)
			_lib.gemm(T1, Bsc, temp, ldw, r, r, (rp, 1), (1, rp), rp, batch=g.nb, batch_strides=(ldw * rp, r * rp, ldw * rp), dtype=gd)
			temps[(ci, b)] = temp
			self._toc("p1_mttkrp", t)
			# P2a: Gram of every temp_i of this rank's bins in fp64 (tall: T^T T, wide: T T^T)
			t = self._tic()
			off, ns = tab["block"][(ci, b)]
			lo, hi = tab["own"][(ci, b)]
			if hi > lo:
				Gb = G_all[off + lo * ns * ns:off + hi * ns * ns]
				if ldw >= r:
					_lib.gemm(temp[lo:hi], temp[lo:hi], Gb, ns, ns, ldw, (1, rp), (rp, 1), ns, batch=hi - lo,
					          batch_strides=(ldw * rp, ldw * rp, ns * ns), dtype=_lib.GEMM_F32_ACC64, epilogue=_lib.EPI_SYMMETRIC)
				else:
					_lib.gemm(temp[lo:hi], temp[lo:hi], Gb, ns, ns, r, (rp, 1), (1, rp), ns, batch=hi - lo,
					          batch_strides=(ldw * rp, ldw * rp, ns * ns), dtype=_lib.GEMM_F32_ACC64, epilogue=_lib.EPI_SYMMETRIC)
			self._toc("polar_bins", t)

		pending, nblk = None, 0
		for ci, ds in enumerate(self.schic):
			r = self.chrom2size[ds.chrom]
			rp = pad4(r)
			Bp, Dp = self._padded_factors(ds.chrom)
			Cc = self._buf(("Cc", ci & 1), Cn, rp)
			if any((ci, b) not in self._Z_valid for b in range(len(ds.geoms))):
				_lib.gemm(V, Dp, Cc, Cn, r, R, (R, 1), (rp, 1), rp, dtype=gd)  # C = V D  (:336)
			for b, g in enumerate(ds.geoms):
				ldw = pad4(g.w)
				P = g.nb * ldw
				t = self._tic()
				X = self._impute_good(ci, b, do_conv, do_rwr, do_col)
				self._toc("rwr", t)
				if first_iter:
					_lib.check(_lib.lib().fh_sqnorm_accum(X.data_ptr(), 1, Cn * P, Cn * P, stats[nch + ci:].data_ptr(),
					                                      _lib.stream_ptr()))
				# P1: T1 = X^T C ; temp_i = T1_i (B diag(A_i))^T. C = V D with the V of the previous sweep's update, and the
				# previous sweep's P5 already formed Z = X^T V for exactly that V (X is re-imputed every sweep but is the same
				# deterministic function of the input), so T1 = Z D: a (P x R)(R x r) product instead of a third pass over X.
				# Only the first sweep (no Z yet) contracts X with C directly.
				t = self._tic()
				T1 = self._buf(("T1", nblk & 1), P, rp)
				nblk += 1
				if (ci, b) in self._Z_valid:
					_lib.gemm(self._buf(("Z", ci, b), P, R), Dp, T1, P, r, R, (R, 1), (rp, 1), rp, dtype=gd)
				else:
					_lib.gemm(X, Cc, T1, P, r, Cn, (1, P), (rp, 1), rp, dtype=gd)
				self._toc("p1_mttkrp", t)
				work = self._allreduce_async(T1)
				if pending is not None:
					finish_block(pending)
				pending = (work, ci, b, T1, Bp)
		if pending is not None:
			finish_block(pending)
		# Streaming input (bench.py's end-to-end leg, a caller that re-stages cells between sweeps): phase A holds the only
		# reads of the block-CSR in a sweep (phases C and P5 reuse the imputed X), so from here on the caller may overwrite
		# the datasets' device arrays with the NEXT sweep's input. `inputs_consumed_hook(event)` is called once per sweep with
		# an event recorded on the compute stream behind the last RWR launch; an upload stream that waits for it overlaps
		# the rest of this sweep (per-bin polar, P3, P4, P5, CP-ALS) instead of starting after the sweep's final read-back.
		hook = getattr(self, "inputs_consumed_hook", None)
		if hook is not None:
			consumed = torch.cuda.Event()
			consumed.record()
			hook(consumed)
		# ---- phase B: G^{-1/2} of all bins of all chromosomes at once (P2b)
		t = self._tic()
		if tab["world"] > 1:
			tab["ssum"].zero_()
		_lib.check(_lib.lib().fh_polar_isqrt_multi(G_all.data_ptr(), WT_all.data_ptr(), tab["n_dev"].data_ptr(),
		                                           tab["off_dev"].data_ptr(), tab["slot_dev"].data_ptr(),
		                                           tab["n_host"].ctypes.data, tab["count"], tab["ssum"].data_ptr(), 0,
		                                           tab["nsweep"].data_ptr(), _lib.stream_ptr()))
		# M = G^{-1/2} = WT^T WT of this rank's bins, over the Gram matrices (the other ranks' slots of G_all stay zero)
		for ci, ds in enumerate(self.schic):
			for b, g in enumerate(ds.geoms):
				off, ns = tab["block"][(ci, b)]
				lo, hi = tab["own"][(ci, b)]
				if hi > lo:
					nn = ns * ns
					WTb, Mb = WT_all[off + lo * nn:off + hi * nn], G_all[off + lo * nn:off + hi * nn]
					_lib.gemm(WTb, WTb, Mb, ns, ns, ns, (1, ns), (ns, 1), ns, batch=hi - lo, batch_strides=(nn, nn, nn), dtype=_lib.GEMM_F64,
					          epilogue=_lib.EPI_SYMMETRIC)  # G^{-1/2} = WT^T WT: symmetric, the tiles below the diagonal are mirrored
		if tab["world"] > 1:
			self._allreduce(G_all)
			self._allreduce(tab["ssum"])
		stats[:nch] = torch.segment_reduce(tab["ssum"], "sum", lengths=tab["chrom_lengths"])
		self._toc("polar_bins", t)
		# ---- phase C: U_i = temp_i M_i, P3
		for ci, ds in enumerate(self.schic):
			r = self.chrom2size[ds.chrom]
			rp = pad4(r)
			for b, g in enumerate(ds.geoms):
				ldw = pad4(g.w)
				P = g.nb * ldw
				X = self._impute_good(ci, b, do_conv, do_rwr, do_col)
				temp = temps.pop((ci, b))
				t = self._tic()
				off, ns = tab["block"][(ci, b)]
				nn = ns * ns
				Mb = G_all[off:off + g.nb * nn]
				U = self.projection_dev[ci][b]
				if ldw >= r:
					_lib.gemm(temp, Mb, U, ldw, r, r, (rp, 1), (ns, 1), rp, batch=g.nb, batch_strides=(ldw * rp, nn, ldw * rp),
					          dtype=_lib.GEMM_F32xF64_F32)
				else:
					_lib.gemm(Mb, temp, U, ldw, r, ldw, (ns, 1), (rp, 1), rp, batch=g.nb, batch_strides=(nn, ldw * rp, ldw * rp),
					          dtype=_lib.GEMM_F64xF32_F32)
				self._toc("polar_bins", t)
				t = self._tic()
				# P3: M += X W,  W_i = (U_i B) (D diag(A_i))^T
				W = self._build_W(ci, b)
				_lib.gemm(X, W, MT, Cn, R, P, (P, 1), (R, 1), R, beta=1.0, dtype=gd)
				self._toc("p3_project", t)
		# P4: V = polar(SVD_term^T) (:483-486)
		self.last_svd_term_T = MT
		t = self._tic()
		Vn = polar_tall(MT, self.group)
		_lib.check(_lib.lib().fh_dot_accum(Vn.data_ptr(), MT.data_ptr(), Cn, R, R, R, stats[2 * nch:].data_ptr(),
		                                   _lib.stream_ptr()))
		self._toc("polar_cells", t)
		self.meta_embedding = Vn
		# P5: Y_i = U_i^T X_i V (:488-529)
		y_works = []
		for ci, ds in enumerate(self.schic):
			r = self.chrom2size[ds.chrom]
			rp = pad4(r)
			Y = self.projected_dev[ds.chrom]
			for b, g in enumerate(ds.geoms):
				ldw = pad4(g.w)
				P = g.nb * ldw
				X = self._impute_good(ci, b, do_conv, do_rwr, do_col)
				t = self._tic()
				Z = self._buf(("Z", ci, b), P, R)  # kept: the next sweep's T1 = Z D
				_lib.gemm(X, Vn, Z, P, R, Cn, (1, P), (R, 1), R, dtype=gd)
				self._Z_valid.add((ci, b))
				U = self.projection_dev[ci][b]
				Yb = Y[ds.global_slice_bin.start + g.row0: ds.global_slice_bin.start + g.row0 + g.nb]
				_lib.gemm(U, Z, Yb, r, R, ldw, (1, rp), (R, 1), R, batch=g.nb, batch_strides=(ldw * rp, ldw * R, r * R), dtype=gd)
				self._toc("p5_tensor", t)
			if ci == self.chrom2id[ds.chrom][-1]:  # every dataset of this chromosome has added its bins: reduce Y under the next one's GEMMs
				y_works.append(self._allreduce_async(Y))
		for w in y_works:
			if w is not None:
				w.wait()
		if self._dist() is not None:
			loc = stats[nch:].clone()
			self._allreduce(loc)
			stats[nch:] = loc
		s = stats.cpu().numpy()
		x_U = s[:nch].reshape(-1, 1).copy()
		x_V = float(s[2 * nch])
		if first_iter:
			return self.projection_dev, self.projected_dev, x_U, x_V, s[nch:2 * nch].reshape(-1, 1).copy()
		return self.projection_dev, self.projected_dev, x_U, x_V

	def _core_norms(self):
		nch = len(self.schic)
		acc = torch.zeros(nch, dtype=torch.float64, device=self.device)
		for ci, ds in enumerate(self.schic):
			core_sqnorm_accum(self.A_dev[ci], self.B_dict[ds.chrom], self.D_dict[ds.chrom], acc[ci:])
		return acc.cpu().numpy().reshape(-1, 1)

	# ------------------------------------------------------------------------------------------
	# parafac2_intergrative.py:544-740
	def prepare(self, schic, size_ratio=0.3, do_conv=True, do_rwr=False, do_col=False, size_list=None,
	            run_init=True, state=None):
		"""Everything `fit` does before its sweep loop (:551-632)."""
		dev, R = self.device, self.rank
		_lib.require_cuda(getattr(self, "device", "cpu"), "Fast_Higashi_core (call .to('cuda') first)")
		_timing = os.environ.get("FH_INIT_TIMING") == "1"  # diagnostics: synchronises between the parts
		_tp = [time.perf_counter()]

		def _mark(what):
			if _timing:
				torch.cuda.synchronize()
				_tp.append(time.perf_counter())
				print("[prepare timing] %s %.2f s" % (what, _tp[-1] - _tp[-2]), flush=True)
		self._setup(schic, size_ratio, size_list)
		self._log("empty params initialized")
		_mark("_setup")
		if state is not None:
			self.load_state(*state)
		elif run_init:
			self.init_params(schic, do_conv, do_rwr, do_col)
		_mark("init_params / load_state")
		self._flags = (do_conv, do_rwr, do_col)
		self.projection_dev = [[torch.zeros(g.nb, pad4(g.w), pad4(self.chrom2size[ds.chrom]), dtype=torch.float32, device=dev)
		                        for g in ds.geoms] for ds in self.schic]
		self.projected_dev = {c: torch.zeros(self.chrom2num_bin[c], self.chrom2size[c], R, dtype=torch.float32, device=dev)
		                      for c in self.chrom2size}
		self._X, self._eig, self._ptab, self._scratch = {}, {}, None, {}
		self.invalidate_cache()
		self._invalidate_Z()
		self.n_rwr_passes = 0
		self._core_norm = self._core_norms()
		self._xnorm = None
		self.re_trace, self.sweep_seconds, self.loss_terms = [], [], []
		self._rec_errors = []
		_mark("buffers + core norms")

	def sweep_once(self, n_iter_parafac=1):
		"""One iteration of the reference's outer loop (:635-737): projections + V update + projected
		tensor, inner CP-ALS per chromosome, loss bookkeeping. Returns (re, per-chromosome re)."""
		do_conv, do_rwr, do_col = self._flags
		dist = self._dist()
		start_time = time.time()
		if self._xnorm is None:
			_, _, x_U, x_V, self._xnorm = self.update_meta_embedding_interactions(
				do_conv=do_conv, do_rwr=do_rwr, do_col=do_col, first_iter=True)
		else:
			_, _, x_U, x_V = self.update_meta_embedding_interactions(do_conv=do_conv, do_rwr=do_rwr, do_col=do_col)
		xnorm, core = self._xnorm, self._core_norm
		self.loss_terms.append(dict(xnorm=xnorm.ravel().copy(), core=core.ravel().copy(), x_U=x_U.ravel().copy(), x_V=x_V))
		err_U = xnorm + core - 2 * x_U
		err_V = xnorm.sum() + core.sum() - 2 * x_V
		# inner CP-ALS per chromosome (:674-695) + the new core norms (:697-706). The chromosomes are
		# independent and every kernel is tiny (r x r work, single-CTA solves): they are spread
		# round-robin over a few CUDA streams so they overlap instead of running back to back.
		t = self._tic()
		nch = len(self.schic)
		acc = torch.zeros(nch, dtype=torch.float64, device=self.device)
		main = torch.cuda.current_stream()
		if getattr(self, "_cp_streams", None) is None:
			n_st = int(os.environ.get("FH_CP_STREAMS", "11"))  # 0: everything on the caller's stream
			self._cp_streams = [torch.cuda.Stream(device=self.device) for _ in range(n_st)] if n_st > 0 else [main]
		ready = torch.cuda.Event()
		ready.record(main)
		# Sharded: the inner CP-ALS is cell independent (it only sees the all-reduced Y), so the chromosomes are dealt to the
		# ranks (largest first, round-robin) and the updated factors exchanged with ONE all-reduce over a packed buffer that
		# is zero where a rank does not own the chromosome (x + 0 is exact: replicas stay bit-identical)
		world = dist.get_world_size(self.group) if dist is not None else 1
		myrank = dist.get_rank(self.group) if dist is not None else 0
		by_cost = sorted(self.chrom2id, key=lambda c: -self.chrom2num_bin[c] * self.chrom2size[c] ** 2)
		owner = {c: i % world for i, c in enumerate(by_cost)}
		for k, (chrom, ids) in enumerate(self.chrom2id.items()):
			if owner[chrom] != myrank:
				continue
			st = self._cp_streams[k % len(self._cp_streams)]
			st.wait_event(ready)
			with torch.cuda.stream(st):
				tag = "cp%d" % (k % len(self._cp_streams))
				if len(ids) == 1:
					A = self.A_dev[ids[0]]
				else:
					A = torch.cat([self.A_dev[i] for i in ids], 0).contiguous()
				cp_als_(self.projected_dev[chrom], A, self.B_dict[chrom], self.D_dict[chrom], n_iter_parafac, tag=tag)
				if len(ids) > 1:
					for i in ids:
						self.A_dev[i].copy_(A[self.schic[i].global_slice_bin])
				if dist is None:
					for i in ids:
						core_sqnorm_accum(self.A_dev[i], self.B_dict[chrom], self.D_dict[chrom], acc[i:], tag=tag)
		for st in self._cp_streams:
			main.wait_stream(st)
		if dist is not None:
			parts, mine = [], []
			for chrom, ids in self.chrom2id.items():
				p_ = [self.A_dev[i] for i in ids] + [self.B_dict[chrom], self.D_dict[chrom]]
				parts += p_
				mine += [owner[chrom] == myrank] * len(p_)
			flat = torch.cat([p_.reshape(-1) if m else torch.zeros(p_.numel(), dtype=p_.dtype, device=p_.device) for p_, m in zip(parts, mine)])
			self._allreduce(flat)
			off = 0
			for p_ in parts:
				p_.copy_(flat[off:off + p_.numel()].view_as(p_))
				off += p_.numel()
			self._core_norm = self._core_norms()
		else:
			self._core_norm = acc.cpu().numpy().reshape(-1, 1)
		self._toc("cp_als", t)
		rec_error = float(np.sqrt(err_V) / np.sqrt(xnorm.sum()))
		self.re_trace.append(rec_error)
		self._rec_errors.append(np.sqrt(err_U.ravel()) / np.sqrt(xnorm.ravel()))
		self.sweep_seconds.append(time.time() - start_time)
		return rec_error, self._rec_errors[-1]

	def fit(self, schic, size_ratio=0.3, n_iter_max=2000, n_iter_parafac=5, do_conv=True, do_rwr=False,
	        do_col=False, tol=1e-8, size_list=None, gpu_id=None, verbose=True, run_init=True, state=None):
		self.gpu_id = gpu_id
		self.verbose = verbose
		self.prepare(schic, size_ratio, do_conv, do_rwr, do_col, size_list, run_init, state)
		for iteration in range(n_iter_max):
			if (iteration % 10) == 0 and iteration > 0 and n_iter_parafac < 10:
				n_iter_parafac += 1
			self._log("Starting iteration", iteration)
			rec_error, _ = self.sweep_once(n_iter_parafac)
			took = self.sweep_seconds[-1]
			if iteration >= 1:
				e, t = self._rec_errors, self.re_trace
				differences = (e[-2] ** 2 - e[-1] ** 2) / (e[-2] ** 2)
				total_differences = (t[-2] ** 2 - t[-1] ** 2) / t[-2] ** 2
				self._log(f"PARAFAC2 re={rec_error:.3f} {total_differences:.2e} "
				          f"variation min{differences.min().item():.1e} at chrom {differences.argmin().item():d}, "
				          f"max{differences.max().item():.1e} at chrom {differences.argmax().item():d}",
				          f"takes {took:.1f}s")
				if iteration >= 3 and tol > 0 and (total_differences < tol or differences.max() < tol * 2):
					self._log("converged in {} iterations.".format(iteration))
					break
			else:
				self._log(f"PARAFAC2 re={rec_error:.3f} takes {took:.1f}s")
		self._export()
		return self

	# optional per-stage device timers (bench.py): CUDA events on the current stream
	def enable_timers(self, on=True):
		self.timers = {} if on else None
		self._pending = []

	def _tic(self):
		if getattr(self, "timers", None) is None:
			return None
		e = torch.cuda.Event(enable_timing=True)
		e.record()
		return e

	def _toc(self, name, start):
		if start is None:
			return
		e = torch.cuda.Event(enable_timing=True)
		e.record()
		self._pending.append((name, start, e))

	def collect_timers(self):
		"""ms per stage since the last call (synchronises)."""
		torch.cuda.synchronize()
		out = {}
		for name, a, b in getattr(self, "_pending", []):
			out[name] = out.get(name, 0.0) + a.elapsed_time(b)
		self._pending = []
		return out

	def _export(self):
		"""Reference-shaped attributes: A_list on the host, projection_list[ci][b] (nb, w, r) on the
		host, projected_tensor_list[chrom] (n, r, R) on the host (:601-620)."""
		self.A_list = [a.cpu() for a in self.A_dev]
		self.projection_list = [[U[:, :g.w, :self.chrom2size[ds.chrom]].cpu() for U, g in zip(self.projection_dev[ci], ds.geoms)]
		                        for ci, ds in enumerate(self.schic)]
		self.projected_tensor_list = {c: y.cpu() for c, y in self.projected_dev.items()}

	# T1: parafac2_intergrative.py:742-834
	@torch.no_grad()
	def transform(self, schic=None, do_conv=None, do_rwr=None, do_col=None):
		f = self._flags
		do_conv = f[0] if do_conv is None else do_conv
		do_rwr = f[1] if do_rwr is None else do_rwr
		do_col = f[2] if do_col is None else do_col
		dev, R = self.device, self.rank
		gd = self._gemm_dtype()
		self._log("start transform")
		Ct = self.total_cell_num
		MT = torch.zeros(Ct, R, dtype=torch.float32, device=dev)
		for ci, ds in enumerate(self.schic):
			for b, g in enumerate(ds.geoms):
				ldw = pad4(g.w)
				P = g.nb * ldw
				W = self._build_W(ci, b)
				X = self._impute_good(ci, b, do_conv, do_rwr, do_col)
				_lib.gemm(X, W, MT, ds.num_cell, R, P, (P, 1), (R, 1), R, beta=1.0, dtype=gd)
				nbad = ds.total_cell_num - ds.num_cell
				if nbad > 0:
					chunk = cells_per_chunk(g.nb, ldw, 1 << 30)
					for c0 in range(0, nbad, chunk):
						nc = min(chunk, nbad - c0)
						Xb = torch.zeros(nc, P, dtype=torch.float32, device=dev)
						rwr_block_csr(ds, b, ds.num_cell + c0, nc, Xb, P, int(self.n_i[ci]), do_conv, do_rwr, do_col,
						              bin_cov=self._cov_all[ci] if do_col else None, use_tc=self.use_tc)
						_lib.gemm(Xb, W, MT[ds.num_cell + c0:], nc, R, P, (P, 1), (R, 1), R, beta=1.0, dtype=gd)
						del Xb
				del W
		meta = polar_tall(MT, self.group)
		self.A_list = [a.cpu() for a in self.A_dev]
		return (None, (self.A_list, self.B_dict.values(), self.D_dict.values(), meta), self.projection_list)

	def fit_transform(self, schic, size_ratio=0.3, n_iter_max=2000, n_iter_parafac=5, do_conv=True, do_rwr=False,
	                  do_col=False, tol=1e-8, size_list=None, gpu_id=None, verbose=True, run_init=True, state=None):
		if verbose:
			print("n_iter_parafac", n_iter_parafac)
		self.fit(schic, size_ratio, n_iter_max, n_iter_parafac, do_conv, do_rwr, do_col, tol, size_list, gpu_id,
		         verbose, run_init, state=state)
		return self.transform(schic, do_conv, do_rwr, do_col)
