"""TEST INFRASTRUCTURE ONLY - CPU restatement (torch-CPU / numpy, fp32 like the reference) of
Fast-Higashi's decomposition hot path. It is the checker for the CUDA path: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it. The
product package never does, and fails loudly without its CUDA library.

Parity pin: tests/golden/*.npz were produced by running the UNMODIFIED reference from
/root/reference (tests/golden/make_golden.py, via oracle/ref_shims.py); tests/test_oracle_golden.py
checks every function below against them, so the oracle is pinned to reference outputs generated in
the build container (the reference ships no tests or golden vectors of its own for this path -
SURVEY.md §4/§8c).

Each function cites the reference lines it restates.
"""
import math
import numpy as np
import torch
import torch.nn.functional as F

FLOOR = 1e-8   # sparse_for_schic.py:320, partial_rwr.py:80
EPS = 1e-15    # partial_rwr.py:90,91,111,134
RESTART = 0.5  # partial_rwr.py:109
MAX_RWR = 60   # partial_rwr.py:101
RWR_TOL = 0.01  # partial_rwr.py:122


# ----------------------------------------------------------------------------------------------
# S4: densify            sparse_for_schic.py:279-320 (transpose=True, do_conv=False branch)
# ----------------------------------------------------------------------------------------------
def densify_block(ds, b, c0, c1, device="cpu"):
	"""(c, nb, w) dense fp32 block of cells [c0, c1) from the block-CSR container; zeros filled,
	values scattered (no accumulation), floor 1e-8. `device`: where the stock-torch ops run ("cpu" for the
	checker; "cuda" only for bench.py's same-box stock-PyTorch baseline)."""
	g = ds.geoms[b]
	rp, col, val = ds.cell_range_csr(b, c0, c1)
	rp, col, val = rp.to(device).long(), col.to(device).long(), val.to(device)
	c = c1 - c0
	dense = torch.zeros(c * g.nb, g.w, dtype=torch.float32, device=device)
	rows = torch.repeat_interleave(torch.arange(c * g.nb, device=device), rp[1:] - rp[:-1])
	dense[rows, col] = val
	return dense.view(c, g.nb, g.w).clamp_(min=FLOOR)


# ----------------------------------------------------------------------------------------------
# R1-R4: partial RWR      partial_rwr.py:45-175
# ----------------------------------------------------------------------------------------------
def conv3x3(x):
	"""partial_rwr.py:77-81 - 3x3 mean, zero padding counted in the divisor, floor 1e-8;
	skipped when the block has a single row."""
	if x.shape[1] <= 1:
		return x
	return F.avg_pool2d(x[:, None], 3, 1, padding=1, ceil_mode=True)[:, 0].clamp_(min=FLOOR)


def transition_matrix(A, s, e):
	"""partial_rwr.py:84-97,111 - column-stochastic P (c, nb, nb) from the local affinity
	0.75 * first-order (diagonal block of A) + 0.25 * second-order (A A^T, zero diagonal)."""
	second = torch.bmm(A, A.transpose(1, 2))
	second.diagonal(dim1=-2, dim2=-1).zero_()
	first = A[:, :, s:e].clone()
	second = second / (second.sum(1, keepdim=True) + EPS) * 0.25
	first = first / (first.sum(1, keepdim=True) + EPS) * 0.75
	local = first + second
	empty = (local.sum(1) == 0).to(local.dtype)
	local.diagonal(dim1=-2, dim2=-1).add_(empty)
	return local / (local.sum(1, keepdim=True) + EPS)


def rwr_iterate(P, force_rwr_epochs):
	"""partial_rwr.py:99-129 - Q <- 0.5 Q P + 0.5 I. Forced mode: exactly k steps. Auto mode
	(k < 0): up to 60 steps; the step is applied, THEN the loop breaks if the largest per-cell
	Frobenius change is < 0.01; the returned count excludes the breaking step."""
	c, n, _ = P.shape
	eye = torch.eye(n, dtype=P.dtype, device=P.device)
	Q = eye[None].repeat(c, 1, 1)
	auto = force_rwr_epochs < 0
	steps = MAX_RWR if auto else int(force_rwr_epochs)
	count = 0
	for _ in range(steps):
		Qn = RESTART * torch.bmm(Q, P)
		Qn.diagonal(dim1=-2, dim2=-1).add_(1 - RESTART)
		if auto:
			delta = (Q - Qn).square().sum(dim=(1, 2)).sqrt()
			Q = Qn
			if float(delta.max()) < RWR_TOL:
				break
		else:
			Q = Qn
		count += 1
	return Q, count


def partial_rwr(x, s, e, do_conv, do_rwr, do_col, bin_cov=None, force_rwr_epochs=-1):
	"""partial_rwr.py:45-175 with final_transpose=False: x (c, nb, w) -> (imputed (c, nb, w), n_iter).
	`bin_cov` (c, w): per-cell coverage of the window columns, used only when do_col."""
	n_iter = 0
	if not (do_conv or do_rwr):
		return x, 0
	if do_conv:
		x = conv3x3(x)
	if do_rwr:
		A = x
		P = transition_matrix(A, s, e)
		Q, n_iter = rwr_iterate(P, force_rwr_epochs)
		if do_col:
			# partial_rwr.py:131-135: symmetrise, clamp at 0, ROW-normalise; A divided by coverage
			Q = (Q + Q.transpose(1, 2)) * 0.5
			Q = Q.clamp_(min=0.0)
			Q = Q / (Q.sum(2, keepdim=True) + EPS)
			A = A / bin_cov[:, None, :]
		x = torch.bmm(Q, A)
	return x, n_iter


# ----------------------------------------------------------------------------------------------
# P2/P4: polar factor      project2orthogonal.py:6-29 (CPU: gesvda raises -> default driver)
# ----------------------------------------------------------------------------------------------
def polar(matrix, rank=None):
	"""U Vh of the thin SVD and the leading singular values."""
	if rank is None:
		rank = min(matrix.shape[-2:])
	if matrix.is_cuda and matrix.shape[-2] / matrix.shape[-1] >= 0.5:
		# project2orthogonal.py:9-19 - the reference's GPU route: gesvda when tall enough, default driver on NaN / failure
		try:
			U, S, Vh = torch.linalg.svd(matrix, full_matrices=False, driver="gesvda")
			if not (torch.isfinite(U).all() and torch.isfinite(Vh).all()):
				raise RuntimeError("gesvda: non-finite")
		except Exception:
			U, S, Vh = torch.linalg.svd(matrix, full_matrices=False)
	else:
		U, S, Vh = torch.linalg.svd(matrix, full_matrices=False)
	return U[..., :rank] @ Vh[..., :rank, :], S[..., :rank]


# ----------------------------------------------------------------------------------------------
# C1: inner CP-ALS         parafac_integrative.py:12-112
# ----------------------------------------------------------------------------------------------
def balance_norm(factors):
	"""parafac_integrative.py:19-26 - unit columns everywhere, product of norms into the last."""
	norms = [torch.norm(f, dim=0) for f in factors]
	total = norms[0] * norms[1] * norms[2]
	out = [f / (nf + EPS) for f, nf in zip(factors, norms)]
	out[-1] = out[-1] * (total + EPS)
	return out


def cp_als(Y, factors, n_iter_max):
	"""parafac_integrative.py:28-112 - Y (n, r, R), factors [A (n,r), B (r,r), D (R,r)].
	Returns (factors, ||Xhat||^2, <Xhat, Y>)."""
	fac = balance_norm([f.clone().float() for f in factors])
	modes = ["ijk,jr,kr->ir", "ijk,ir,kr->jr", "ijk,ir,jr->kr"]
	prev = None
	norm_hat = inner = 0.0
	for _ in range(n_iter_max):
		for m in range(3):
			others = [fac[j] for j in range(3) if j != m]
			G = (others[0].T @ others[0]) * (others[1].T @ others[1])
			M = torch.einsum(modes[m], Y, *others)
			G.diagonal().add_(1e-10)
			fac[m] = torch.linalg.solve(G, M.T).T.contiguous()
		Xhat = torch.einsum("ir,jr,kr->ijk", *fac)
		loss = float(torch.linalg.norm(Y - Xhat).square())
		norm_hat = float(torch.linalg.norm(Xhat).square())
		inner = float((Xhat * Y).sum())
		wdiff = float("nan") if prev is None else (prev - loss) / prev
		prev = loss
		if wdiff < 1e-5:
			break
		fac = balance_norm(fac)
	return fac, norm_hat, inner


def core_sqnorm(A, B, D):
	"""||[[A,B,D]]||^2 = sum((A^T A) * (B^T B) * (D^T D)); the reference forms the dense tensor
	per bin batch (parafac2_intergrative.py:623-632) - same number."""
	return float(torch.einsum("ir,jr,kr->kij", A, B, D).square().sum())


# ----------------------------------------------------------------------------------------------
# I1 / P1-P5 / C2 / T1: the integrative PARAFAC2 driver   parafac2_intergrative.py:45-850
# ----------------------------------------------------------------------------------------------
class OracleCore:
	"""Restatement of Fast_Higashi_core on the block-CSR container (CPU, fp32)."""

	def __init__(self, rank, off_diag, res_list, device="cpu"):
		self.rank, self.off_diag, self.res_list = rank, off_diag, res_list
		self.device = torch.device(device)  # "cuda": the same stock-torch ops on the GPU (bench.py's torch_gpu_baseline only)

	# parafac2_intergrative.py:558-567
	def set_sizes(self, schic, size_ratio):
		self.chrom2size = {}
		for ds in schic:
			r = min(int(ds.num_bin * size_ratio * ds.resolution / 1000000), self.rank)
			self.chrom2size[ds.chrom] = min(self.chrom2size.get(ds.chrom, r), r)
		# multi-resolution stacking (:581-592): the datasets of one chromosome (one per resolution) share B and D and
		# their bins are stacked along mode 0 of that chromosome's projected tensor
		self.chrom2id = {c: [] for c in self.chrom2size}
		self.chrom2num_bin = {}
		self.gslice = []
		for ci, ds in enumerate(schic):
			self.chrom2id[ds.chrom].append(ci)
			start = self.chrom2num_bin.get(ds.chrom, 0)
			self.gslice.append(slice(start, start + ds.num_bin))
			self.chrom2num_bin[ds.chrom] = start + ds.num_bin

	def _imputed(self, ds, ci, b, c0, c1, do_conv, do_rwr, do_col, bad=False, k=None):
		g = ds.geoms[b]
		off = ds.num_cell if bad else 0
		x = densify_block(ds, b, off + c0, off + c1, self.device)
		cov = None
		if do_col:
			src = self.bad_bin_cov_list[ci] if bad else self.bin_cov_list[ci]
			cov = src[c0:c1, g.col0:g.col0 + g.w]
		return partial_rwr(x, g.s, g.e, do_conv, do_rwr, do_col, cov,
		                   force_rwr_epochs=self.n_i[ci] if k is None else k)[0]

	# parafac2_intergrative.py:61-301
	def init_params(self, schic, do_conv, do_rwr, do_col):
		from sklearn.decomposition import TruncatedSVD
		R = self.rank
		sizes = [self.chrom2size[ds.chrom] for ds in schic]
		self.A_list = [torch.randn(ds.num_bin, r) * 1e-2 + 1 for ds, r in zip(schic, sizes)]
		self.B_dict = {ch: torch.eye(r).add_(torch.randn(r), alpha=1e-2) for ch, r in self.chrom2size.items()}
		uniq = list(self.chrom2size.values()) * len(self.res_list)
		cum = np.concatenate([[0], np.cumsum(uniq)])
		C = None
		self.bin_cov_list, self.bad_bin_cov_list, n_i_all = [], [], []
		for ci, ds in enumerate(schic):
			if C is None:
				C = np.empty((ds.num_cell, cum[-1])); cstart = 0
			nbad = ds.total_cell_num - ds.num_cell
			cov = torch.full((ds.num_cell, ds.num_bin), 1e-4)
			bad_cov = torch.full((nbad, ds.num_bin), 1e-4)
			n1m = int(math.ceil(ds.num_bin * ds.resolution / 1000000))
			size1 = min(int(math.ceil(ds.num_bin / ds.num_bin_batch * ds.resolution / 1000000))
			            + 2 * self.off_diag + 1, n1m)
			feats = np.empty((ds.num_cell, int(math.ceil(n1m * size1))))
			ll = int(math.ceil(1000000 / ds.resolution))
			pool = (lambda t: F.avg_pool2d(t[:, None], ll, ll, padding=0, ceil_mode=False)[:, 0]) if ll > 1 else (lambda t: t)
			n_i_list = []
			fstart = 0
			for b, g in enumerate(ds.geoms):
				width = 0
				for sl in ds.cell_slice_list[:ds.num_cell_batch]:
					x, n_i = partial_rwr(densify_block(ds, b, sl.start, sl.stop), g.s, g.e,
					                     do_conv, do_rwr, False, None, -1)
					cov[sl, g.col0:g.col0 + g.w] += x.sum(1)
					n_i_list.append(n_i)
					if not do_col:
						Bf = pool(x).reshape(x.shape[0], -1).numpy()
						feats[sl, fstart:fstart + Bf.shape[1]] = Bf
						width = Bf.shape[1]
				fstart += width
				for sl in ds.cell_slice_list[ds.num_cell_batch:]:
					x, _ = partial_rwr(densify_block(ds, b, sl.start, sl.stop), g.s, g.e,
					                   do_conv, do_rwr, False, None, -1)
					bad_cov[sl.start - ds.num_cell:sl.stop - ds.num_cell, g.col0:g.col0 + g.w] += x.sum(1)
			if do_col:
				# second pass with the (not yet inf-masked) coverage; auto-stop again (:201-247)
				for b, g in enumerate(ds.geoms):
					width = 0
					for sl in ds.cell_slice_list[:ds.num_cell_batch]:
						x, _ = partial_rwr(densify_block(ds, b, sl.start, sl.stop), g.s, g.e, do_conv, do_rwr,
						                   True, cov[sl, g.col0:g.col0 + g.w], -1)
						Bf = pool(x).reshape(x.shape[0], -1).numpy()
						feats[sl, fstart:fstart + Bf.shape[1]] = Bf
						width = Bf.shape[1]
					fstart += width
			r = self.chrom2size[ds.chrom]
			emb = TruncatedSVD(n_components=r, n_iter=2).fit_transform(feats[:, :fstart])
			C[:, cstart:cstart + emb.shape[1]] = emb
			cstart += emb.shape[1]
			n_i_all.append(max(n_i_list) if n_i_list else 0)
			cov[cov <= 1e-4] = float("inf")
			if ds.num_cell_batch_bad > 0:
				bad_cov[bad_cov <= 1e-4] = float("inf")
				self.bad_bin_cov_list.append(bad_cov)
			else:
				self.bad_bin_cov_list.append(0)
			self.bin_cov_list.append(cov)
		self.n_i = np.array(n_i_all)
		U, S, Vh = torch.linalg.svd(torch.from_numpy(C).float(), full_matrices=False)
		self.meta_embedding = U[:, :R].contiguous()
		SVh = Vh[:R] * S[:R, None]
		self.D_dict = {ch: SVh[:, a:b].clone() for ch, a, b in zip(self.chrom2size, cum[:-1], cum[1:])}

	def load_state(self, A_list, B_list, D_list, meta_embedding, bin_cov_list, bad_bin_cov_list, n_i):
		chroms = list(self.chrom2size)
		dev = self.device
		self.A_list = [torch.as_tensor(a).clone().float().to(dev) for a in A_list]
		self.B_dict = {c: torch.as_tensor(b).clone().float().to(dev) for c, b in zip(chroms, B_list)}
		self.D_dict = {c: torch.as_tensor(d).clone().float().to(dev) for c, d in zip(chroms, D_list)}
		self.meta_embedding = torch.as_tensor(meta_embedding).clone().float().to(dev)
		self.bin_cov_list = [torch.as_tensor(b).float().to(dev) for b in bin_cov_list]
		self.bad_bin_cov_list = [torch.as_tensor(b).float().to(dev) if not np.isscalar(b) else 0 for b in bad_bin_cov_list]
		self.n_i = np.asarray(n_i)

	# parafac2_intergrative.py:304-540
	def sweep_projections(self, schic, do_conv, do_rwr, do_col, want_norm):
		V = self.meta_embedding
		R = self.rank
		svd_term = torch.zeros(R, V.shape[0], device=self.device)
		x_U = np.zeros(len(schic)); xnorm = np.zeros(len(schic))
		self.projection_list = []
		for ci, ds in enumerate(schic):
			A, B, D = self.A_list[ci], self.B_dict[ds.chrom], self.D_dict[ds.chrom]
			Cc = V @ D
			proj = []
			for b, g in enumerate(ds.geoms):
				rows = slice(g.row0, g.row0 + g.nb)
				temp = 0
				Xs = []
				for sl in ds.cell_slice_list[:ds.num_cell_batch]:
					X = self._imputed(ds, ci, b, sl.start, sl.stop, do_conv, do_rwr, do_col).permute(1, 2, 0)
					if want_norm:
						xnorm[ci] += float(torch.linalg.norm(X).square())
					lhs = torch.einsum("ir,jr,kr->ikj", A[rows], B, Cc[sl])
					temp = temp + torch.bmm(X, lhs)
					Xs.append((sl, X))
				U, S = polar(temp, temp.shape[-1])
				proj.append(U)
				x_U[ci] += float(S.sum())
				lhs2 = torch.einsum("ir,jr,kr->kij", A[rows], B, D).reshape(R, -1)
				for sl, X in Xs:
					p = torch.bmm(U.transpose(1, 2), X)
					svd_term[:, sl] += lhs2 @ p.reshape(-1, p.shape[-1])
			self.projection_list.append(proj)
		return svd_term, x_U, xnorm

	def sweep(self, schic, do_conv, do_rwr, do_col, want_norm=False):
		svd_term, x_U, xnorm = self.sweep_projections(schic, do_conv, do_rwr, do_col, want_norm)
		self.last_svd_term = svd_term
		V, _ = polar(svd_term.T, self.rank)
		x_V = float((V * svd_term.T).sum())
		self.meta_embedding = V
		self.projected = {c: torch.zeros(self.chrom2num_bin[c], self.chrom2size[c], self.rank, device=self.device) for c in self.chrom2size}
		for ci, ds in enumerate(schic):
			Y = self.projected[ds.chrom][self.gslice[ci]]   # this resolution's rows of the stacked tensor (:526-529)
			for b, g in enumerate(ds.geoms):
				acc = 0
				for sl in ds.cell_slice_list[:ds.num_cell_batch]:
					X = self._imputed(ds, ci, b, sl.start, sl.stop, do_conv, do_rwr, do_col).permute(1, 2, 0)
					acc = acc + torch.einsum("ijk,km,ijl->ilm", X, V[sl], self.projection_list[ci][b])
				Y[g.row0:g.row0 + g.nb] = acc
		return x_U, x_V, xnorm

	def core_norms(self, schic):
		return np.array([core_sqnorm(self.A_list[ci], self.B_dict[ds.chrom], self.D_dict[ds.chrom])
		                 for ci, ds in enumerate(schic)])

	# parafac2_intergrative.py:544-740
	def fit(self, schic, size_ratio, n_iter_max, n_iter_parafac, do_conv, do_rwr, do_col, tol,
	        run_init=True, state=None):
		self.set_sizes(schic, size_ratio)
		if state is not None:
			self.load_state(*state)
		elif run_init:
			self.init_params(schic, do_conv, do_rwr, do_col)
		core = self.core_norms(schic)
		self.re_trace, per_chrom, self.loss_terms = [], [], []
		xnorm = None
		for it in range(n_iter_max):
			if it % 10 == 0 and it > 0 and n_iter_parafac < 10:
				n_iter_parafac += 1
			x_U, x_V, xn = self.sweep(schic, do_conv, do_rwr, do_col, want_norm=xnorm is None)
			if xnorm is None:
				xnorm = xn
			self.loss_terms.append(dict(xnorm=xnorm.copy(), core=core.copy(), x_U=x_U.copy(), x_V=x_V))
			err_U = xnorm + core - 2 * x_U
			err_V = xnorm.sum() + core.sum() - 2 * x_V
			for chrom, ids in self.chrom2id.items():   # one CP-ALS per chromosome over the stacked bins (:670-695)
				fac, _, _ = cp_als(self.projected[chrom], [torch.cat([self.A_list[i] for i in ids], 0), self.B_dict[chrom],
				                                           self.D_dict[chrom]], n_iter_parafac)
				for i in ids:
					self.A_list[i] = fac[0][self.gslice[i]].clone()
				self.B_dict[chrom], self.D_dict[chrom] = fac[1], fac[2]
			core = self.core_norms(schic)
			self.re_trace.append(float(np.sqrt(err_V) / np.sqrt(xnorm.sum())))
			per_chrom.append(np.sqrt(err_U) / np.sqrt(xnorm))
			if it >= 1:
				d_chrom = (per_chrom[-2] ** 2 - per_chrom[-1] ** 2) / per_chrom[-2] ** 2
				d_tot = (self.re_trace[-2] ** 2 - self.re_trace[-1] ** 2) / self.re_trace[-2] ** 2
				if it >= 3 and tol > 0 and (d_tot < tol or d_chrom.max() < tol * 2):
					break
		return self

	# parafac2_intergrative.py:742-834
	def transform(self, schic, do_conv, do_rwr, do_col):
		R = self.rank
		svd_term = torch.zeros(R, schic[0].total_cell_num, device=self.device)
		for ci, ds in enumerate(schic):
			A, B, D = self.A_list[ci], self.B_dict[ds.chrom], self.D_dict[ds.chrom]
			for b, g in enumerate(ds.geoms):
				lhs2 = torch.einsum("ir,jr,kr->kij", A[g.row0:g.row0 + g.nb], B, D).reshape(R, -1)
				U = self.projection_list[ci][b]
				for sl in ds.cell_slice_list[:ds.num_cell_batch]:
					X = self._imputed(ds, ci, b, sl.start, sl.stop, do_conv, do_rwr, do_col).permute(1, 2, 0)
					p = torch.bmm(U.transpose(1, 2), X)
					svd_term[:, sl] += lhs2 @ p.reshape(-1, p.shape[-1])
				for sl in ds.cell_slice_list[ds.num_cell_batch:]:
					X = self._imputed(ds, ci, b, sl.start - ds.num_cell, sl.stop - ds.num_cell,
					                  do_conv, do_rwr, do_col, bad=True).permute(1, 2, 0)
					p = torch.bmm(U.transpose(1, 2), X)
					svd_term[:, sl] += lhs2 @ p.reshape(-1, p.shape[-1])
		V, _ = polar(svd_term.T, R)
		return V


def embed_all(V, D_list):
	"""FastHigashi_Wrapper.py:755-761 - concat over chromosomes of V @ colnorm(D)."""
	out = []
	for D in D_list:
		D = np.asarray(D, dtype=np.float64)
		out.append(np.asarray(V, dtype=np.float64) @ (D / np.linalg.norm(D, axis=0, keepdims=True)))
	return np.concatenate(out, axis=1)
