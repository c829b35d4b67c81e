"""Cell-sharded truncated SVDs for `init_params` (SURVEY.md 8f N1, 8e "init").

The reference initialises the cell factor with two host SVDs over ALL cells
(parafac2_intergrative.py:257-258: sklearn `TruncatedSVD(n_components=r, n_iter=2)` of the pooled
1 Mb features per chromosome; :283-286: a dense SVD of the concatenated embeddings). With one
process per GPU that means gathering (cells x features) to one host - 50 GB per chromosome at 100k
cells. Here both factorisations work on the rank-local cell rows and exchange only small matrices:

  * `sharded_truncated_svd`  randomized range finder (Halko et al.; the algorithm behind sklearn's
    TruncatedSVD) with the cell axis sharded: the sketch `Y = F Omega` and every power iteration are
    local GEMMs; only `F^T Y` (features x k) is all-reduced. The orthonormal basis of the sharded
    `Y` comes from a Cholesky-QR2 on the all-reduced k x k Gram (fp64).
  * `sharded_svd_gram`       thin SVD of a tall cell-sharded matrix through its all-reduced Gram
    (fp64 eigh), for the joint embedding (cells x sum r).

Plain torch tensor algebra (GEMMs plus k x k Cholesky / eigh factorisations): one-off initialisation, not
the sweep. Works on any device and with `group=None` (single process), which is how the CPU tests
check it against an exact SVD and against sklearn. The default `init_params` path keeps the
reference's host SVD (bit-reproducible with its seed); this module is the `init_svd="device"` option.
"""
import torch


def _allreduce(t, group):
	if group is not None:
		import torch.distributed as dist
		dist.all_reduce(t, group=group)
	return t


def _small_cholesky(G):
	"""Lower Cholesky factor of a k x k matrix (k ~ 150), on the host: cuSOLVER's potrf / syevd take tens of milliseconds on
	matrices of this size, LAPACK well under one (measured: the 22 per-chromosome SVDs of init_params 1.8 s -> see DESIGN 7)."""
	return torch.linalg.cholesky(G.cpu()).to(G.device)


def _small_eigh(G):
	lam, W = torch.linalg.eigh(G.cpu())
	return lam.to(G.device), W.to(G.device)


def _orthonormalize_sharded(Y, group):
	"""Q with orthonormal columns spanning the (row-sharded) Y: two rounds of Cholesky-QR in fp64, the
	k x k Gram all-reduced. A tiny ridge keeps the factorisation defined for rank-deficient sketches."""
	for _ in range(2):
		G = _allreduce(Y.T @ Y, group)
		G = (G + G.T) * 0.5
		ridge = torch.finfo(G.dtype).eps * G.diagonal().max().clamp_min(1e-300) * G.shape[0]
		L = _small_cholesky(G + ridge * torch.eye(G.shape[0], dtype=G.dtype, device=G.device))
		Y = torch.linalg.solve_triangular(L, Y.T, upper=False).T
	return Y


def _svd_wide_gram(B):
	"""Thin SVD of the wide k x m matrix B (k ~ 150 sketch rows, m = thousands of features) through eigh of its k x k Gram
	in fp64: B = Uh diag(S) Vt. Both tall-skinny factorisations of the range finder (the QR of the features x k power
	iterate and this SVD) went through cuSOLVER's geqrf / gesvd, which took ~50 ms per chromosome whatever the cell count
	(1.1 s of a 1.4 s init at 2,500 cells); as GEMMs + a k x k eigh they are a few ms. The Gram squares the condition
	number: the relative error of component i is ~eps (S_0 / S_i)^2, 1e-10 for the leading components kept here."""
	G = B @ B.T
	lam, W = _small_eigh((G + G.T) * 0.5)
	lam, W = lam.flip(0), W.flip(1)
	S = lam.clamp_min(0).sqrt()
	floor = S[0] * 1e-12 if S.numel() else S
	Vt = (W.T @ B) / S.clamp_min(floor)[:, None]
	return W, S, Vt


@torch.no_grad()
def sharded_truncated_svd(F_local, n_components, n_iter=2, n_oversamples=10, group=None, seed=0):
	"""Rank-`n_components` randomized SVD of F = [F_0; F_1; ...] (rows = cells, sharded over the ranks
	of `group`). Returns (U_local * S, S, Vt): the embedding rows of the local cells (what
	`TruncatedSVD.fit_transform` returns), the singular values and the right singular vectors
	(replicated). The random test matrix comes from a seeded generator of F's device, so every rank draws
	the same one and a run is reproducible for a given world size."""
	F = F_local.to(torch.float64)
	n_feat = F.shape[1]
	k = min(int(n_components) + int(n_oversamples), n_feat)
	# drawn on the device of F (2.4 M doubles per chromosome from the CPU generator cost ~25 ms each): every rank seeds the
	# same generator of the same device type, so all ranks still draw the same matrix
	gen = torch.Generator(device=F.device).manual_seed(int(seed))
	omega = torch.randn(n_feat, k, generator=gen, dtype=torch.float64, device=F.device)
	Y = F @ omega
	for _ in range(int(n_iter)):
		# power iteration with the right factor re-orthonormalised (replicated, local QR): (F F^T)^q F Omega
		Z = _allreduce(F.T @ Y, group)
		Z = _orthonormalize_sharded(Z, None)           # replicated: Cholesky-QR2 without an exchange (see below)
		Y = F @ Z
	Q = _orthonormalize_sharded(Y, group)
	B = _allreduce(Q.T @ F, group)                     # k x features, replicated
	Uh, S, Vt = _svd_wide_gram(B)
	r = min(int(n_components), S.shape[0])
	U = Q @ Uh[:, :r]
	# deterministic signs (sklearn's svd_flip, v-based): largest |entry| of every right vector positive
	idx = Vt[:r].abs().argmax(dim=1)
	sign = torch.sign(Vt[:r][torch.arange(r, device=Vt.device), idx])
	sign[sign == 0] = 1
	return (U * sign) * S[:r], S[:r], Vt[:r] * sign[:, None]


@torch.no_grad()
def sharded_svd_gram(C_local, rank, group=None):
	"""Thin SVD of the tall cell-sharded C (cells x m) through eigh of the all-reduced m x m Gram in
	fp64: returns (U_local[:, :rank], S[:rank] * Vh[:rank]) = (meta_embedding rows, the `SVh` that
	parafac2_intergrative.py:285-290 splits into the D factors). Adequate here because kappa(C) is the
	spread of the leading singular values of the chromosome embeddings (~1e3-1e4)."""
	C = C_local.to(torch.float64)
	G = _allreduce(C.T @ C, group)
	lam, V = torch.linalg.eigh((G + G.T) * 0.5)
	lam, V = lam.flip(0), V.flip(1)
	r = min(int(rank), V.shape[1])
	S = lam[:r].clamp_min(0).sqrt()
	idx = V[:, :r].abs().argmax(dim=0)
	sign = torch.sign(V[idx, torch.arange(r, device=V.device)])
	sign[sign == 0] = 1
	V = V[:, :r] * sign
	floor = S[0] * 1e-12 if r else S
	U = (C @ V) / S.clamp_min(floor)
	return U.to(C_local.dtype), (V * S).T.to(C_local.dtype)
