"""Cell-sharded SVDs of the device init option (fasthigashi_b200.dist_svd, SURVEY.md 8f N1 / 8e "init"):
single process against an exact SVD and sklearn's TruncatedSVD (what parafac2_intergrative.py:257 calls),
and world_size 2 over gloo against the single-process result."""
import os
import socket
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from fasthigashi_b200.dist_svd import sharded_truncated_svd, sharded_svd_gram


def decaying_matrix(rows, cols, seed, power=1.0):
	g = torch.Generator().manual_seed(seed)
	k = min(rows, cols)
	U, _ = torch.linalg.qr(torch.randn(rows, k, generator=g, dtype=torch.float64))
	V, _ = torch.linalg.qr(torch.randn(cols, k, generator=g, dtype=torch.float64))
	s = 10.0 / torch.arange(1, k + 1, dtype=torch.float64) ** power
	return ((U * s) @ V.T).float(), s


def test_truncated_svd_matches_exact_and_is_no_worse_than_sklearn():
	from sklearn.decomposition import TruncatedSVD
	F, s = decaying_matrix(300, 700, 0)
	r = 24
	emb, S, Vt = sharded_truncated_svd(F, r, n_iter=2, seed=3)
	assert emb.shape == (300, r) and S.shape == (r,) and Vt.shape == (r, 700)
	np.testing.assert_allclose(S.numpy(), s[:r].numpy(), rtol=2e-2)            # 2 power iterations on a 1/j spectrum
	np.testing.assert_allclose(S[:8].numpy(), s[:8].numpy(), rtol=1e-4)
	gram = (emb / S).T @ (emb / S)
	assert torch.allclose(gram, torch.eye(r, dtype=gram.dtype), atol=1e-9)    # orthonormal left vectors
	assert torch.allclose(Vt @ Vt.T, torch.eye(r, dtype=Vt.dtype), atol=1e-9)
	resid = torch.linalg.norm(F.double() - emb @ Vt)
	best = torch.sqrt((s[r:] ** 2).sum())
	assert resid <= 1.02 * best
	np.random.seed(0)
	sk = TruncatedSVD(n_components=r, n_iter=2).fit(F.numpy().astype(np.float64))
	sk_emb = sk.transform(F.numpy().astype(np.float64))
	sk_resid = np.linalg.norm(F.numpy().astype(np.float64) - sk_emb @ sk.components_)
	assert float(resid) <= 1.005 * sk_resid                                   # captures at least as much as the reference's init
	# the leading directions agree with sklearn's up to sign
	for j in range(6):
		c = abs(np.corrcoef(emb[:, j].numpy(), sk_emb[:, j])[0, 1])
		assert c > 0.9999, (j, c)


def test_truncated_svd_edge_shapes():
	F, s = decaying_matrix(40, 12, 1)
	emb, S, Vt = sharded_truncated_svd(F, 12, n_iter=2)        # k capped at the feature count: exact
	np.testing.assert_allclose(S.numpy(), s.numpy(), rtol=1e-5)
	F0 = torch.zeros(10, 30)
	emb, S, Vt = sharded_truncated_svd(F0, 4)                  # all-zero features: finite output
	assert torch.isfinite(emb).all() and float(emb.abs().max()) == 0.0


def test_svd_gram_matches_dense_svd():
	C, s = decaying_matrix(500, 60, 2, power=0.7)
	R = 16
	U, SVh = sharded_svd_gram(C, R)
	Ue, Se, Vhe = torch.linalg.svd(C.double(), full_matrices=False)
	ref_SVh = Vhe[:R] * Se[:R, None]
	for j in range(R):
		sg = torch.sign((U[:, j].double() * Ue[:, j]).sum())
		assert torch.allclose(U[:, j].double() * sg, Ue[:, j], atol=2e-5), j
		assert torch.allclose(SVh[j].double() * sg, ref_SVh[j], atol=2e-5 * float(Se[0])), j
	assert torch.allclose(U.double() @ SVh.double(), (Ue[:, :R] * Se[:R]) @ Vhe[:R], atol=1e-4)


def _free_port():
	s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
	os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
	dist.init_process_group("gloo", rank=rank, world_size=world)
	F, _ = decaying_matrix(301, 400, 5)
	C, _ = decaying_matrix(301, 50, 6, power=0.7)
	from fasthigashi_b200.sharding import cell_slab
	lo, hi = cell_slab(301, world, rank)
	emb, S, Vt = sharded_truncated_svd(F[lo:hi], 20, n_iter=2, group=dist.group.WORLD, seed=9)
	U, SVh = sharded_svd_gram(C[lo:hi], 12, group=dist.group.WORLD)
	q.put((rank, lo, hi, emb.numpy(), S.numpy(), Vt.numpy(), U.numpy(), SVh.numpy()))
	dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process():
	ctx = mp.get_context("spawn")
	q = ctx.Queue()
	port = _free_port()
	procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
	for p in procs: p.start()
	res = sorted((q.get(timeout=180) for _ in procs), key=lambda x: x[0])
	for p in procs: p.join(timeout=60)
	F, _ = decaying_matrix(301, 400, 5)
	C, _ = decaying_matrix(301, 50, 6, power=0.7)
	emb1, S1, Vt1 = sharded_truncated_svd(F, 20, n_iter=2, seed=9)
	U1, SVh1 = sharded_svd_gram(C, 12)
	emb2 = np.concatenate([r[3] for r in res], 0)
	U2 = np.concatenate([r[6] for r in res], 0)
	assert [(r[1], r[2]) for r in res] == [(0, 151), (151, 301)]
	np.testing.assert_allclose(res[0][4], S1.numpy(), rtol=1e-9)
	np.testing.assert_allclose(res[0][4], res[1][4], rtol=0, atol=0)          # replicated results identical on both ranks
	np.testing.assert_allclose(res[0][5], res[1][5], rtol=0, atol=0)
	np.testing.assert_allclose(emb2, emb1.numpy(), atol=1e-8)
	np.testing.assert_allclose(res[0][5], Vt1.numpy(), atol=1e-9)
	np.testing.assert_allclose(U2, U1.numpy(), atol=1e-5)
	np.testing.assert_allclose(res[0][7], SVh1.numpy(), atol=1e-5)
