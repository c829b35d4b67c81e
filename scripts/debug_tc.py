import sys, os, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fasthigashi_b200
from fasthigashi_b200 import _lib as L
DEV = "cuda:0"
torch.manual_seed(0)
def run(M, N, K, layout):
	A = torch.randn(M, K); B = torch.randn(K, N)
	lda = (K + 3) // 4 * 4 if layout[0] == "n" else (M + 3) // 4 * 4
	ldb = (N + 3) // 4 * 4 if layout[1] == "n" else (K + 3) // 4 * 4
	if layout[0] == "n":
		Ad = torch.zeros(M, lda); Ad[:, :K] = A; sa = (lda, 1)
	else:
		Ad = torch.zeros(K, lda); Ad[:, :M] = A.T; sa = (1, lda)
	if layout[1] == "n":
		Bd = torch.zeros(K, ldb); Bd[:, :N] = B; sb = (ldb, 1)
	else:
		Bd = torch.zeros(N, ldb); Bd[:, :K] = B.T; sb = (1, ldb)
	Cd = torch.full((M, N), float("nan"), device=DEV)
	L.gemm(Ad.to(DEV), Bd.to(DEV), Cd, M, N, K, sa, sb, N, dtype=L.GEMM_TF32X3)
	C = Cd.cpu().double()
	ref = A.double() @ B.double()
	err = float((C - ref).norm() / ref.norm())
	msg = "M%d N%d K%d %s: rel err %.2e" % (M, N, K, layout, err)
	if err > 1e-5:
		# which k-blocks are present? least squares on per-k-block partial products
		nkb = (K + 31) // 32
		parts = torch.stack([(A[:, kb*32:(kb+1)*32].double() @ B[kb*32:(kb+1)*32].double()).reshape(-1) for kb in range(nkb)], 1)
		coef = torch.linalg.lstsq(parts, C.reshape(-1, 1)).solution.ravel()
		res = float((parts @ coef - C.reshape(-1)).norm() / C.norm())
		msg += " | kblock coefs " + " ".join("%.2f" % c for c in coef.tolist()) + " | resid %.2e" % res
	print(msg)
for K in [32, 64, 96, 128, 160, 256]:
	run(128, 128, K, "nt")
for lay in ["nn", "tn", "tt"]:
	for K in [8, 32, 64]:
		run(128, 128, K, lay)
run(32, 32, 8, "nn"); run(32, 32, 8, "tn")

print("--- ragged / batched via pytest shapes")
def runb(M, N, K, batch, layout):
	g = torch.Generator().manual_seed(1)
	lda = (K + 3) // 4 * 4 if layout[0] == "n" else (M + 3) // 4 * 4
	ldb = (N + 3) // 4 * 4 if layout[1] == "n" else (K + 3) // 4 * 4
	A = torch.randn(batch, M, K, generator=g); B = torch.randn(batch, K, N, generator=g)
	ref = torch.bmm(A.double(), B.double())
	if layout[0] == "n":
		Ad = torch.zeros(batch, M, lda); Ad[:, :, :K] = A; sa = (lda, 1); ba = M * lda
	else:
		Ad = torch.zeros(batch, K, lda); Ad[:, :, :M] = A.transpose(1, 2); sa = (1, lda); ba = K * lda
	if layout[1] == "n":
		Bd = torch.zeros(batch, K, ldb); Bd[:, :, :N] = B; sb = (ldb, 1); bb = K * ldb
	else:
		Bd = torch.zeros(batch, N, ldb); Bd[:, :, :K] = B.transpose(1, 2); sb = (1, ldb); bb = N * ldb
	Cd = torch.full((batch, M, N), float("nan"), device=DEV)
	L.gemm(Ad.to(DEV), Bd.to(DEV), Cd, M, N, K, sa, sb, N, batch=batch, batch_strides=(ba, bb, M * N), dtype=L.GEMM_TF32X3)
	C = Cd.cpu().double()
	per = [(float((C[b] - ref[b]).norm() / ref[b].norm())) for b in range(batch)]
	print("M%d N%d K%d b%d %s:" % (M, N, K, batch, layout), " ".join("%.1e" % e for e in per), "nan:", int(torch.isnan(C).sum()))
runb(115, 115, 316, 3, "nt"); runb(128, 128, 316, 1, "nt"); runb(115, 128, 64, 1, "nt"); runb(128, 115, 64, 1, "nt"); runb(300, 137, 1000, 2, "nt")
runb(115, 115, 316, 3, "nn"); runb(300, 137, 1000, 2, "tn"); runb(64, 260, 40, 3, "tt")
