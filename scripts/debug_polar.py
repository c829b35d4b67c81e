import sys, os, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fasthigashi_b200
from fasthigashi_b200.project2orthogonal import polar_batched
DEV = "cuda:0"
g = torch.Generator().manual_seed(0)
for (batch, rows, cols) in [(6, 316, 137), (9, 72, 21), (3, 152, 150), (4, 12, 20), (5, 64, 32)]:
	for lk in [2.0, 4.0, 5.5]:
		n = min(rows, cols)
		Uq, _ = torch.linalg.qr(torch.randn(batch, max(rows, cols), n, generator=g, dtype=torch.float64))
		Vq, _ = torch.linalg.qr(torch.randn(batch, n, n, generator=g, dtype=torch.float64))
		sv = torch.logspace(0, -lk, n, dtype=torch.float64)
		T = (Uq * sv) @ Vq.transpose(1, 2)
		if rows < cols: T = T.transpose(1, 2)
		T = T.float().contiguous()
		Ud, Sd, Vhd = torch.linalg.svd(T.double(), full_matrices=False)
		truth = Ud @ Vhd
		U, ssum, sig = polar_batched(T.to(DEV), rows, cols, cols, want_sigma=True)
		U = U.cpu().double()
		G = U.transpose(1, 2) @ U if rows >= cols else U @ U.transpose(1, 2)
		orth = (G - torch.eye(n, dtype=torch.float64)).abs().amax(dim=(1, 2))
		sg = torch.sort(sig.cpu(), dim=1, descending=True).values
		print("shape", (batch, rows, cols), "logk", lk, "ortho", ["%.1e" % x for x in orth.tolist()],
		      "U-vs-truth %.1e" % float((U - truth).norm() / truth.norm()),
		      "sig rel max %.1e" % float(((sg - Sd).abs() / Sd).max()), "sig_min rel %.1e" % float(((sg[:, -1] - Sd[:, -1]).abs() / Sd[:, -1]).max()))
