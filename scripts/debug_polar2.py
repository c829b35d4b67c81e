import sys, os, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fasthigashi_b200
from fasthigashi_b200 import _lib
from fasthigashi_b200.project2orthogonal import polar_batched
DEV = "cuda:0"
def al(x): return (x + 255) // 256 * 256
for sweeps in [8, 12, 16, 24, 40]:
	os.environ["FH_POLAR_SWEEPS"] = str(sweeps)
	g = torch.Generator().manual_seed(0)
	for (batch, rows, cols) in [(6, 316, 137), (3, 152, 150)]:
		for lk in [4.0, 5.5]:
			n = min(rows, cols)
			Uq, _ = torch.linalg.qr(torch.randn(batch, max(rows, cols), n, generator=g, dtype=torch.float64))
			Vq, _ = torch.linalg.qr(torch.randn(batch, n, n, generator=g, dtype=torch.float64))
			sv = torch.logspace(0, -lk, n, dtype=torch.float64)
			T = ((Uq * sv) @ Vq.transpose(1, 2)).float().contiguous()
			U, ssum, sig = polar_batched(T.to(DEV), rows, cols, cols, want_sigma=True)
			torch.cuda.synchronize()
			ws = _lib._ws_cache[(DEV, "polar")]
			m = (n + 1) & ~1
			off = 3 * al(batch * n * n * 8) + al(batch * n * 8) + al(batch * 40 * (m - 1) * (m // 2) * 16)
			ns = ws[off:off + 4 * batch].view(torch.int32).cpu().tolist()
			U = U.cpu().double()
			orth = ((U.transpose(1, 2) @ U) - torch.eye(n, dtype=torch.float64)).abs().amax(dim=(1, 2))
			print("cap", sweeps, (batch, rows, cols), "logk", lk, "nsweeps", ns, "ortho max %.1e" % float(orth.max()))
