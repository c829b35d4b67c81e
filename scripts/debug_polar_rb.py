"""Size sweep of the per-bin polar kernel (register-blocked Jacobi) on well-conditioned and graded inputs: orthogonality defect,
Jacobi sweeps, singular values against torch's fp64 SVD; and the Newton-Schulz iteration count / residual of the cell-mode polar."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fasthigashi_b200  # noqa
from fasthigashi_b200 import _lib
from fasthigashi_b200.project2orthogonal import polar_batched
g = torch.Generator().manual_seed(0)
for n in [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "4,8,9,12,16,20,32,33,40,64,65,96,100,128,129,137,144,150,160".split(","))]:
	for logk in (1.0, 5.0):
		rows, batch = 2 * n + 3, 3
		Uq, _ = torch.linalg.qr(torch.randn(batch, rows, n, generator=g, dtype=torch.float64))
		Vq, _ = torch.linalg.qr(torch.randn(batch, n, n, generator=g, dtype=torch.float64))
		sv = torch.logspace(0, -logk, n, dtype=torch.float64)
		T = ((Uq * sv) @ Vq.transpose(1, 2)).float().contiguous()
		Sd = torch.linalg.svdvals(T.double())
		U, ssum, sig, nsw = polar_batched(T.cuda(), rows, n, n, want_sigma=True, want_sweeps=True)
		U = U.cpu().double()
		err = float((U.transpose(1, 2) @ U - torch.eye(n, dtype=torch.float64)).abs().max())
		sg = torch.sort(sig.cpu(), dim=1, descending=True).values
		print("n %3d logk %.0f  ortho %.2e  sweeps %2d  sigma err %.2e  ssum err %.2e" % (
			n, logk, err, nsw, float(((sg - Sd).abs() / Sd[:, :1]).max()), float((ssum.cpu() - Sd.sum(1)).abs().max() / Sd.sum(1).max())))
# Newton-Schulz on a Gram matrix like SVD_term^T's (kappa(M) ~ 4e3)
for n, logk in ((64, 3.6), (256, 3.6), (256, 4.5)):
	Uq, _ = torch.linalg.qr(torch.randn(4000, n, generator=g, dtype=torch.float64))
	Vq, _ = torch.linalg.qr(torch.randn(n, n, generator=g, dtype=torch.float64))
	M = ((Uq * torch.logspace(0, -logk, n, dtype=torch.float64)) @ Vq.T)
	G = (M.T @ M).cuda()
	out = torch.empty_like(G)
	ws = _lib.workspace((5 * n * n + 512) * 8, G.device, "ns")
	it = C.c_int(0)
	rc = _lib.lib().fh_inv_sqrt_spd(G.data_ptr(), out.data_ptr(), n, ws.data_ptr(), ws.numel(), C.byref(it), _lib.stream_ptr())
	res = float(((out @ G @ out) - torch.eye(n, dtype=torch.float64, device="cuda")).norm() ** 2)
	print("NS n %d logk %.1f rc %d iters %d  ||I - Gi G Gi||_F^2 %.3e" % (n, logk, rc, it.value, res))
