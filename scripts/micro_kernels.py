"""Representative single launches of every hot kernel at BASELINE configs[1] sizes (for ncu)."""
import sys, os, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import fasthigashi_b200
from fasthigashi_b200 import _lib, synth
from fasthigashi_b200.partial_rwr import rwr_block_csr, pad4
from fasthigashi_b200.sparse_for_schic import Sparse, Chrom_Dataset
dev = torch.device("cuda:0")
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
n = 457
rng = np.random.default_rng(0)
idx, val = synth.synth_chrom(n, cells, 0.05, 100, 1, rng.integers(0, 8, cells), 8, device=dev, cell_chunk=256)
sp = Sparse.__new__(Sparse); sp.indices, sp.values, sp.shape, sp.ndim, sp.indptr = idx, val, np.array([n, n, cells]), 3, None
ds = Chrom_Dataset(sp, bs_bin=bench.bs_bin_rule(n, bench.RES), bs_cell=cells, compact=True, flank=100, chrom="chr1", resolution=bench.RES, device=dev)
b = 1
g = ds.geoms[b]; ldw = pad4(g.w); P = g.nb * ldw; r = 137; rp = pad4(r); R = 256
X = torch.zeros(cells, P, device=dev)
for rep in range(2):
	rwr_block_csr(ds, b, 0, cells, X, P, 4, True, True, False, use_tc=True)
torch.cuda.synchronize()
Cc = torch.randn(cells, rp, device=dev); V = torch.randn(cells, R, device=dev); W = torch.randn(P, R, device=dev)
T1 = torch.empty(P, rp, device=dev); MT = torch.zeros(cells, R, device=dev); Z = torch.empty(P, R, device=dev)
for rep in range(2):
	_lib.gemm(X, Cc, T1, P, r, cells, (1, P), (rp, 1), rp, dtype=_lib.GEMM_TF32X3)       # P1
	_lib.gemm(X, W, MT, cells, R, P, (P, 1), (R, 1), R, beta=1.0, dtype=_lib.GEMM_TF32X3)  # P3
	_lib.gemm(X, V, Z, P, R, cells, (1, P), (R, 1), R, dtype=_lib.GEMM_TF32X3)            # P5
torch.cuda.synchronize()
print("nb", g.nb, "w", g.w, "cells", cells, "nnz block", ds.val[b].numel(), "P1 flops %.3e P3 %.3e" % (2.0 * P * r * cells, 2.0 * P * R * cells))
