"""Small multi-cell launch of the fused 3xFP16 RWR kernel (debugging aid: run under compute-sanitizer)."""
import sys, os, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fasthigashi_b200
from fasthigashi_b200 import synth
from fasthigashi_b200.partial_rwr import rwr_block_csr, pad4
from fasthigashi_b200.sparse_for_schic import Sparse, Chrom_Dataset
dev = torch.device("cuda:0")
cells, n = int(sys.argv[1]) if len(sys.argv) > 1 else 330, 250
ks = [int(a) for a in sys.argv[2:]] or [1, 4]
rng = np.random.default_rng(0)
idx, val = synth.synth_chrom(n, cells, 0.10, 100, 1, rng.integers(0, 4, cells), 4, device=dev, cell_chunk=256)
sp = Sparse.__new__(Sparse); sp.indices, sp.values, sp.shape, sp.ndim, sp.indptr = idx, val, np.array([n, n, cells]), 3, None
ds = Chrom_Dataset(sp, bs_bin=125, bs_cell=cells, compact=True, flank=100, chrom="chr1", resolution=1000000, device=dev)
for b, g in enumerate(ds.geoms):
	ldw = pad4(g.w); P = g.nb * ldw
	for k in ks:
		X = torch.zeros(cells, P, device=dev)
		rwr_block_csr(ds, b, 0, cells, X, P, k, True, True, False, use_tc=True, chunk=cells)
		torch.cuda.synchronize()
		print("block", b, "nb", g.nb, "w", g.w, "s", g.s, "k", k, "ok", float(X.sum()))
