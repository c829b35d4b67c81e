"""Phase stamps of the fused RWR kernel (FH_CHAIN_TRACE=1 must be set): one launch per k."""
import sys, os, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import fasthigashi_b200
from fasthigashi_b200 import synth
from fasthigashi_b200.partial_rwr import rwr_block_csr, pad4
from fasthigashi_b200.sparse_for_schic import Sparse, Chrom_Dataset
dev = torch.device("cuda:0")
cells, n = 2072, 457
rng = np.random.default_rng(0)
idx, val = synth.synth_chrom(n, cells, 0.05, 100, 1, rng.integers(0, 8, cells), 8, device=dev, cell_chunk=256)
sp = Sparse.__new__(Sparse); sp.indices, sp.values, sp.shape, sp.ndim, sp.indptr = idx, val, np.array([n, n, cells]), 3, None
ds = Chrom_Dataset(sp, bs_bin=bench.bs_bin_rule(n, bench.RES), bs_cell=cells, compact=True, flank=100, chrom="chr1", resolution=bench.RES, device=dev)
g = ds.geoms[1]; ldw = pad4(g.w); P = g.nb * ldw
X = torch.zeros(cells, P, device=dev)
for k in (1, 4, 4):
	rwr_block_csr(ds, 1, 0, cells, X, P, k, True, True, False, use_tc=True, chunk=cells)
torch.cuda.synchronize()
