"""Time the RWR of one chr1-sized bin block (nb=115, w=315) for k = 0 (conv only), 1, 2, 4 steps:
the differences give the per-phase cost of the fused kernel. usage: python scripts/rwr_phase_times.py [cells]"""
import sys, os, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import fasthigashi_b200
from fasthigashi_b200 import synth
from fasthigashi_b200.partial_rwr import rwr_block_csr, pad4
from fasthigashi_b200.sparse_for_schic import Sparse, Chrom_Dataset
dev = torch.device("cuda:0")
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 2072
n = 457
rng = np.random.default_rng(0)
idx, val = synth.synth_chrom(n, cells, 0.05, 100, 1, rng.integers(0, 8, cells), 8, device=dev, cell_chunk=256)
sp = Sparse.__new__(Sparse); sp.indices, sp.values, sp.shape, sp.ndim, sp.indptr = idx, val, np.array([n, n, cells]), 3, None
ds = Chrom_Dataset(sp, bs_bin=bench.bs_bin_rule(n, bench.RES), bs_cell=cells, compact=True, flank=100, chrom="chr1", resolution=bench.RES, device=dev)
b = 1
g = ds.geoms[b]; ldw = pad4(g.w); P = g.nb * ldw
X = torch.zeros(cells, P, device=dev)


def t(k, do_rwr=True, reps=10):
	for _ in range(3):
		rwr_block_csr(ds, b, 0, cells, X, P, k, True, do_rwr, False, use_tc=True, chunk=cells)
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	torch.cuda.synchronize(); e0.record()
	for _ in range(reps):
		rwr_block_csr(ds, b, 0, cells, X, P, k, True, do_rwr, False, use_tc=True, chunk=cells)
	e1.record(); torch.cuda.synchronize()
	return e0.elapsed_time(e1) * 1e3 / reps


print("FH_RWR_FUSED", os.environ.get("FH_RWR_FUSED", "2"), "nb", g.nb, "w", g.w, "cells", cells)
base = t(0, do_rwr=False)
print("densify+conv only: %.1f us" % base)
for k in (1, 2, 3, 4, 6):
	us = t(k)
	print("k=%d: %.1f us total, %.1f us RWR part, %.2f us per cell per SM" % (k, us, us - base, (us - base) / (cells / 148.0)))
