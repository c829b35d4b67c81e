"""RWR stage time of one pass over all bin blocks for a BASELINE geometry other than the bench's (config 1: hg19 at 1 Mb with
do_col, the reference's GPU block rule gives blocks of <= 128 rows; config 5: hg19 at 100 kb, blocks of ~150 rows, which the
fused kernels do not take). The library reads its path switches once per process, so run it once per setting:
  python scripts/rwr_geometry_time.py 1000000 4238 1                    # fused 3xFP16 kernel (do_col inside)
  FH_RWR_FUSED=0 python scripts/rwr_geometry_time.py 1000000 4238 1     # per-op tcgen05 chain"""
import sys, os, math, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import fasthigashi_b200
from fasthigashi_b200 import synth, _lib
from fasthigashi_b200.partial_rwr import rwr_block_csr, pad4
from fasthigashi_b200.sparse_for_schic import Sparse, Chrom_Dataset
res, cells, do_col = int(sys.argv[1]), int(sys.argv[2]), bool(int(sys.argv[3]))
nchrom = int(sys.argv[4]) if len(sys.argv) > 4 else 22
dev = torch.device("cuda:0")
bins = synth.chrom_bins("hg19", res)[:nchrom]
rng = np.random.default_rng(0)
cluster = rng.integers(0, 8, size=cells)
dss = []
for ci, n in enumerate(bins):
	idx, val = synth.synth_chrom(n, cells, bench.DENSITY if res >= 500000 else 0.004, bench.OFF_DIAG, 1000 + ci, cluster, 8, device=dev, cell_chunk=256)
	sp = Sparse.__new__(Sparse); sp.indices, sp.values, sp.shape, sp.ndim, sp.indptr = idx, val, np.array([n, n, cells]), 3, None
	dss.append(Chrom_Dataset(sp, bs_bin=bench.bs_bin_rule(n, res), bs_cell=cells, compact=True, flank=bench.OFF_DIAG, chrom="chr%d" % (ci + 1), resolution=res, device=dev))
	del idx, val, sp
nbs = sorted({g.nb for ds in dss for g in ds.geoms})
big = max(g.nb * pad4(g.w) for ds in dss for g in ds.geoms)
X = torch.empty(cells, big, device=dev)
covs = [torch.rand(cells, ds.num_bin, device=dev) + 0.5 for ds in dss]


def one_pass():
	for ds, cov in zip(dss, covs):
		for b, g in enumerate(ds.geoms):
			rwr_block_csr(ds, b, 0, cells, X, g.nb * pad4(g.w), 4, True, True, do_col, bin_cov=cov, use_tc=True)


one_pass(); torch.cuda.synchronize()
fb0 = _lib.lib().fh_tc_fallback_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
	one_pass()
e1.record(); torch.cuda.synchronize()
print("res %d cells %d chromosomes %d do_col %d  blocks %d (rows %d..%d)  FH_RWR_FUSED=%s FH_RWR_F16=%s : %.2f ms per RWR pass, %.2f us per cell" % (
	res, cells, len(bins), do_col, sum(len(ds.geoms) for ds in dss), nbs[0], nbs[-1], os.environ.get("FH_RWR_FUSED", "2"), os.environ.get("FH_RWR_F16", "1"),
	e0.elapsed_time(e1) / 3, e0.elapsed_time(e1) / 3 * 1e3 / cells))
