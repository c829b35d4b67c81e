"""Stage-by-stage comparison CUDA vs oracle on a mid-size synthetic problem (debug aid)."""
import math, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import fasthigashi_b200
from fasthigashi_b200 import synth, _lib
from fasthigashi_b200.sparse_for_schic import Sparse, Chrom_Dataset
from fasthigashi_b200.partial_rwr import rwr_block_csr, pad4
from fasthigashi_b200.project2orthogonal import polar_batched
from oracle import fh_oracle as O
from conftest import rel_fro
DEV = "cuda:0"
bins, ncell = [250, 160], 120
chroms, _ = synth.synth_dataset(bins, ncell, 0.10, off_diag=100, seed=1, num_cluster=5)
def mk(device):
	out = []
	for ch in chroms:
		n = ch["n"]; bb = math.ceil(n / max(math.ceil(n / 128), 1))
		out.append(Chrom_Dataset(Sparse(ch["indices"], ch["values"], ch["shape"]), bs_bin=bb, bs_cell=ncell, compact=True,
		                         flank=100, chrom=ch["chrom"], resolution=1000000, device=device))
	return out
cds, gds = mk("cpu"), mk(DEV)
gen = torch.Generator().manual_seed(0)
for ci in range(2):
	ds_c, ds_g = cds[ci], gds[ci]
	cov = torch.rand(ncell, ds_c.num_bin, generator=gen) + 0.5
	for b, g in enumerate(ds_c.geoms):
		ldw = pad4(g.w)
		for (do_col, k) in [(False, 4), (True, 4), (False, -1)]:
			out = torch.zeros(ncell, g.nb * ldw, device=DEV)
			n_it = rwr_block_csr(ds_g, b, 0, ncell, out, g.nb * ldw, k, True, True, do_col, bin_cov=cov.to(DEV))
			ref, n_ref = O.partial_rwr(O.densify_block(ds_c, b, 0, ncell), g.s, g.e, True, True, do_col, cov[:, g.col0:g.col0 + g.w], k)
			got = out.view(ncell, g.nb, ldw)[:, :, :g.w].cpu()
			print("rwr chrom%d blk%d nb=%d w=%d do_col=%d k=%d: rel=%.2e n_it=%d/%d" % (ci, b, g.nb, g.w, do_col, k, rel_fro(got.numpy(), ref.numpy()), n_it, n_ref))
		# polar on a realistic temp: X_i * random lhs
		X = ref.permute(1, 2, 0)  # (nb, w, c)
		r = 40
		lhs = torch.randn(g.nb, ncell, r, generator=gen)
		temp = torch.bmm(X, lhs)
		Ud, Sd, Vhd = torch.linalg.svd(temp.double(), full_matrices=False)
		truth = Ud @ Vhd
		kap = (Sd[:, 0] / Sd[:, -1])
		Uo, So = O.polar(temp, r)
		tp = torch.zeros(g.nb, ldw, r); tp[:, :g.w] = temp
		Ug, ssum, sig = polar_batched(tp.to(DEV), ldw, r, r, want_sigma=True)
		Ug = Ug.cpu()[:, :g.w]
		orth = (Ug.double().transpose(1, 2) @ Ug.double() - torch.eye(r, dtype=torch.float64)).abs().amax(dim=(1, 2))
		print("  polar: kappa med %.1e max %.1e | oracle-vs-truth %.2e | cuda-vs-truth %.2e | cuda ortho max %.1e | sigma_sum rel %.1e" % (
			kap.median(), kap.max(), rel_fro(Uo.numpy(), truth.numpy()), rel_fro(Ug.numpy(), truth.numpy()), orth.max(),
			float(((ssum.cpu() - Sd.sum(1)).abs() / Sd.sum(1)).max())))
		obj_t = (truth * temp.double()).sum(); obj_g = (Ug.double() * temp.double()).sum()
		print("  objective rel diff %.2e" % float(abs(obj_t - obj_g) / obj_t))
