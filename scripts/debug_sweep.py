"""One sweep CUDA vs oracle from the same state: compares every loss term and factor (debug aid)."""
import math, sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fasthigashi_b200
from fasthigashi_b200 import synth
from fasthigashi_b200.sparse_for_schic import Sparse, Chrom_Dataset
from fasthigashi_b200.parafac2_intergrative import Fast_Higashi_core
from oracle import fh_oracle as O
from conftest import rel_fro
DEV = "cuda:0"
bins, ncell = [250, 160, 100, 60], 120
do_col = "--nocol" not in sys.argv
chroms, _ = synth.synth_dataset(bins, ncell, 0.10, off_diag=100, seed=1, num_cluster=5)
def mk(device):
	out = []
	for ch in chroms:
		n = ch["n"]; bb = math.ceil(n / max(math.ceil(n / 128), 1))
		out.append(Chrom_Dataset(Sparse(ch["indices"], ch["values"], ch["shape"]), bs_bin=bb, bs_cell=ncell, compact=True,
		                         flank=100, chrom=ch["chrom"], resolution=1000000, device=device))
	return out
ocore = O.OracleCore(32, 100, [1000000])
torch.manual_seed(0); np.random.seed(0)
ods = mk("cpu")
ocore.set_sizes(ods, 0.3)
ocore.init_params(ods, True, True, do_col)
state = ([a.clone() for a in ocore.A_list], [b.clone() for b in ocore.B_dict.values()],
         [d.clone() for d in ocore.D_dict.values()], ocore.meta_embedding.clone(),
         [c.clone() for c in ocore.bin_cov_list], [0] * 4, ocore.n_i.copy())
print("n_i", ocore.n_i)
x_U, x_V, xnorm = ocore.sweep(ods, True, True, do_col, want_norm=True)
core = Fast_Higashi_core(32, 100, [1000000]).to(DEV)
core.verbose = False
gds = mk(DEV)
core._setup(gds, 0.3, None)
core.load_state(*state)
core._flags = (True, True, do_col)
from fasthigashi_b200.partial_rwr import pad4
core.projection_dev = [[torch.zeros(g.nb, pad4(g.w), core.chrom2size[ds.chrom], device=DEV) for g in ds.geoms] for ds in core.schic]
core.projected_dev = {c: torch.zeros(core.chrom2num_bin[c], core.chrom2size[c], 32, device=DEV) for c in core.chrom2size}
core.invalidate_cache()
_, _, gx_U, gx_V, gxnorm = core.update_meta_embedding_interactions(do_conv=True, do_rwr=True, do_col=do_col, first_iter=True)
print("xnorm  oracle", xnorm, "\n       cuda  ", gxnorm.ravel(), "\n rel", np.abs(gxnorm.ravel() - xnorm) / xnorm)
print("x_U    oracle", x_U, "\n       cuda  ", gx_U.ravel(), "\n rel", np.abs(gx_U.ravel() - x_U) / x_U)
print("x_V    oracle", x_V, " cuda", gx_V, " rel", abs(gx_V - x_V) / x_V)
print("MT rel", rel_fro(core.last_svd_term_T.cpu().numpy(), ocore.last_svd_term.T.numpy()))
print("V rel", rel_fro(core.meta_embedding.cpu().numpy(), ocore.meta_embedding.numpy()))
for ci, ds in enumerate(ods):
	print("Y rel chrom", ci, rel_fro(core.projected_dev[ds.chrom].cpu().numpy(), ocore.projected[ds.chrom].numpy()))
	for b, g in enumerate(ds.geoms):
		Ug = core.projection_dev[ci][b][:, :g.w].cpu()
		Uo = ocore.projection_list[ci][b]
		print("   U rel blk", b, rel_fro(Ug.numpy(), Uo.numpy()), "ortho", float((Ug.double().transpose(1, 2) @ Ug.double() - torch.eye(Ug.shape[-1], dtype=torch.float64)).abs().max()))
