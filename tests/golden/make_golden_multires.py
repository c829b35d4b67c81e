"""Golden fixture for the multi-resolution path of `Fast_Higashi_core.fit_transform` (two resolutions of the same
chromosomes: shared B and D per chromosome, bins stacked along mode 0 of the projected tensor,
parafac2_intergrative.py:581-592, 670-695), produced by the UNMODIFIED reference on CPU in the build container.
Re-run:  python tests/golden/make_golden_multires.py   ->  tests/golden/core_multires.npz
The reference's loop runs untouched; values are observed by wrapping `update_meta_embedding_interactions`."""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import fasthigashi_b200  # noqa: E402,F401
from fasthigashi_b200 import synth  # noqa: E402
from oracle import ref_shims  # noqa: E402

NCELL, OFF_DIAG, RANK, NSWEEP = 40, 12, 12, 5
# (resolution, bins per chromosome, bs_bin, density, seed): 1 Mb first, then 500 kb (resolution-major order, the wrapper's)
LEVELS = [(1000000, [70, 44], 32, 0.12, 21), (500000, [140, 88], 48, 0.06, 22)]


def to_np(x):
	return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


def main():
	torch.set_num_threads(4)
	mods = ref_shims.import_reference()
	out = dict(ncell=np.array(NCELL), off_diag=np.array(OFF_DIAG), rank=np.array(RANK), nlevel=np.array(len(LEVELS)),
	           res=np.array([l[0] for l in LEVELS]), bs_bin=np.array([l[2] for l in LEVELS]))
	ds_list = []
	for li, (res, bins, bs_bin, density, seed) in enumerate(LEVELS):
		chroms, _ = synth.synth_dataset(bins, NCELL, density, off_diag=OFF_DIAG, seed=seed, num_cluster=4)
		out["bins%d" % li] = np.array(bins)
		for ch in chroms:
			out["l%d_%s_idx" % (li, ch["chrom"])] = to_np(ch["indices"]).astype(np.int16)
			out["l%d_%s_val" % (li, ch["chrom"])] = to_np(ch["values"]).astype(np.float32)
		ds_list += ref_shims.build_reference_datasets(mods, chroms, off_diag=OFF_DIAG, res=res, bs_bin=bs_bin, bs_cell=NCELL)
	core = mods["parafac2_intergrative"].Fast_Higashi_core(rank=RANK, off_diag=OFF_DIAG, res_list=[l[0] for l in LEVELS]).to("cpu")
	trace = []
	orig = core.update_meta_embedding_interactions

	def wrapped(*a, **k):
		snap = dict(A=[to_np(x).copy() for x in core.A_list], B=[to_np(x).copy() for x in core.B_dict.values()],
		            D=[to_np(x).copy() for x in core.D_dict.values()], V=to_np(core.meta_embedding).copy())
		res = orig(*a, **k)
		snap["x_U"] = np.array(res[2]).ravel().copy()
		snap["x_V"] = float(res[3])
		if len(res) == 5:
			snap["xnorm"] = np.array(res[4]).ravel().copy()
		snap["Y"] = [to_np(v).copy() for v in res[1].values()]
		trace.append(snap)
		return res
	core.update_meta_embedding_interactions = wrapped
	torch.manual_seed(0); np.random.seed(0)
	with contextlib.redirect_stdout(io.StringIO()) as buf:
		res = core.fit_transform(ds_list, size_ratio=0.3, n_iter_max=NSWEEP, n_iter_parafac=1, do_conv=True, do_rwr=True,
		                         do_col=False, tol=0.0, gpu_id=None, run_init=True)
	printed = [float(l.split("re=")[1].split()[0]) for l in buf.getvalue().splitlines() if "PARAFAC2 re=" in l]
	A_list, B_list, D_list, V_final = res[1]
	xnorm = trace[0]["xnorm"]
	re = []
	nds, nchrom = len(ds_list), len(LEVELS[0][1])
	for t, s in enumerate(trace):
		core_n = np.array([float(torch.einsum("ir,jr,kr->kij", torch.from_numpy(s["A"][i]), torch.from_numpy(s["B"][i % nchrom]),
		                                      torch.from_numpy(s["D"][i % nchrom])).square().sum()) for i in range(nds)])
		re.append(np.sqrt(xnorm.sum() + core_n.sum() - 2 * s["x_V"]) / np.sqrt(xnorm.sum()))
		out["t%d_x_U" % t] = s["x_U"]; out["t%d_x_V" % t] = np.array(s["x_V"])
		if t in (0, 1):
			for i in range(nds):
				out["t%d_A%d" % (t, i)] = s["A"][i]
			for c in range(nchrom):
				out["t%d_B%d" % (t, c)] = s["B"][c]; out["t%d_D%d" % (t, c)] = s["D"][c]; out["t%d_Y%d" % (t, c)] = s["Y"][c]
			out["t%d_V" % t] = s["V"]
	assert np.allclose(np.round(re, 3), printed, atol=1.1e-3), (re, printed)
	out.update(re=np.array(re), xnorm=xnorm, n_i=np.asarray(core.n_i), nsweep=np.array(len(trace)), final_V=to_np(V_final),
	           chrom2size=np.array(list(core.chrom2size.values())))
	for i in range(nds):
		out["final_A%d" % i] = to_np(A_list[i])
		out["bin_cov%d" % i] = to_np(core.bin_cov_list[i])
	for c in range(nchrom):
		out["final_B%d" % c] = to_np(list(B_list)[c]); out["final_D%d" % c] = to_np(list(D_list)[c])
	np.savez_compressed(os.path.join(HERE, "core_multires.npz"), **out)
	print("n_i", core.n_i, "sizes", core.chrom2size, "re", np.round(re, 5))


if __name__ == "__main__":
	main()
