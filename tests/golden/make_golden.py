"""Generates the committed golden fixtures by running the UNMODIFIED reference from
/root/reference on CPU (this container only). Re-run:  python tests/golden/make_golden.py

Outputs (tests/golden/):
  data_small.npz   the synthetic COO tensors the fixtures were computed on
  rwr_cases.npz    reference densify (`Chrom_Dataset.fetch`) and `partial_rwr` outputs
  polar_cases.npz  reference `project2orthogonal` outputs
  cp_cases.npz     reference `parafac` outputs
  core_*.npz       reference `Fast_Higashi_core.fit_transform`: init state, per-sweep loss terms,
                   final factors / embeddings
The reference's own loop runs untouched; values are observed by wrapping
`update_meta_embedding_interactions` (entry snapshot of the factors + its return values).
"""
import os
import sys
import io
import contextlib
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import fasthigashi_b200  # noqa: E402
from fasthigashi_b200 import synth  # noqa: E402
from oracle import ref_shims  # noqa: E402

BINS = [90, 70, 40]
NCELL = 48
OFF_DIAG = 12
RES = 1000000
BS_BIN = 32
RANK = 16


def to_np(x):
	return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


def main():
	torch.set_num_threads(4)
	mods = ref_shims.import_reference()
	chroms, cluster = synth.synth_dataset(BINS, NCELL, 0.12, off_diag=OFF_DIAG, seed=3, num_cluster=4)
	data = {"bins": np.array(BINS), "ncell": NCELL, "off_diag": OFF_DIAG, "res": RES, "cluster": cluster}
	for ch in chroms:
		data[ch["chrom"] + "_idx"] = to_np(ch["indices"]).astype(np.int16)
		data[ch["chrom"] + "_val"] = to_np(ch["values"]).astype(np.float32)
	np.savez_compressed(os.path.join(HERE, "data_small.npz"), **data)

	prwr = mods["partial_rwr"].partial_rwr
	# ---------------- RWR stage fixtures ----------------
	ds_list = ref_shims.build_reference_datasets(mods, chroms, off_diag=OFF_DIAG, res=RES, bs_bin=BS_BIN,
	                                             bs_cell=20, good_qc_num=44)
	out = {}
	case = 0
	gen = torch.Generator().manual_seed(11)
	for ci, b, cb in [(0, 0, 0), (0, 1, 1), (0, 2, 0), (1, 1, 2), (2, 1, 0), (2, 0, 1)]:
		ds = ds_list[ci]
		(x, _), kind = ds.fetch(b, cb, save_context=dict(device="cpu"), transpose=True, do_conv=False)
		sl = ds.local_bin_slice_list[b]
		cov = torch.rand(x.shape[0], x.shape[2], generator=gen) + 0.5
		cov[0, :3] = float("inf")
		out["c%d_meta" % case] = np.array([ci, b, cb, sl.start, sl.stop])
		out["c%d_dense" % case] = to_np(x)
		out["c%d_cov" % case] = to_np(cov)
		for name, kw in [
			("conv_rwr_k3", dict(do_conv=True, do_rwr=True, do_col=False, force_rwr_epochs=3)),
			("conv_rwr_k0", dict(do_conv=True, do_rwr=True, do_col=False, force_rwr_epochs=0)),
			("conv_rwr_k1", dict(do_conv=True, do_rwr=True, do_col=False, force_rwr_epochs=1)),
			("conv_rwr_col_k4", dict(do_conv=True, do_rwr=True, do_col=True, force_rwr_epochs=4)),
			("rwr_only_k2", dict(do_conv=False, do_rwr=True, do_col=False, force_rwr_epochs=2)),
			("conv_only", dict(do_conv=True, do_rwr=False, do_col=False)),
			("auto", dict(do_conv=True, do_rwr=True, do_col=False, force_rwr_epochs=-1)),
			("auto_col", dict(do_conv=True, do_rwr=True, do_col=True, force_rwr_epochs=-1)),
		]:
			y, n_it = prwr(x.clone(), slice_start=sl.start, slice_end=sl.stop, bin_cov=cov.clone(),
			               return_rwr_iter=True, final_transpose=False, **kw)
			out["c%d_%s" % (case, name)] = to_np(y).astype(np.float32)
			out["c%d_%s_niter" % (case, name)] = np.array(n_it)
		case += 1
	out["ncase"] = np.array(case)
	np.savez_compressed(os.path.join(HERE, "rwr_cases.npz"), **out)

	# ---------------- polar fixtures ----------------
	p2o = mods["project2orthogonal"].project2orthogonal
	g = torch.Generator().manual_seed(5)
	out = {}
	mats = [torch.randn(5, 40, 12, generator=g),
	        torch.randn(3, 12, 12, generator=g),
	        torch.randn(2, 9, 14, generator=g),  # wide: rows < cols
	        torch.randn(1, 200, 16, generator=g)]
	# ill-conditioned tall batch (kappa ~ 1e6, as H3)
	Uq, _ = torch.linalg.qr(torch.randn(4, 60, 20, generator=g))
	Vq, _ = torch.linalg.qr(torch.randn(4, 20, 20, generator=g))
	sv = torch.logspace(0, -6, 20)[None].repeat(4, 1)
	mats.append(Uq * sv[:, None, :] @ Vq.transpose(1, 2))
	for i, m in enumerate(mats):
		with contextlib.redirect_stdout(io.StringIO()):
			U, S = p2o(m.clone(), m.shape[-1], compute_device=torch.device("cpu"))
		out["m%d" % i], out["u%d" % i], out["s%d" % i] = to_np(m), to_np(U), to_np(S)
	out["n"] = np.array(len(mats))
	np.savez_compressed(os.path.join(HERE, "polar_cases.npz"), **out)

	# ---------------- inner CP-ALS fixtures ----------------
	parafac = mods["parafac_integrative"].parafac
	out = {}
	g = torch.Generator().manual_seed(9)
	for i, (n, r, R, iters) in enumerate([(50, 8, 16, 1), (33, 12, 16, 3), (20, 5, 7, 10)]):
		A0 = torch.randn(n, r, generator=g) * 1e-2 + 1
		B0 = torch.eye(r) + 0.01 * torch.randn(r, generator=g)
		D0 = torch.randn(R, r, generator=g)
		Y = torch.einsum("ir,jr,kr->ijk", torch.randn(n, r, generator=g), torch.randn(r, r, generator=g),
		                 torch.randn(R, r, generator=g)) + 0.1 * torch.randn(n, r, R, generator=g)
		fac, cn, lx = parafac(Y.clone(), rank=r, init=[A0.clone(), B0.clone(), D0.clone()], n_iter_max=iters)
		for nm, v in zip(["Y", "A0", "B0", "D0", "A1", "B1", "D1"], [Y, A0, B0, D0] + list(fac)):
			out["p%d_%s" % (i, nm)] = to_np(v)
		out["p%d_scal" % i] = np.array([iters, cn, lx])
	out["n"] = np.array(3)
	np.savez_compressed(os.path.join(HERE, "cp_cases.npz"), **out)

	# ---------------- full core runs ----------------
	for tag, kw in [
		("col", dict(do_col=True, good_qc_num=44, bs_cell=20, n_iter_max=12)),
		("nocol", dict(do_col=False, good_qc_num=-1, bs_cell=NCELL, n_iter_max=6)),
	]:
		ds_list = ref_shims.build_reference_datasets(mods, chroms, off_diag=OFF_DIAG, res=RES, bs_bin=BS_BIN,
		                                             bs_cell=kw["bs_cell"], good_qc_num=kw["good_qc_num"])
		core = mods["parafac2_intergrative"].Fast_Higashi_core(rank=RANK, off_diag=OFF_DIAG, res_list=[RES]).to("cpu")
		trace = []
		orig = core.update_meta_embedding_interactions

		def wrapped(*a, _orig=orig, _core=core, _trace=trace, **k):
			snap = dict(A=[to_np(x).copy() for x in _core.A_list],
			            B=[to_np(x).copy() for x in _core.B_dict.values()],
			            D=[to_np(x).copy() for x in _core.D_dict.values()],
			            V=to_np(_core.meta_embedding).copy())
			res = _orig(*a, **k)
			snap["x_U"] = np.array(res[2]).ravel().copy()
			snap["x_V"] = float(res[3])
			if len(res) == 5:
				snap["xnorm"] = np.array(res[4]).ravel().copy()
			snap["Y"] = [to_np(v).copy() for v in res[1].values()]
			snap["V_new"] = to_np(_core.meta_embedding).copy()
			_trace.append(snap)
			return res
		core.update_meta_embedding_interactions = wrapped
		torch.manual_seed(0); np.random.seed(0)
		with contextlib.redirect_stdout(io.StringIO()) as buf:
			res = core.fit_transform(ds_list, size_ratio=0.3, n_iter_max=kw["n_iter_max"], n_iter_parafac=1,
			                         do_conv=True, do_rwr=True, do_col=kw["do_col"], tol=0.0, gpu_id=None,
			                         run_init=True)
		printed = [float(l.split("re=")[1].split()[0]) for l in buf.getvalue().splitlines() if "PARAFAC2 re=" in l]
		A_list, B_list, D_list, V_final = res[1]
		out = dict(n_i=np.asarray(core.n_i), nsweep=np.array(len(trace)), printed_re=np.array(printed),
		           do_col=np.array(kw["do_col"]), good_qc_num=np.array(ds_list[0].num_cell),
		           bs_cell=np.array(kw["bs_cell"]), rank=np.array(RANK), bs_bin=np.array(BS_BIN))
		xnorm = trace[0]["xnorm"]
		re = []
		for t, s in enumerate(trace):
			core_n = np.array([float(torch.einsum("ir,jr,kr->kij", torch.from_numpy(a), torch.from_numpy(b),
			                                      torch.from_numpy(d)).square().sum())
			                   for a, b, d in zip(s["A"], s["B"], s["D"])])
			re.append(np.sqrt(xnorm.sum() + core_n.sum() - 2 * s["x_V"]) / np.sqrt(xnorm.sum()))
			out["t%d_x_U" % t] = s["x_U"]; out["t%d_x_V" % t] = np.array(s["x_V"])
			if t in (0, 1, len(trace) - 1):
				for i in range(len(BINS)):
					out["t%d_A%d" % (t, i)] = s["A"][i]; out["t%d_B%d" % (t, i)] = s["B"][i]
					out["t%d_D%d" % (t, i)] = s["D"][i]; out["t%d_Y%d" % (t, i)] = s["Y"][i]
				out["t%d_V" % t] = s["V"]; out["t%d_V_new" % t] = s["V_new"]
		out["re"] = np.array(re); out["xnorm"] = xnorm
		assert np.allclose(np.round(re, 3), printed, atol=1.1e-3), (re, printed)
		for i in range(len(BINS)):
			out["final_A%d" % i] = to_np(A_list[i]); out["final_B%d" % i] = to_np(list(B_list)[i])
			out["final_D%d" % i] = to_np(list(D_list)[i])
			out["bin_cov%d" % i] = to_np(core.bin_cov_list[i])
			bb = core.bad_bin_cov_list[i]
			out["bad_bin_cov%d" % i] = to_np(bb) if torch.is_tensor(bb) else np.zeros((0, BINS[i]), np.float32)
			for b, U in enumerate(core.projection_list[i]):
				out["final_U%d_%d" % (i, b)] = to_np(U)
		out["final_V"] = to_np(V_final)
		np.savez_compressed(os.path.join(HERE, "core_%s.npz" % tag), **out)
		print(tag, "n_i", core.n_i, "re", np.round(re, 5))


if __name__ == "__main__":
	main()
