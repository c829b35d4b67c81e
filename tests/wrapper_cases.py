"""Wrapper-level cases shared by the GPU tests (tests/test_zz_gpu_round2_candidates.py: the real library on a B200) and
the CPU orchestration tests (tests/test_orchestration_emulated.py: the host-memory stand-in of tests/emu/fake_abi.py).
`make_wrapper(tmp_path, off_diag, resolution, chroms)` returns a `FastHigashi` object bound to the device under test."""
import os
import numpy as np
import torch
from conftest import GOLDEN
from oracle import fh_oracle as O


def case_only_partial_rwr_matches_oracle(make_wrapper, tmp_path):
	"""FastHigashi_Wrapper.py:569-655: per-cell imputed maps, symmetrised, keyed by the original cell id."""
	d = np.load(os.path.join(GOLDEN, "data_small.npz"))
	ncell, off, res = int(d["ncell"]), int(d["off_diag"]), int(d["res"])
	chroms = ["chr1", "chr2", "chr3"]
	tensors = {res: [(d[c + "_idx"].astype(np.int64), d[c + "_val"], (int(n), int(n), ncell)) for c, n in zip(chroms, d["bins"])]}
	qc = np.ones(ncell); qc[[5, 17]] = 0
	w = make_wrapper(tmp_path, off, res, chroms)
	w.set_tensors(tensors, qc=qc, readcount=np.linspace(8, 10, ncell))
	w.prep_dataset()
	files = w.only_partial_rwr(out_format="npz")
	assert len(files) == 3
	for ds, path in zip(w.all_matrix, files):
		got = np.load(path)
		assert list(got["shape"]) == [ds.num_bin, ds.num_bin]
		cpu = ds.select_cells(0, ds.total_cell_num, good_qc_num=ds.num_cell).to("cpu")
		n = ds.num_bin
		seen = 0
		for sl in ds.cell_slice_list:
			nc = sl.stop - sl.start
			full = np.zeros((nc, n, n))
			for b, g in enumerate(cpu.geoms):
				x, _ = O.partial_rwr(O.densify_block(cpu, b, sl.start, sl.stop), g.s, g.e, True, True, False, None, -1)
				full[:, g.row0:g.row0 + g.nb, g.col0:g.col0 + g.w] = x.numpy()
			full = full + full.transpose(0, 2, 1)
			for i in range(nc):
				m = full[i] - np.diag(np.diag(full[i]) / 2)
				a = got[str(w.reorder[sl.start + i])]
				assert a.dtype == np.float32 and a.shape == (n, n)
				assert np.linalg.norm(a - m) <= 1e-5 * np.linalg.norm(m)
				seen += 1
		assert seen == ncell and len(got.files) == ncell + 1


def write_raw_files(G, tmp_path):
	"""raw/{chrom}_sparse_adj.npy (object arrays of per-cell scipy CSR) from tests/golden/ingest_cases.npz."""
	from scipy.sparse import csr_matrix
	chroms = [str(c) for c in G["chroms"]]
	ncell = int(G["ncell"])
	os.makedirs(os.path.join(str(tmp_path), "raw"))
	for ch in chroms:
		n = int(G["raw_%s_n" % ch])
		indptr = G["raw_%s_indptr" % ch].reshape(ncell, n + 1)
		arr, offp = np.empty(ncell, dtype=object), 0
		for c in range(ncell):
			nnz = int(indptr[c, -1])
			arr[c] = csr_matrix((G["raw_%s_data" % ch][offp:offp + nnz], G["raw_%s_indices" % ch][offp:offp + nnz], indptr[c]), shape=(n, n))
			offp += nnz
		np.save(os.path.join(str(tmp_path), "raw", "%s_sparse_adj.npy" % ch), arr, allow_pickle=True)


def case_wrapper_from_raw_files(make_wrapper, tmp_path):
	"""prep_dataset straight from raw/{chrom}_sparse_adj.npy (ingest.py) -> run_model -> embeddings, against the
	oracle fed with the reference's own packed tensors of the same raw files (tests/golden/ingest_cases.npz)."""
	from scipy.sparse import csr_matrix
	from fasthigashi_b200.sparse_for_schic import Sparse, Chrom_Dataset
	G = np.load(os.path.join(GOLDEN, "ingest_cases.npz"), allow_pickle=True)
	chroms = [str(c) for c in G["chroms"]]
	ncell, res = int(G["ncell"]), int(G["res"])
	write_raw_files(G, tmp_path)
	w = make_wrapper(tmp_path, 12, res, chroms)
	w.prep_dataset()
	assert np.array_equal(w.reorder, G["reorder"]) and w.good_qc_num == int(G["qc"].sum())
	torch.manual_seed(0); np.random.seed(0)
	w.run_model(dim1=0.6, rank=8, n_iter_parafac=1, n_iter_max=4, tol=0.0)
	emb = w.fetch_cell_embedding(final_dim=4)
	ods = []
	for ds, ch in zip(w.all_matrix, chroms):
		sp = Sparse(G["plain_%s_idx" % ch].astype(np.int64), G["plain_%s_val" % ch], tuple(int(x) for x in G["plain_%s_shape" % ch]))
		ods.append(Chrom_Dataset(sp, bs_bin=ds.bs_bin, bs_cell=ds.bs_cell, good_qc_num=ds.num_cell, compact=True, flank=12,
		                         chrom=ch, resolution=res))
	oc = O.OracleCore(8, 12, [res])
	torch.manual_seed(0); np.random.seed(0)
	oc.fit(ods, 0.6, 4, 1, True, True, w.final_do_col, 0.0)
	Vo = oc.transform(ods, True, True, w.final_do_col)
	Eo = O.embed_all(Vo.numpy(), [x.numpy() for x in oc.D_dict.values()])
	pear = [abs(np.corrcoef(emb["embed_all"][:, j], Eo[:, j])[0, 1]) for j in range(Eo.shape[1])]
	assert min(pear) > 0.999, min(pear)
