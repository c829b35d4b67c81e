"""Host-side methods of the `FastHigashi` wrapper that never touch the device (meta/QC, raw ingest,
embedding post-processing), against reference fixtures (tests/golden/make_golden_ingest.py).
The constructor itself demands a CUDA device, so the object is assembled with __new__ here."""
import os
import numpy as np
import pandas as pd
import pytest
from conftest import GOLDEN
from fasthigashi_b200.FastHigashi_Wrapper import FastHigashi
from test_ingest_golden import G, CHROMS, NCELL, RES, raw_dir  # noqa: F401  (fixture)


def host_only_wrapper(tmp, cache, with_batch=False):
	fh = FastHigashi.__new__(FastHigashi)
	fh.config = {"chrom_list": CHROMS, "temp_dir": tmp, "data_dir": tmp, "resolution": RES, "resolution_fh": [RES]}
	if with_batch:
		fh.config["batch_id"] = "batch"
	fh.chrom_list, fh.temp_dir, fh.data_dir = CHROMS, tmp, tmp
	fh.path2input_cache = fh.path2result_dir = cache
	fh.off_diag, fh.filter, fh.fh_resolutions = 12, True, [RES]
	fh._tensors, fh._meta, fh._batch_norm = None, None, True
	fh.embedding_storage = None
	return fh


def test_constructor_refuses_cpu(tmp_path):
	import torch
	if torch.cuda.is_available():
		pytest.skip("CUDA present")
	with pytest.raises(RuntimeError, match="CUDA"):
		FastHigashi({"chrom_list": CHROMS, "temp_dir": str(tmp_path), "resolution": RES, "resolution_fh": [RES]},
		            None, None, 12, True, True, True, False, False)


def test_preprocess_meta_and_raw_ingest(raw_dir, tmp_path):  # noqa: F811
	import pickle
	pickle.dump({"batch": list(G["batch"])}, open(os.path.join(raw_dir, "label_info.pickle"), "wb"))
	fh = host_only_wrapper(raw_dir, str(tmp_path), with_batch=True)
	label_info, reorder, readcount, qc = fh.preprocess_meta()
	assert np.array_equal(reorder, G["reorder"]) and np.array_equal(qc, G["qc"])
	assert np.array_equal(fh.batch_id, G["batch"][reorder])
	for f in ("qc.npy", "read_count_all.npy", "reorder.npy"):  # the reference writes these too (:205-209)
		assert os.path.exists(os.path.join(str(tmp_path), f))
	# second call takes the cached QC
	assert np.array_equal(fh.preprocess_meta()[3], qc)
	mats = fh._load_tensors(RES, reorder)
	assert os.path.exists(os.path.join(str(tmp_path), "cache_intra_%d_offdiag_12_b200.pkl" % RES))
	for ch, m in zip(CHROMS, mats):
		assert tuple(int(x) for x in m.shape) == tuple(int(x) for x in G["batch_%s_shape" % ch])
		assert len(m.values) == len(G["batch_%s_val" % ch])
		np.testing.assert_allclose(np.sort(m.values.numpy()), np.sort(G["batch_%s_val" % ch]), rtol=2e-6, atol=1e-7)
	# method with the reference signature
	idx, val, shape = fh.pack_training_data_one_process(os.path.join(raw_dir, "raw"), "chr1", reorder, off_diag=12, fac_size=1,
	                                                    merge_fac_row=1, merge_fac_col=1, is_sym=True, force_shift=False)
	assert idx.shape[1] == len(G["batch_chr1_val"])
	with pytest.raises(NotImplementedError):
		fh.pack_training_data_one_process(os.path.join(raw_dir, "raw"), "chr1", reorder, fac_size=2)


def test_fetch_cell_embedding_matches_reference(tmp_path):
	fh = host_only_wrapper(str(tmp_path), str(tmp_path))
	reorder = G["reorder"]
	fh.rank = 8
	fh.meta_embedding = G["emb_meta"]
	fh.D_list = [G["emb_D0"], G["emb_D1"]]
	fh.coverage_feats = G["readcount"][reorder].reshape(-1, 1)
	fh.reorder = reorder
	fh.label_info = pd.DataFrame({"batch": G["batch"]}).iloc[reorder].reset_index()
	np.random.seed(0)
	store = fh.fetch_cell_embedding(final_dim=6, restore_order=True)
	np.random.seed(1)
	store = fh.correct_batch_linear("batch", add_intercept_back=True)
	assert set(store) == {"embed_all", "embed_raw", "embed_l2_norm", "restore_order", "embed_correct_coverage_fh",
	                      "embed_l2_norm_correct_coverage_fh", "embed_correct_batch", "embed_l2_norm_correct_batch"}
	np.testing.assert_allclose(store["embed_all"], G["emb_out_embed_all"], rtol=1e-10, atol=1e-12)
	np.testing.assert_allclose(np.asarray(fh.label_info["coverage_fh"]), G["emb_coverage_fh"], rtol=1e-10)
	for k in ("embed_raw", "embed_l2_norm", "embed_correct_coverage_fh", "embed_correct_batch", "embed_l2_norm_correct_batch"):
		a, b = store[k], G["emb_out_" + k]
		assert a.shape == b.shape
		# randomized SVD: same seed gives the same basis; allow a per-component sign and a loose tolerance
		sgn = np.sign(np.sum(a * b, axis=0))
		np.testing.assert_allclose(a * sgn, b, rtol=1e-5, atol=1e-7)
	assert fh.correct_batch_linear("missing") is None


def test_only_partial_rwr_assembly_logic(tmp_path, monkeypatch):
	"""only_partial_rwr (FastHigashi_Wrapper.py:569-655) with the device call replaced by the CPU oracle: checks what the
	wrapper itself does - one auto-stopped call per (bin-block, cell batch), pasting the block windows into the (n, n) map,
	m + m^T with the diagonal halved, datasets keyed by the ORIGINAL cell id, good and bad cells - against the same steps
	written as in the reference (float64 maps, then cast). The device call itself is covered by the GPU tests."""
	import torch
	from conftest import load_small_dataset
	from oracle import fh_oracle as O
	import fasthigashi_b200.partial_rwr as prw

	calls = []

	def fake_rwr_block_csr(ds, b, cell0, ncell, out, out_cell_stride, k, do_conv, do_rwr, do_col, bin_cov=None, use_tc=False, chunk=None):
		g = ds.geoms[b]
		x, n_it = O.partial_rwr(O.densify_block(ds, b, cell0, cell0 + ncell), g.s, g.e, do_conv, do_rwr, do_col, None, k)
		ldw = prw.pad4(g.w)
		view = out.view(ncell, g.nb, ldw)
		view[:, :, :g.w] = x
		calls.append((ds.chrom, b, cell0, ncell, k, do_col))
		return n_it
	monkeypatch.setattr(prw, "rwr_block_csr", fake_rwr_block_csr)
	fh = host_only_wrapper(str(tmp_path), str(tmp_path))
	fh.device = "cpu"
	fh.do_conv, fh.do_rwr = True, True
	fh.all_matrix = load_small_dataset(good_qc_num=44, bs_cell=20)
	rng = np.random.default_rng(0)
	fh.reorder = rng.permutation(48)
	files = fh.only_partial_rwr(out_format="npz")
	assert [os.path.basename(f) for f in files] == ["impute_prwr_chr1.npz", "impute_prwr_chr2.npz", "impute_prwr_chr3.npz"]
	assert all(k == -1 and not col for (_, _, _, _, k, col) in calls)          # auto-stop, do_col=False (:606-607)
	for ds, path in zip(fh.all_matrix, files):
		got = np.load(path)
		n = ds.num_bin
		assert list(got["shape"]) == [n, n] and len(got.files) == 49
		assert sorted(int(k) for k in got.files if k != "shape") == list(range(48))
		assert [(c0, nc) for (ch, b, c0, nc, _, _) in calls if ch == ds.chrom and b == 0] == [(0, 20), (20, 20), (40, 4), (44, 4)]
		for sl in ds.cell_slice_list:
			nc = sl.stop - sl.start
			full = np.zeros((nc, n, n))
			for b, g in enumerate(ds.geoms):
				x, _ = O.partial_rwr(O.densify_block(ds, b, sl.start, sl.stop), g.s, g.e, True, True, False, None, -1)
				full[:, g.row0:g.row0 + g.nb, g.col0:g.col0 + g.w] = x.numpy()
			full = full + full.transpose(0, 2, 1)
			for i in range(nc):
				m = (full[i] - np.diag(np.diag(full[i]) / 2)).astype("float32")
				assert np.array_equal(got[str(fh.reorder[sl.start + i])], m)        # fp32 on the device == float64-then-cast
	with pytest.raises(ValueError):
		fh.only_partial_rwr(out_format="csv")


def test_fetch_cell_embedding_device_svd_option(tmp_path):
	"""svd="device" (dist_svd randomized SVD instead of sklearn's): same subspace and singular directions as the reference
	output up to sign for the well-separated leading components; here it runs on the CPU device."""
	fh = host_only_wrapper(str(tmp_path), str(tmp_path))
	fh.device = "cpu"
	reorder = G["reorder"]
	fh.rank = 8
	fh.meta_embedding = G["emb_meta"]
	fh.D_list = [G["emb_D0"], G["emb_D1"]]
	fh.coverage_feats = G["readcount"][reorder].reshape(-1, 1)
	fh.reorder = reorder
	fh.label_info = pd.DataFrame({"batch": G["batch"]}).iloc[reorder].reset_index()
	np.random.seed(0)
	store = fh.fetch_cell_embedding(final_dim=6, restore_order=True, svd="device")
	np.testing.assert_allclose(store["embed_all"], G["emb_out_embed_all"], rtol=1e-10, atol=1e-12)
	a, b = store["embed_raw"], G["emb_out_embed_raw"]
	assert a.shape == b.shape
	# rank-8 input, 6 components, 5 power iterations: both randomized SVDs have converged to the exact one
	sgn = np.sign(np.sum(a * b, axis=0))
	np.testing.assert_allclose(a * sgn, b, rtol=1e-6, atol=1e-8)
	assert "embed_correct_coverage_fh" in store and store["embed_correct_coverage_fh"].shape == b.shape
	with pytest.raises(ValueError):
		fh.fetch_cell_embedding(final_dim=6, svd="gpu")


def test_preprocess_contact_map_method(raw_dir, tmp_path):  # noqa: F811
	"""FastHigashi.preprocess_contact_map with the reference's call (FastHigashi_Wrapper.py:484-494): one dim-0-sorted
	`Sparse` per chromosome, cached; coarsening through merge_fac_row/col (the reference's 'merge2' fixture)."""
	fh = host_only_wrapper(raw_dir, str(tmp_path))
	reorder = G["reorder"]
	cache = os.path.join(str(tmp_path), "cache_intra_x.pkl")
	mats = fh.preprocess_contact_map(fh.config, reorder=reorder, path2input_cache=cache, batch_norm=False, is_sym=True, off_diag=12,
	                                 fac_size=1, merge_fac_row=1, merge_fac_col=1, filename_pattern="%s_sparse_adj.npy", force_shift=False)
	assert os.path.exists(cache) and len(mats) == len(CHROMS)
	for ch, m in zip(CHROMS, mats):
		assert tuple(int(x) for x in m.shape) == tuple(int(x) for x in G["plain_%s_shape" % ch])
		assert m.indptr is not None and bool((m.indices[0][1:] >= m.indices[0][:-1]).all())
		ref = np.zeros(tuple(int(x) for x in m.shape), np.float32)
		ri = G["plain_%s_idx" % ch]
		ref[ri[0], ri[1], ri[2]] = G["plain_%s_val" % ch]
		np.testing.assert_allclose(m.to_dense().numpy(), ref, rtol=2e-6, atol=1e-7)
	again = fh.preprocess_contact_map(fh.config, reorder=reorder, path2input_cache=cache, batch_norm=False, off_diag=12)
	assert all(np.array_equal(a.values.numpy(), b.values.numpy()) for a, b in zip(mats, again))
	m2 = fh.preprocess_contact_map(fh.config, reorder=reorder, path2input_cache=None, batch_norm=False, off_diag=8, merge_fac_row=2, merge_fac_col=2)
	for ch, m in zip(CHROMS, m2):
		assert tuple(int(x) for x in m.shape) == tuple(int(x) for x in G["merge2_%s_shape" % ch])
	with pytest.raises(NotImplementedError):
		fh.preprocess_contact_map(fh.config, reorder=reorder, path2input_cache=None, batch_norm=False, fac_size=2)
