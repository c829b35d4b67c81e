"""Golden fixtures for the ingest stage: runs the UNMODIFIED reference's
`FastHigashi.get_qc`, `pack_training_data_one_process` and `fetch_cell_embedding` /
`correct_batch_linear` (FastHigashi_Wrapper.py:221-366, 428-458, 750-878) on a small on-disk
raw dataset, in this container only.  Re-run:  python tests/golden/make_golden_ingest.py

Output tests/golden/ingest_cases.npz:
  raw_{chrom}_{indptr,indices,data}   per-cell CSR matrices of `raw/{chrom}_sparse_adj.npy`, concatenated
  qc, readcount                       reference get_qc()
  {case}_{chrom}_{idx,val,shape}      reference pack_training_data_one_process per case
  emb_*                               inputs and outputs of fetch_cell_embedding
"""
import os
import sys
import tempfile
import numpy as np
from scipy.sparse import csr_matrix

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

CHROMS = ["chr1", "chr2"]
BINS = {"chr1": 60, "chr2": 37}
NCELL = 30
RES = 500000


def make_raw(rng):
	"""Symmetric per-cell count matrices with empty bins, a few sparse (bad-QC) cells, fp32 data."""
	raw = {}
	for ch in CHROMS:
		n = BINS[ch]
		dead = rng.choice(n, size=3, replace=False)
		mats = []
		for c in range(NCELL):
			ncontact = rng.integers(150, 400) if c % 7 else rng.integers(3, 12)
			i = rng.integers(0, n, size=ncontact)
			d = np.minimum(rng.geometric(0.18, size=ncontact) - 1, n - 1)
			j = np.clip(i + d * rng.choice([-1, 1], size=ncontact), 0, n - 1)
			keep = ~(np.isin(i, dead) | np.isin(j, dead))
			i, j = i[keep], j[keep]
			m = np.zeros((n, n), dtype=np.float32)
			np.add.at(m, (i, j), 1.0)
			m = m + m.T - np.diag(np.diag(m))
			mats.append(csr_matrix(m.astype(np.float32)))
		raw[ch] = mats
	return raw


def write_raw(raw, temp_dir):
	os.makedirs(os.path.join(temp_dir, "raw"), exist_ok=True)
	for ch, mats in raw.items():
		arr = np.empty(len(mats), dtype=object)
		for i, m in enumerate(mats):
			arr[i] = m
		np.save(os.path.join(temp_dir, "raw", "%s_sparse_adj.npy" % ch), arr, allow_pickle=True)


def main():
	mods = ref_shims.import_reference()
	import importlib
	W = importlib.import_module("fasthigashi.FastHigashi_Wrapper")
	rng = np.random.default_rng(5)
	raw = make_raw(rng)
	out = {"chroms": np.array(CHROMS), "ncell": NCELL, "res": RES}
	for ch, mats in raw.items():
		out["raw_%s_indptr" % ch] = np.concatenate([m.indptr for m in mats]).astype(np.int32)
		out["raw_%s_indices" % ch] = np.concatenate([m.indices for m in mats]).astype(np.int32)
		out["raw_%s_data" % ch] = np.concatenate([m.data for m in mats]).astype(np.float32)
		out["raw_%s_n" % ch] = BINS[ch]
	batch = np.array(["b%d" % (c % 3) for c in range(NCELL)])
	out["batch"] = batch
	blacklist = {"chr1": np.array([5, 17, 200]), "chr2": np.array([2])}
	out["bl_chr1"], out["bl_chr2"] = blacklist["chr1"], blacklist["chr2"]

	with tempfile.TemporaryDirectory() as tmp:
		write_raw(raw, tmp)
		fh = W.FastHigashi.__new__(W.FastHigashi)
		fh.config = {"chrom_list": CHROMS, "temp_dir": tmp, "data_dir": tmp, "resolution": RES, "resolution_fh": [RES]}
		fh.temp_dir, fh.chrom_list = tmp, CHROMS
		qc, readcount = fh.get_qc()
		out["qc"], out["readcount"] = qc, readcount
		good, bad = np.where(qc > 0)[0], np.where(qc <= 0)[0]
		reorder = np.concatenate([np.sort(good), np.sort(bad)])
		out["reorder"] = reorder
		print("qc good/bad:", len(good), len(bad))

		def run(case, off_diag, merge, with_batch, batch_norm, with_bl):
			cfg = dict(fh.config)
			if with_batch:
				cfg["batch_id"] = "batch"
				fh.batch_id = batch[reorder]
			fh.config = cfg
			blp = os.path.join(tmp, "raw", "blacklist.npy")
			if with_bl:
				np.save(blp, blacklist, allow_pickle=True)
			elif os.path.exists(blp):
				os.remove(blp)
			for ch in CHROMS:
				idx, val, shape = fh.pack_training_data_one_process(
					raw_dir=os.path.join(tmp, "raw"), chrom=ch, reorder=reorder, batch_norm=batch_norm, is_sym=True,
					off_diag=off_diag, fac_size=1, merge_fac_row=merge, merge_fac_col=merge,
					filename_pattern="%s_sparse_adj.npy", force_shift=False)
				out["%s_%s_idx" % (case, ch)] = idx
				out["%s_%s_val" % (case, ch)] = val
				out["%s_%s_shape" % (case, ch)] = np.asarray(shape)
				print(case, ch, shape, idx.shape)
			fh.config = {k: v for k, v in cfg.items() if k != "batch_id"}

		run("plain", 12, 1, False, False, False)
		run("merge2", 8, 2, False, False, False)
		run("batch", 12, 1, True, True, False)
		run("batchoff", 12, 1, True, False, False)
		run("bl", 12, 1, False, False, True)

		# fetch_cell_embedding / correct_batch_linear on injected factors (host post-processing only)
		import pandas as pd
		rg = np.random.default_rng(9)
		R = 8
		fh.rank = R
		fh.meta_embedding = rg.standard_normal((NCELL, R))
		fh.D_list = [rg.standard_normal((R, 5)), rg.standard_normal((R, 7))]
		fh.coverage_feats = readcount[reorder].reshape(-1, 1)
		fh.reorder = reorder
		fh.label_info = pd.DataFrame({"batch": batch}).iloc[reorder].reset_index()
		fh.embedding_storage = None
		np.random.seed(0)
		store = fh.fetch_cell_embedding(final_dim=6, restore_order=True)
		np.random.seed(1)
		store = fh.correct_batch_linear("batch", add_intercept_back=True)
		out["emb_meta"], out["emb_D0"], out["emb_D1"] = fh.meta_embedding, fh.D_list[0], fh.D_list[1]
		for k, v in store.items():
			if k != "restore_order":
				out["emb_out_" + k] = np.asarray(v)
		out["emb_coverage_fh"] = np.asarray(fh.label_info["coverage_fh"])
	np.savez_compressed(os.path.join(HERE, "ingest_cases.npz"), **out)
	print("wrote ingest_cases.npz with", len(out), "arrays")


if __name__ == "__main__":
	main()
