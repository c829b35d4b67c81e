"""CPU timing of the ingest stage on synthetic raw files: this package's
`ingest.pack_training_data_one_process` / `get_qc` (libfh_host.so) and the block-CSR staging, next to the
unmodified reference when /root/reference is present (this container only).

  python scripts/ingest_bench.py [--cells 2000] [--bins 499] [--contacts 1500] [--off-diag 100]
Prints one JSON line."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np
from scipy.sparse import coo_matrix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fasthigashi_b200  # noqa: E402,F401
import __graft_entry__ as ge  # noqa: E402
ge.build_host()
from fasthigashi_b200 import ingest  # noqa: E402
from fasthigashi_b200.sparse_for_schic import Sparse, Chrom_Dataset  # noqa: E402


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument("--cells", type=int, default=2000)
	ap.add_argument("--bins", type=int, default=499)
	ap.add_argument("--contacts", type=int, default=6000, help="contacts per cell before symmetrising")
	ap.add_argument("--off-diag", type=int, default=100)
	args = ap.parse_args()
	rng = np.random.default_rng(0)
	n = args.bins
	mats = np.empty(args.cells, dtype=object)
	for c in range(args.cells):
		i = rng.integers(0, n, size=args.contacts)
		j = np.clip(i + (rng.geometric(0.05, size=args.contacts) - 1) * rng.choice([-1, 1], size=args.contacts), 0, n - 1)
		m = coo_matrix((np.ones(len(i), np.float32), (i, j)), shape=(n, n)).tocsr()
		mats[c] = (m + m.T).tocsr()
	out = {"cells": args.cells, "bins": n, "nnz_per_cell": float(np.mean([m.nnz for m in mats])), "cores": os.cpu_count()}
	with tempfile.TemporaryDirectory() as tmp:
		os.makedirs(os.path.join(tmp, "raw"))
		np.save(os.path.join(tmp, "raw", "chr1_sparse_adj.npy"), mats, allow_pickle=True)
		reorder = np.arange(args.cells)
		t = time.perf_counter()
		kept, reads = ingest.get_qc(os.path.join(tmp, "raw"), ["chr1"], 500000)
		out["get_qc_s"] = time.perf_counter() - t
		t = time.perf_counter()
		idx, val, shape = ingest.pack_training_data_one_process(os.path.join(tmp, "raw"), "chr1", reorder, args.off_diag)
		out["pack_s"] = time.perf_counter() - t
		t = time.perf_counter()
		ds = Chrom_Dataset(Sparse(idx, val, shape, copy=False), bs_bin=125, bs_cell=args.cells, compact=True,
		                   flank=args.off_diag, chrom="chr1", resolution=500000, device="cpu")
		out["block_csr_cpu_s"] = time.perf_counter() - t
		out["nnz"] = int(len(val))
		try:
			from oracle import ref_shims
			mods = ref_shims.import_reference()
			import importlib
			W = importlib.import_module("fasthigashi.FastHigashi_Wrapper")
			fh = W.FastHigashi.__new__(W.FastHigashi)
			fh.config = {"chrom_list": ["chr1"], "temp_dir": tmp, "data_dir": tmp, "resolution": 500000, "resolution_fh": [500000]}
			fh.temp_dir, fh.chrom_list = tmp, ["chr1"]
			t = time.perf_counter()
			fh.get_qc()
			out["ref_get_qc_s"] = time.perf_counter() - t
			t = time.perf_counter()
			ridx, rval, rshape = fh.pack_training_data_one_process(
				raw_dir=os.path.join(tmp, "raw"), chrom="chr1", reorder=reorder, batch_norm=False, is_sym=True, off_diag=args.off_diag,
				fac_size=1, merge_fac_row=1, merge_fac_col=1, filename_pattern="%s_sparse_adj.npy", force_shift=False)
			out["ref_pack_s"] = time.perf_counter() - t
			sp = mods["sparse_for_schic"]
			t = time.perf_counter()
			obj = sp.Sparse(ridx, rval, np.asarray(rshape), copy=False)
			obj.sort_indices()
			sp.Chrom_Dataset(tensor=obj, bs_bin=125, bs_cell=args.cells, good_qc_num=-1, kind="hic", upper_sim=False, compact=True,
			                 flank=args.off_diag, chrom="chr1", resolution=500000)
			out["ref_chrom_dataset_s"] = time.perf_counter() - t
			out["same_nnz"] = bool(len(rval) == len(val))
		except ImportError:
			out["reference"] = "not present"
	print(json.dumps(out))


if __name__ == "__main__":
	main()
