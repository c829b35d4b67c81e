import os
import sys
import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
	sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
	config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
	if torch.cuda.is_available():
		return
	skip = pytest.mark.skip(reason="no CUDA device")
	for item in items:
		if "gpu" in item.keywords:
			item.add_marker(skip)


def load_small_dataset(good_qc_num=-1, bs_cell=None, bs_bin=None, device="cpu"):
	"""The committed synthetic tensors of tests/golden/data_small.npz as block-CSR datasets."""
	import fasthigashi_b200  # noqa: F401
	from fasthigashi_b200.sparse_for_schic import Sparse, Chrom_Dataset
	d = np.load(os.path.join(GOLDEN, "data_small.npz"))
	ncell, off, res = int(d["ncell"]), int(d["off_diag"]), int(d["res"])
	out = []
	for i, n in enumerate(d["bins"]):
		ch = "chr%d" % (i + 1)
		sp = Sparse(d[ch + "_idx"].astype(np.int64), d[ch + "_val"], (int(n), int(n), ncell))
		out.append(Chrom_Dataset(sp, bs_bin=bs_bin or 32, bs_cell=bs_cell or ncell, good_qc_num=good_qc_num,
		                         compact=True, flank=off, chrom=ch, resolution=res, device=device))
	return out


def rel_fro(a, b):
	a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
	return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def load_multires_dataset(device="cpu"):
	"""tests/golden/core_multires.npz: the same chromosomes at two resolutions, resolution-major (the wrapper's order)."""
	import fasthigashi_b200  # noqa: F401
	from fasthigashi_b200.sparse_for_schic import Sparse, Chrom_Dataset
	g = np.load(os.path.join(GOLDEN, "core_multires.npz"))
	ncell, off = int(g["ncell"]), int(g["off_diag"])
	out = []
	for li in range(int(g["nlevel"])):
		for i, n in enumerate(g["bins%d" % li]):
			ch = "chr%d" % (i + 1)
			sp = Sparse(g["l%d_%s_idx" % (li, ch)].astype(np.int64), g["l%d_%s_val" % (li, ch)], (int(n), int(n), ncell))
			out.append(Chrom_Dataset(sp, bs_bin=int(g["bs_bin"][li]), bs_cell=ncell, compact=True, flank=off, chrom=ch,
			                         resolution=int(g["res"][li]), device=device))
	return out, g
