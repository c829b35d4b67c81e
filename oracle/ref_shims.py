"""TEST INFRASTRUCTURE ONLY. Imports the *unmodified* reference from /root/reference (this
container only; the path does not exist on the GPU box) with the four out-of-tree shims of
SURVEY.md §8c. Used by tests/golden/make_golden.py to generate the committed fixtures and by
tests that pin oracle/fh_oracle.py against the real reference when it is present.

Nothing in the product path may import this module.
"""
import os
import sys
import types
import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("FH_REFERENCE_ROOT", "/root/reference")


def reference_available():
	return os.path.isdir(os.path.join(REFERENCE_ROOT, "fasthigashi"))


def import_reference():
	"""Returns the dict of reference modules. Raises ImportError when the reference is absent."""
	if not reference_available():
		raise ImportError("reference not present at %s" % REFERENCE_ROOT)
	# shim 1: opt_einsum.contract -> torch.einsum (parafac2_intergrative.py:5, parafac_integrative.py:8)
	if "opt_einsum" not in sys.modules:
		oe = types.ModuleType("opt_einsum")
		oe.contract = lambda f, *ops: torch.einsum(f.replace(" ", ""), *ops)
		sys.modules["opt_einsum"] = oe
	# shim 2: h5py is imported but unused on this path (FastHigashi_Wrapper.py:5)
	if "h5py" not in sys.modules:
		try:
			import h5py  # noqa: F401
		except Exception:
			sys.modules["h5py"] = types.ModuleType("h5py")
	if REFERENCE_ROOT not in sys.path:
		sys.path.insert(0, REFERENCE_ROOT)
	import importlib
	mods = {}
	for name in ["partial_rwr", "project2orthogonal", "sparse_for_schic", "parafac_integrative",
	             "parafac2_intergrative"]:
		mods[name] = importlib.import_module("fasthigashi." + name)
	sp = mods["sparse_for_schic"]
	# shim 3: numpy-2 "Unable to avoid copy" for a tuple shape (sparse_for_schic.py:65)
	if not getattr(sp.Sparse, "_fh_shimmed", False):
		_orig = sp.Sparse.__init__

		def _init(self, indices, values, shape, *a, **k):
			return _orig(self, indices, values, np.asarray(shape), *a, **k)
		sp.Sparse.__init__ = _init
		sp.Sparse._fh_shimmed = True
		# shim 4: pin_memory needs a CUDA driver (sparse_for_schic.py:576); no-op on CPU boxes
		if not torch.cuda.is_available():
			sp.Chrom_Dataset.pin_memory = lambda self: None
	return mods


def build_reference_datasets(mods, chroms, bs_bin_rule="cpu", off_diag=100, res=1000000,
                             good_qc_num=-1, bs_bin=None, bs_cell=None):
	"""chroms: output of fasthigashi_b200.synth.synth_dataset. Builds the reference's own
	Sparse -> Chrom_Dataset objects (FastHigashi_Wrapper.py:396,525-535)."""
	sp = mods["sparse_for_schic"]
	out = []
	for ch in chroms:
		idx = ch["indices"].cpu().numpy().astype(np.int32)
		# reference wants dim-0 sorted (Sparse.sort_indices :133-151)
		order = np.lexsort((idx[1], idx[2], idx[0]))
		idx = np.ascontiguousarray(idx[:, order])
		val = np.ascontiguousarray(ch["values"].cpu().numpy()[order])
		obj = sp.Sparse(idx, val, np.asarray(ch["shape"]), copy=True)
		obj.sort_indices()
		n = ch["n"]
		ncell = ch["shape"][2]
		if bs_bin is not None:
			bb = bs_bin if isinstance(bs_bin, int) else bs_bin[ch["chrom"]]
		elif bs_bin_rule == "cpu":
			bb = n
		else:
			import math
			rec = min(max(int(15000000 / res), 128), 256)
			bb = math.ceil(n / max(math.ceil(n / rec), 1))
		bc = ncell if bs_cell is None else bs_cell
		out.append(sp.Chrom_Dataset(tensor=obj, bs_bin=bb, bs_cell=bc, good_qc_num=good_qc_num,
		                            kind="hic", upper_sim=False, compact=True, flank=off_diag,
		                            chrom=ch["chrom"], resolution=res))
	return out
