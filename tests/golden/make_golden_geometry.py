"""Golden block geometry at the FULL sizes of every BASELINE.json config: the slice lists the
unmodified reference `Chrom_Dataset` (sparse_for_schic.py:356-510) builds for each chromosome with
the wrapper's GPU batching rule (FastHigashi_Wrapper.py:501-512), off_diag = 100.
Re-run (this container only):  python tests/golden/make_golden_geometry.py
Output tests/golden/geometry_cases.npz: per case/chromosome an int array (blocks, 6):
[row_start, row_stop, local_start, local_stop, col_start, col_stop] (stops clipped as python slices are)."""
import math
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import fasthigashi_b200  # noqa: E402,F401
from fasthigashi_b200 import synth  # noqa: E402
from oracle import ref_shims  # noqa: E402

CASES = {"pfc_500kb": ("pfc", 500000), "hg19_1mb": ("hg19", 1000000), "hg19_500kb": ("hg19", 500000), "hg19_100kb": ("hg19", 100000)}
OFF_DIAG = 100


def bs_bin_rule(n, res):
	rec = min(max(int(15000000 / res), 128), 256)
	return math.ceil(n / max(math.ceil(n / rec), 1))


def main():
	mods = ref_shims.import_reference()
	sp = mods["sparse_for_schic"]
	out = {}
	for case, (kind, res) in CASES.items():
		for ci, n in enumerate(synth.chrom_bins(kind, res)):
			# one cell, one diagonal contact per bin: geometry does not depend on the data
			idx = np.stack([np.arange(n), np.arange(n), np.zeros(n, dtype=np.int64)]).astype(np.int32)
			t = sp.Sparse(idx, np.ones(n, dtype=np.float32), np.asarray([n, n, 1]), copy=True)
			t.sort_indices()
			ds = sp.Chrom_Dataset(tensor=t, bs_bin=bs_bin_rule(n, res), bs_cell=1, good_qc_num=-1, kind="hic", upper_sim=False,
			                      compact=True, flank=OFF_DIAG, chrom="chr%d" % (ci + 1), resolution=res)
			rows = []
			for b in range(ds.num_bin_batch):
				r, l, c = ds.bin_slice_list[b], ds.local_bin_slice_list[b], ds.col_bin_slice_list[b]
				rows.append([r.start, min(r.stop, n), l.start, l.stop, c.start or 0, min(c.stop, n) if c.stop is not None else n])
			out["%s_chr%d" % (case, ci + 1)] = np.asarray(rows, dtype=np.int64)
			out["%s_chr%d_n" % (case, ci + 1)] = n
	np.savez_compressed(os.path.join(HERE, "geometry_cases.npz"), **out)
	print("wrote", len(out) // 2, "chromosome geometries")


if __name__ == "__main__":
	main()
