"""Host logic: block geometry / block-CSR staging (sparse_for_schic.py:356-510 mirror)."""
import numpy as np
import pytest
import torch
from conftest import load_small_dataset
from fasthigashi_b200.sparse_for_schic import Sparse, Chrom_Dataset, block_geometry


def test_geometry_rule():
	# hand-derived from sparse_for_schic.py:457-492 (n=90, bs_bin=32, flank=12)
	g = block_geometry(90, 32, 12)
	assert [(x.row0, x.nb, x.col0, x.w, x.s, x.e) for x in g] == [
		(0, 32, 0, 44, 0, 32), (32, 32, 20, 56, 12, 44), (64, 26, 52, 38, 12, 38)]
	# i == flank keeps absolute columns (reference uses `i > flank`)
	g = block_geometry(40, 12, 12)
	assert (g[1].col0, g[1].s) == (0, 12)
	assert (g[2].col0, g[2].s) == (12, 12)
	g = block_geometry(50, 50, 100)
	assert (g[0].w, g[0].s, g[0].e) == (50, 0, 50)


def test_csr_roundtrip_and_cell_order():
	ds = load_small_dataset(good_qc_num=44, bs_cell=20)[0]
	assert ds.num_cell == 44 and ds.total_cell_num == 48
	assert [(s.start, s.stop) for s in ds.cell_slice_list] == [(0, 20), (20, 40), (40, 44), (44, 48)]
	assert ds.num_cell_batch == 3 and ds.num_cell_batch_bad == 1
	d = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "data_small.npz"))
	idx, val = d["chr1_idx"].astype(np.int64), d["chr1_val"]
	dense = np.zeros((90, 90, 48), np.float32)
	dense[idx[0], idx[1], idx[2]] = val
	for b, g in enumerate(ds.geoms):
		rp, col, v = ds.cell_range_csr(b, 0, 48)
		rows = np.repeat(np.arange(48 * g.nb), np.diff(rp.numpy()))
		out = np.zeros((48 * g.nb, g.w), np.float32)
		out[rows, col.numpy()] = v.numpy()
		ref = dense[g.row0:g.row0 + g.nb, g.col0:g.col0 + g.w, :].transpose(2, 0, 1).reshape(-1, g.w)
		assert np.array_equal(out, ref)
	assert ds.nnz() == len(val)


def test_rejects_out_of_window_and_duplicates():
	idx = np.array([[0, 5], [40, 6], [0, 0]])
	with pytest.raises(ValueError):
		Chrom_Dataset(Sparse(idx, np.ones(2, np.float32), (60, 60, 1)), 16, 1, compact=True, flank=10)
	idx = np.array([[3, 3], [4, 4], [0, 0]])
	with pytest.raises(ValueError):
		Chrom_Dataset(Sparse(idx, np.ones(2, np.float32), (60, 60, 1)), 16, 1, compact=True, flank=10)


def test_empty_cells_and_select_cells():
	idx = np.array([[0, 1, 7], [1, 0, 7], [2, 2, 2]])
	ds = Chrom_Dataset(Sparse(idx, np.array([1., 2., 3.], np.float32), (9, 9, 4)), 4, 4, compact=True, flank=3)
	assert ds.nnz() == 3 and len(ds.geoms) == 3
	sub = ds.select_cells(2, 4)
	assert sub.total_cell_num == 2 and sub.nnz() == 3
	sub0 = ds.select_cells(0, 2)
	assert sub0.nnz() == 0 and all(int(r[-1]) == 0 for r in sub0.rowptr)


def test_geometry_matches_reference_at_full_baseline_sizes():
	"""All four BASELINE geometries (PFC @500 kb, hg19 @1 Mb / 500 kb / 100 kb, off_diag 100, the wrapper's GPU
	batching rule): the slice lists of the unmodified reference's Chrom_Dataset (tests/golden/make_golden_geometry.py)."""
	import math, os
	from conftest import GOLDEN
	from fasthigashi_b200 import synth
	G = np.load(os.path.join(GOLDEN, "geometry_cases.npz"))
	blocks = 0
	for case, (kind, res) in {"pfc_500kb": ("pfc", 500000), "hg19_1mb": ("hg19", 1000000), "hg19_500kb": ("hg19", 500000),
	                          "hg19_100kb": ("hg19", 100000)}.items():
		for ci, n in enumerate(synth.chrom_bins(kind, res)):
			ref = G["%s_chr%d" % (case, ci + 1)]
			assert int(G["%s_chr%d_n" % (case, ci + 1)]) == n
			rec = min(max(int(15000000 / res), 128), 256)
			bs = math.ceil(n / max(math.ceil(n / rec), 1))
			mine = np.asarray([[g.row0, g.row0 + g.nb, g.s, g.e, g.col0, g.col0 + g.w] for g in block_geometry(n, bs, 100)])
			assert np.array_equal(mine, ref), (case, ci)
			blocks += len(mine)
	assert blocks == 348


def _random_coo(rng, n, ncell, flank, density, dtype):
	"""Unique (row, col, cell) entries inside the band |col - row| <= flank, in random order."""
	r, c, z = np.meshgrid(np.arange(n), np.arange(n), np.arange(ncell), indexing="ij")
	keep = (np.abs(r - c) <= flank) & (rng.random(r.shape) < density)
	idx = np.stack([r[keep], c[keep], z[keep]]).astype(dtype)
	perm = rng.permutation(idx.shape[1])
	return idx[:, perm], rng.random(idx.shape[1]).astype(np.float32)


def _sort_route(idx, val, shape, **kw):
	"""The device-sort route of Chrom_Dataset (taken when the COO is already on the GPU), run on CPU tensors."""
	ds = object.__new__(Chrom_Dataset)
	ref = Chrom_Dataset(Sparse(np.zeros((3, 0), np.int64), np.zeros(0, np.float32), shape), **kw)
	ds.__dict__.update(ref.__dict__)
	ds._build_sort(torch.as_tensor(idx.astype(np.int64)), torch.as_tensor(val), torch.device("cpu"))
	return ds


@pytest.mark.parametrize("dtype", [np.int32, np.int64])
@pytest.mark.parametrize("n,bs_bin,flank,compact", [(90, 32, 12, True), (50, 50, 100, True), (37, 8, 5, True), (40, 16, 40, False)])
def test_native_block_csr_equals_sort_route(dtype, n, bs_bin, flank, compact, monkeypatch):
	"""libfh_host's counting passes (host COO) and the 64-bit key sort (device COO) give the same arrays bit for bit, for
	shuffled input, both index widths, empty cells/rows, and any host thread count."""
	rng = np.random.default_rng(n + bs_bin)
	ncell = 7
	idx, val = _random_coo(rng, n, ncell, flank, 0.3, dtype)
	idx = idx[:, idx[2] != 3]            # one cell without contacts
	val = val[:idx.shape[1]]
	kw = dict(bs_bin=bs_bin, bs_cell=4, good_qc_num=5, compact=compact, flank=flank)
	want = _sort_route(idx, val, (n, n, ncell), **kw)
	for threads in ("1", "3", "0"):
		monkeypatch.setenv("FH_HOST_THREADS", threads)
		got = Chrom_Dataset(Sparse(idx, val, (n, n, ncell), copy=False), **kw)
		assert len(got.rowptr) == len(want.rowptr) == len(got.geoms)
		for b in range(len(got.geoms)):
			assert got.rowptr[b].dtype == torch.int32 and got.col[b].dtype == torch.int16 and got.val[b].dtype == torch.float32
			assert torch.equal(got.rowptr[b], want.rowptr[b])
			assert torch.equal(got.col[b], want.col[b])
			assert torch.equal(got.val[b], want.val[b])
		assert got.nnz() == idx.shape[1]


def test_native_block_csr_empty_tensor_and_errors():
	from fasthigashi_b200 import ingest
	ds = Chrom_Dataset(Sparse(np.zeros((3, 0), np.int64), np.zeros(0, np.float32), (20, 20, 3)), 8, 3, compact=True, flank=4)
	assert ds.nnz() == 0 and [len(r) for r in ds.rowptr] == [3 * 8 + 1, 3 * 8 + 1, 3 * 4 + 1]
	# error codes of the C ABI (include/fh_host.h)
	geom = dict(num_bin=20, bs_bin=8, num_cell=3, nb=[8, 8, 4], col0=[0, 4, 12], w=[12, 16, 8])
	with pytest.raises(ingest.IngestError) as e:
		ingest.block_csr(np.array([[9], [3], [0]]), np.ones(1, np.float32), **geom)       # column left of the window
	assert e.value.code == -4
	with pytest.raises(ingest.IngestError) as e:
		ingest.block_csr(np.array([[9, 9], [5, 5], [1, 1]]), np.ones(2, np.float32), **geom)
	assert e.value.code == -5 and "cell 1" in str(e.value)
	with pytest.raises(ingest.IngestError) as e:
		ingest.block_csr(np.array([[20], [19], [0]]), np.ones(1, np.float32), **geom)     # row outside the tensor
	assert e.value.code == -1
	with pytest.raises(ingest.IngestError) as e:
		ingest.block_csr(np.array([[1], [1], [3]]), np.ones(1, np.float32), **geom)       # cell outside the tensor
	assert e.value.code == -1
	with pytest.raises(ingest.IngestError):
		ingest.block_csr(np.zeros((3, 0), np.int64), np.zeros(0, np.float32), 20, 8, 3, [8, 8, 4], [0, 4, 12], [12, 40000, 8])


def test_sparse_container_known_answers_of_the_reference():
	"""The reference's only test for this path, `sparse_for_schic.test()` (:634-661), restated on this package's `Sparse`:
	sort, permute, reshape, slicing and their composition against dense numpy."""
	def new_obj():
		return Sparse([[0, 2, 1, 0], [0, 3, 2, 1]], [1, 2, 3, 4.], (3, 4))
	dense = lambda o: np.asarray(o.to_scipy().todense())
	base = dense(new_obj())
	assert np.array_equal(base, np.array([[1, 4, 0, 0], [0, 0, 3, 0], [0, 0, 0, 2.]]))
	o = new_obj(); o.sort_indices()
	assert np.array_equal(dense(o), base) and list(o.indptr) == [0, 2, 3, 4]
	assert np.array_equal(dense(new_obj().permute(0, 1)), base)
	assert np.array_equal(dense(new_obj().permute(1, 0)), base.T)
	assert np.array_equal(dense(new_obj().reshape(1, 12)), base.reshape(1, 12))
	assert np.array_equal(dense(new_obj().reshape(4, 3)), base.reshape(4, 3))
	assert np.array_equal(dense(new_obj().reshape(-1, 6)), base.reshape(2, 6))
	for s in [slice(2), slice(10), slice(0, None), slice(2, None), slice(1, 3)]:
		assert np.array_equal(dense(new_obj().slicing(s)), base[s])
		assert np.array_equal(dense(new_obj()[s]), base[s])
	o = new_obj().permute(1, 0).reshape(6, 2).slicing(slice(1, 3)).permute(1, 0)
	assert np.array_equal(dense(o), base.T.reshape(6, 2)[1:3].T)
	assert np.array_equal(new_obj().to_dense().numpy(), base) and new_obj().numel() == 12 and len(new_obj()) == 3
	assert np.array_equal(new_obj()[0].to_dense().numpy(), base[0])
	o = new_obj(); o.filter_max_distance(1)
	assert np.array_equal(dense(o), np.triu(np.tril(base, 1), -1))
	assert np.array_equal(np.asarray(new_obj().to_csr().todense()), base)


def test_sparse_container_matches_reference_class_on_random_tensors():
	from oracle import ref_shims
	if not ref_shims.reference_available():
		pytest.skip("the reference checkout is only present in the build container")
	R = ref_shims.import_reference()["sparse_for_schic"].Sparse
	rng = np.random.default_rng(4)
	shape = (7, 5, 6)
	flat = rng.choice(int(np.prod(shape)), size=60, replace=False)
	idx = np.stack(np.unravel_index(flat, shape)).astype(np.int64)
	val = rng.random(60).astype(np.float32)
	ours, ref = Sparse(idx, val, shape), R(idx.copy(), val.copy(), np.asarray(shape))
	assert np.array_equal(ours.to_dense().numpy(), ref.to_dense())
	for perm in [(2, 0, 1), (1, 2, 0)]:
		assert np.array_equal(ours.permute(*perm).to_dense().numpy(), ref.permute(*perm).to_dense())
	for dims in [(35, 6), (7, 30), (5, -1, 3)]:
		assert np.array_equal(ours.reshape(*dims).to_dense().numpy(), ref.reshape(*dims).to_dense())
	for sl in [slice(1, 4), slice(3, None), slice(0, 20)]:
		a, b = ours.get_slice_idx_value(sl), ref.get_slice_idx_value(sl)
		assert a[2] == tuple(b[2]) and a[3] == b[3]
		assert np.array_equal(ours[sl].to_dense().numpy(), ref[sl].to_dense())
	assert np.array_equal(ours[2].to_dense().numpy(), ref[2].to_dense())
