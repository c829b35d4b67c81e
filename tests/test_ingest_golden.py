"""Ingest stage against fixtures produced by the unmodified reference
(tests/golden/make_golden_ingest.py): get_qc and pack_training_data_one_process in five
configurations, for BOTH the product (fasthigashi_b200.ingest -> libfh_host.so, C++/OpenMP) and its
checker (oracle/ingest_oracle.py, numpy); then product vs checker on larger random inputs, the
per-resolution cache and the edge cases."""
import os
import numpy as np
import pytest
from scipy.sparse import csr_matrix
from conftest import GOLDEN
import __graft_entry__ as ge
ge.build_host()
from fasthigashi_b200 import ingest
from oracle import ingest_oracle

IMPLS = {"native": ingest, "oracle": ingest_oracle}

G = np.load(os.path.join(GOLDEN, "ingest_cases.npz"), allow_pickle=True)
CHROMS = [str(c) for c in G["chroms"]]
NCELL = int(G["ncell"])
RES = int(G["res"])


def raw_cells(ch):
	n = int(G["raw_%s_n" % ch])
	indptr = G["raw_%s_indptr" % ch].reshape(NCELL, n + 1)
	indices, data = G["raw_%s_indices" % ch], G["raw_%s_data" % ch]
	out, off = [], 0
	for c in range(NCELL):
		nnz = int(indptr[c, -1])
		out.append(csr_matrix((data[off:off + nnz], indices[off:off + nnz], indptr[c]), shape=(n, n)))
		off += nnz
	return out


@pytest.fixture(scope="module")
def raw_dir(tmp_path_factory):
	d = tmp_path_factory.mktemp("ingest")
	os.makedirs(d / "raw")
	for ch in CHROMS:
		arr = np.empty(NCELL, dtype=object)
		for i, m in enumerate(raw_cells(ch)):
			arr[i] = m
		np.save(d / "raw" / ("%s_sparse_adj.npy" % ch), arr, allow_pickle=True)
	return str(d)


def as_sorted(idx, val, shape):
	key = (idx[2].astype(np.int64) * shape[0] + idx[0]) * shape[1] + idx[1]
	o = np.argsort(key, kind="stable")
	return key[o], val[o]


def test_flatten_keeps_csr_order():
	mats = raw_cells(CHROMS[0])
	rc = ingest_oracle.flatten_cells(mats)
	coo = [m.tocoo() for m in mats]
	assert np.array_equal(rc.row, np.concatenate([m.row for m in coo]))
	assert np.array_equal(rc.col, np.concatenate([m.col for m in coo]))
	assert np.array_equal(rc.cell, np.concatenate([np.full(m.nnz, i) for i, m in enumerate(coo)]))
	assert rc.num_cell == NCELL and rc.shape == mats[0].shape


@pytest.mark.parametrize("impl", list(IMPLS))
def test_get_qc_matches_reference(raw_dir, impl):
	kept, reads = IMPLS[impl].get_qc(os.path.join(raw_dir, "raw"), CHROMS, RES)
	assert np.array_equal(kept, G["qc"]) and kept.dtype == G["qc"].dtype
	np.testing.assert_allclose(reads, G["readcount"], rtol=1e-6)
	assert 0 < kept.sum() < NCELL  # the fixture has both good and bad cells


CASES = {"plain": dict(off_diag=12, merge=1, batch=False, batch_norm=False, bl=False),
         "merge2": dict(off_diag=8, merge=2, batch=False, batch_norm=False, bl=False),
         "batch": dict(off_diag=12, merge=1, batch=True, batch_norm=True, bl=False),
         "batchoff": dict(off_diag=12, merge=1, batch=True, batch_norm=False, bl=False),
         "bl": dict(off_diag=12, merge=1, batch=False, batch_norm=False, bl=True)}


@pytest.mark.parametrize("impl", list(IMPLS))
@pytest.mark.parametrize("case", list(CASES))
def test_pack_training_data_matches_reference(raw_dir, case, impl):
	c = CASES[case]
	reorder = G["reorder"]
	bl = {"chr1": G["bl_chr1"], "chr2": G["bl_chr2"]} if c["bl"] else None
	batch = G["batch"][reorder] if c["batch"] else None
	for ch in CHROMS:
		idx, val, shape = IMPLS[impl].pack_training_data_one_process(
			os.path.join(raw_dir, "raw"), ch, reorder, c["off_diag"], c["merge"], c["merge"], batch, c["batch_norm"], bl)
		ridx, rval, rshape = G["%s_%s_idx" % (case, ch)], G["%s_%s_val" % (case, ch)], G["%s_%s_shape" % (case, ch)]
		assert tuple(shape) == tuple(int(s) for s in rshape)
		assert idx.dtype == np.int32 and val.dtype == np.float32
		k, v = as_sorted(idx, val, shape)
		rk, rv = as_sorted(ridx, rval, shape)
		if c["merge"] > 1:
			# the reference's sum_duplicates() after the in-place coarsening is a no-op (scipy keeps the
			# has_canonical_format flag of the CSR it came from), so colliding contacts stay separate
			# entries and its densify keeps an arbitrary one; here they are summed (DESIGN.md deviations)
			assert len(rk) > len(np.unique(rk))
			rk, inv = np.unique(rk, return_inverse=True)
			clipped = (np.bincount(inv, weights=(rv >= rv.max())) > 0) | (v >= v.max())  # at a mean+15 sigma cap (either side)
			rv = np.log1p(np.bincount(inv, weights=np.expm1(rv.astype(np.float64)))).astype(np.float32)
			assert np.array_equal(k, rk)
			assert clipped.sum() < 0.01 * len(v)
			np.testing.assert_allclose(v[~clipped], rv[~clipped], rtol=1e-5)
			continue
		assert np.array_equal(k, rk)  # same (row, col, cell) set: bit-exact integer work
		np.testing.assert_allclose(v, rv, rtol=2e-6, atol=1e-7)
		assert np.abs(idx[0].astype(int) - idx[1]).max() <= c["off_diag"]


def test_pack_feeds_block_csr(raw_dir):
	"""ingest -> Sparse -> Chrom_Dataset: the window rule of the staging layer accepts what ingest emits."""
	from fasthigashi_b200.sparse_for_schic import Sparse, Chrom_Dataset
	idx, val, shape = ingest.pack_training_data_one_process(os.path.join(raw_dir, "raw"), "chr1", G["reorder"], 12)
	ds = Chrom_Dataset(Sparse(idx.astype(np.int64), val, shape, copy=False), bs_bin=20, bs_cell=NCELL,
	                   good_qc_num=int(G["qc"].sum()), compact=True, flank=12, chrom="chr1", resolution=RES, device="cpu")
	assert ds.nnz() == len(val) and ds.total_cell_num == NCELL and ds.num_cell == int(G["qc"].sum())


def test_preprocess_contact_map_cache_roundtrip(raw_dir, tmp_path):
	cfg = {"chrom_list": CHROMS, "temp_dir": raw_dir, "resolution": RES}
	cache = str(tmp_path / "cache.pkl")
	a = ingest.preprocess_contact_map(cfg, G["reorder"], cache, 12, RES)
	assert os.path.exists(cache)
	b = ingest.preprocess_contact_map({"chrom_list": CHROMS, "temp_dir": "/nonexistent", "resolution": RES}, G["reorder"], cache, 12, RES)
	for (i0, v0, s0), (i1, v1, s1) in zip(a, b):
		assert np.array_equal(i0, i1) and np.array_equal(v0, v1) and tuple(s0) == tuple(s1)
	for ch, (i0, v0, s0) in zip(CHROMS, a):
		assert np.array_equal(np.sort(v0), np.sort(G["plain_%s_val" % ch])) or np.allclose(np.sort(v0), np.sort(G["plain_%s_val" % ch]), rtol=2e-6)


@pytest.mark.parametrize("impl", list(IMPLS))
def test_empty_cell_and_all_filtered(impl):
	n = 12
	mats = [csr_matrix(np.eye(n, dtype=np.float32) * 3), csr_matrix((n, n), dtype=np.float32),
	        csr_matrix(np.diag(np.ones(n - 1, dtype=np.float32), 1) + np.diag(np.ones(n - 1, dtype=np.float32), -1))]
	wrap = ingest.CellMatrices if impl == "native" else ingest_oracle.flatten_cells
	idx, val, shape = IMPLS[impl].pack_training_data_one_process(None, "chrX", None, off_diag=0, raw=wrap(mats))
	assert shape == (n, n, 3)
	assert set(idx[2].tolist()) == {0}  # cell 1 is empty, cell 2 only has off-diagonal contacts
	assert np.all(idx[0] == idx[1])
	# every contact filtered out: empty tensor, no exception
	idx, val, shape = IMPLS[impl].pack_training_data_one_process(None, "chrX", None, off_diag=0, raw=wrap(mats[2:]))
	assert idx.shape == (3, 0) and val.shape == (0,) and shape == (0, 0, 1)


def random_cells(rng, ncell, n, dtype, index_dtype, symmetric=True):
	mats = []
	for c in range(ncell):
		k = int(rng.integers(0, 6 * n))
		i = rng.integers(0, n, size=k)
		j = np.clip(i + rng.integers(-25, 26, size=k), 0, n - 1)
		v = rng.integers(1, 5, size=k) if np.issubdtype(dtype, np.integer) else rng.random(k) * 3 + 0.1
		m = np.zeros((n, n))
		np.add.at(m, (i, j), v)
		if symmetric:
			m = m + m.T
		m = csr_matrix(m.astype(dtype))
		m.indices = m.indices.astype(index_dtype)
		m.indptr = m.indptr.astype(index_dtype)
		mats.append(m)
	return mats


@pytest.mark.parametrize("dtype,index_dtype,merge,batches,bl", [
	(np.float32, np.int32, 1, 0, False), (np.float64, np.int64, 1, 3, False), (np.float64, np.int32, 3, 2, True),
	(np.int64, np.int32, 2, 0, True), (np.int32, np.int64, 1, 4, False)])
def test_native_matches_checker_on_random_inputs(dtype, index_dtype, merge, batches, bl):
	"""libfh_host.so against the pinned numpy restatement: same (row, col, cell) set bit for bit, values to
	fp32 rounding; non-integer weights, every scipy dtype combination, coarsening + blacklist + batches."""
	rng = np.random.default_rng(hash((str(dtype), merge, batches)) % 2 ** 31)
	ncell, n = 157, 83
	mats = random_cells(rng, ncell, n, dtype, index_dtype)
	batch = rng.integers(0, batches, size=ncell).astype(str) if batches else None
	blk = {"chrT": rng.choice(n, size=5, replace=False)} if bl else None
	a = ingest.pack_training_data_one_process(None, "chrT", None, 9, merge, merge, batch, True, blk, raw=ingest.CellMatrices(mats))
	b = ingest_oracle.pack_training_data_one_process(None, "chrT", None, 9, merge, merge, batch, True, blk,
	                                                 raw=ingest_oracle.flatten_cells(mats))
	assert a[2] == tuple(int(x) for x in b[2])
	ka, va = as_sorted(a[0], a[1], a[2])
	kb, vb = as_sorted(b[0], b[1], b[2])
	assert np.array_equal(ka, kb)
	np.testing.assert_allclose(va, vb, rtol=3e-6, atol=1e-7)
	if merge == 1:  # same entry ORDER as the reference too (cell-major, CSR order inside a cell)
		assert np.array_equal(a[0], b[0])


def test_native_thread_count_does_not_change_the_result(monkeypatch):
	rng = np.random.default_rng(1)
	mats = random_cells(rng, 64, 40, np.float32, np.int32)
	batch = rng.integers(0, 2, size=64)
	out = []
	for t in ("1", "3", "0"):
		monkeypatch.setenv("FH_HOST_THREADS", t)
		out.append(ingest.pack_training_data_one_process(None, "c", None, 6, 1, 1, batch, True, None, raw=ingest.CellMatrices(mats)))
	for o in out[1:]:
		assert np.array_equal(o[0], out[0][0])
		np.testing.assert_allclose(o[1], out[0][1], rtol=1e-6)


def test_native_rejects_bad_input():
	m = csr_matrix(np.eye(5, dtype=np.float32))
	m2 = csr_matrix(np.eye(6, dtype=np.float32))
	with pytest.raises(ingest.IngestError):
		ingest.CellMatrices([m, m2])
	with pytest.raises(ingest.IngestError):
		ingest.CellMatrices([])
	bad = csr_matrix(np.eye(5, dtype=np.float32))
	bad.indices = bad.indices.copy(); bad.indices[2] = 9
	with pytest.raises(ingest.IngestError, match="column index"):
		ingest.pack_training_data_one_process(None, "c", None, 2, raw=ingest.CellMatrices([bad]))
	with pytest.raises(ingest.IngestError, match="batch_id"):
		ingest.pack_training_data_one_process(None, "c", None, 2, batch_id=[0, 1], raw=ingest.CellMatrices([m]))
