"""The product's Python orchestration (`Fast_Higashi_core`: operand layouts / strides handed to the C ABI, the order of
the stages, what is all-reduced when cells are sharded) run on the CPU against tests/emu/fake_abi.py - a host-memory
stand-in that evaluates the documented contract of every include/fh_b200.h entry point in fp64. The CUDA kernels are NOT
exercised here (tests/test_gpu_parity.py does that on the B200); these tests pin everything between the reference-facing
API and the C ABI to the reference's own runs (tests/golden/core_*.npz)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_small_dataset, load_multires_dataset, rel_fro

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
import fake_abi  # noqa: E402
from oracle import fh_oracle as O  # noqa: E402


@pytest.fixture
def fake():
	f, undo = fake_abi.install()
	try:
		yield f
	finally:
		undo()


def _core(rank, off_diag, res_list, **kw):
	from fasthigashi_b200.parafac2_intergrative import Fast_Higashi_core
	return Fast_Higashi_core(rank, off_diag, res_list, **kw).to("cpu")


def _state(g, nds, nchrom):
	return ([g["t0_A%d" % i] for i in range(nds)], [g["t0_B%d" % c] for c in range(nchrom)], [g["t0_D%d" % c] for c in range(nchrom)],
	        g["t0_V"], [g["bin_cov%d" % i] for i in range(nds)],
	        [g["bad_bin_cov%d" % i] if "bad_bin_cov%d" % i in g.files else 0 for i in range(nds)], g["n_i"])


@pytest.mark.parametrize("tag", ["nocol", "col"])
def test_full_run_matches_reference_fixture(fake, tag):
	"""fit_transform from the shared seeds: init (auto-stop RWR counts, coverage, pooled features, host SVD), every sweep's
	loss, transform incl. bad-QC cells, embeddings - against the unmodified reference's run."""
	g = np.load(os.path.join(GOLDEN, "core_%s.npz" % tag))
	good = int(g["good_qc_num"])
	ds = load_small_dataset(good_qc_num=good if good < 48 else -1, bs_cell=int(g["bs_cell"]))
	core = _core(int(g["rank"]), 12, [1000000])
	torch.manual_seed(0); np.random.seed(0)
	nsweep = int(g["nsweep"])
	_, (A_list, B_list, D_list, V), proj = core.fit_transform(ds, size_ratio=0.3, n_iter_max=nsweep, n_iter_parafac=1, do_conv=True,
	                                                         do_rwr=True, do_col=bool(g["do_col"]), tol=0.0, verbose=False)
	assert list(core.n_i) == list(g["n_i"])
	re = np.array(core.re_trace)
	assert np.max(np.abs(re - g["re"]) / g["re"]) < 1e-4, (re, g["re"])
	assert tuple(V.shape) == (48, int(g["rank"]))                      # good AND bad-QC cells
	E = O.embed_all(V.numpy(), [d.numpy() for d in D_list])
	Eref = O.embed_all(g["final_V"], [g["final_D%d" % i] for i in range(3)])
	for j in range(E.shape[1]):
		assert abs(np.corrcoef(E[:, j], Eref[:, j])[0, 1]) > 0.999
	for i in range(3):
		fin = np.isfinite(g["bin_cov%d" % i])
		assert np.array_equal(np.isfinite(core.bin_cov_list[i].numpy()), fin)
		assert rel_fro(core.bin_cov_list[i].numpy()[fin], g["bin_cov%d" % i][fin]) < 1e-5
		for b, U in enumerate(proj[i]):
			assert tuple(U.shape) == g["final_U%d_%d" % (i, b)].shape
	assert fake.calls["rwr_batched"] > 0 and fake.calls["polar_isqrt_multi"] == nsweep and fake.calls["cp_als"] == 3 * nsweep


def test_lockstep_from_reference_state_and_rwr_cache_modes(fake):
	"""From the reference's init state: per-sweep loss terms and the projected tensors; cache='run' (one RWR pass per
	run) gives the same numbers as cache='sweep' with a third of the RWR calls."""
	g = np.load(os.path.join(GOLDEN, "core_nocol.npz"))
	out = {}
	for cache in ("sweep", "run"):
		fake.calls.clear()
		core = _core(int(g["rank"]), 12, [1000000], cache=cache)
		core.fit(load_small_dataset(), 0.3, 3, 1, True, True, False, 0.0, verbose=False, state=_state(g, 3, 3))
		out[cache] = (np.array(core.re_trace), fake.calls["rwr_batched"], core)
	re, calls, core = out["sweep"]
	assert np.max(np.abs(re - g["re"][:3]) / g["re"][:3]) < 1e-4
	for t in range(3):
		assert rel_fro(core.loss_terms[t]["x_U"], g["t%d_x_U" % t]) < 1e-4
		assert abs(core.loss_terms[t]["x_V"] - float(g["t%d_x_V" % t])) / abs(float(g["t%d_x_V" % t])) < 1e-4
	assert rel_fro(core.loss_terms[0]["xnorm"], g["xnorm"]) < 1e-5
	assert np.allclose(out["run"][0], re, rtol=1e-6) and out["run"][1] * 3 == calls


def test_multi_resolution_matches_reference_fixture(fake):
	"""Two resolutions of the same chromosomes (shared B / D, bins stacked in the projected tensor, one CP-ALS per
	chromosome): lock-step from the reference's state, then the full run from the seeds."""
	ds, g = load_multires_dataset()
	res_list = [int(r) for r in g["res"]]
	nchrom, nsweep = len(g["chrom2size"]), int(g["nsweep"])
	core = _core(int(g["rank"]), int(g["off_diag"]), res_list)
	core.fit(ds, 0.3, 2, 1, True, True, False, 0.0, verbose=False, state=_state(g, len(ds), nchrom))
	assert list(core.chrom2size.values()) == list(g["chrom2size"])
	assert np.max(np.abs(np.array(core.re_trace) - g["re"][:2]) / g["re"][:2]) < 1e-4
	for c, chrom in enumerate(core.chrom2size):
		assert rel_fro(core.projected_tensor_list[chrom].numpy(), g["t1_Y%d" % c]) < 1e-3    # after sweep 2 = reference's t1 return
	for i in range(len(ds)):
		assert ds[i].global_slice_bin == slice(0 if i < nchrom else ds[i - nchrom].num_bin, (0 if i < nchrom else ds[i - nchrom].num_bin) + ds[i].num_bin)
	core = _core(int(g["rank"]), int(g["off_diag"]), res_list)
	torch.manual_seed(0); np.random.seed(0)
	ds2, _ = load_multires_dataset()
	_, (A_list, B_list, D_list, V), _ = core.fit_transform(ds2, size_ratio=0.3, n_iter_max=nsweep, n_iter_parafac=1, do_conv=True,
	                                                      do_rwr=True, do_col=False, tol=0.0, verbose=False)
	assert list(core.n_i) == list(g["n_i"])
	re = np.array(core.re_trace)
	assert np.max(np.abs(re - g["re"]) / g["re"]) < 1e-4, (re, g["re"])
	E = O.embed_all(V.numpy(), [d.numpy() for d in D_list])
	Eref = O.embed_all(g["final_V"], [g["final_D%d" % c] for c in range(nchrom)])
	for j in range(E.shape[1]):
		assert abs(np.corrcoef(E[:, j], Eref[:, j])[0, 1]) > 0.999


def test_wide_windows_take_the_row_gram_route(fake):
	"""r > window width (the last block of chr3: 8 rows x 20 columns against r = 36): temp_i is wide, the Gram is
	T T^T and U = M T (the FH_GEMM_F64xF32_F32 call) - against the pinned oracle from the same seeds."""
	core = _core(40, 12, [1000000])
	torch.manual_seed(0); np.random.seed(0)
	core.fit(load_small_dataset(), 0.9, 3, 1, True, True, False, 0.0, verbose=False)
	assert core.chrom2size["chr3"] == 36 and core.schic[2].geoms[1].w == 20
	oc = O.OracleCore(40, 12, [1000000])
	torch.manual_seed(0); np.random.seed(0)
	oc.fit(load_small_dataset(), 0.9, 3, 1, True, True, False, 0.0)
	re, ro = np.array(core.re_trace), np.array(oc.re_trace)
	assert np.max(np.abs(re - ro) / ro) < 1e-4, (re, ro)
	U = core.projection_list[1][2]                       # chr2's last block, (6, 18, 40): wide too, and well conditioned
	assert tuple(U.shape) == (6, 18, 40)
	assert torch.allclose(U @ U.transpose(1, 2), torch.eye(18).expand(6, 18, 18), atol=1e-4)   # orthonormal ROWS when wide


def test_device_init_svd_reaches_the_host_init_loss(fake):
	"""init_svd='device' (cell-sharded randomized SVD, no gather) is another random start: same loss level after a few sweeps."""
	res = {}
	for mode in ("host", "device"):
		core = _core(16, 12, [1000000], init_svd=mode)
		torch.manual_seed(0); np.random.seed(0)
		core.fit(load_small_dataset(), 0.3, 6, 1, True, True, False, 0.0, verbose=False)
		res[mode] = core.re_trace[-1]
	assert abs(res["device"] - res["host"]) / res["host"] < 0.02, res


def _cpu_wrapper(tmp_path, off, res, chroms):
	"""A `FastHigashi` object bound to the host-memory stand-in: the real constructor (with the three torch.cuda probes it
	makes answered for it), then device 'cpu'. Only valid while the `fake` fixture is installed."""
	import json
	from unittest import mock
	from fasthigashi_b200.FastHigashi_Wrapper import FastHigashi
	cfg = dict(chrom_list=chroms, temp_dir=str(tmp_path), data_dir=str(tmp_path), resolution=res, resolution_fh=[res])
	cfg_path = tmp_path / ("config_%d.JSON" % os.getpid())         # per process: the gloo workers share tmp_path
	with open(cfg_path, "w") as f:
		json.dump(cfg, f)
	with mock.patch("torch.cuda.is_available", lambda: True), mock.patch("torch.cuda.current_device", lambda: 0), \
	     mock.patch("torch.cuda.mem_get_info", lambda *a: (64 << 30, 180 << 30)):
		w = FastHigashi(str(cfg_path), None, None, off, True, True, True, False, False)
	w.device = "cpu"
	return w


def test_wrapper_only_partial_rwr(fake, tmp_path):
	"""FastHigashi.only_partial_rwr (FastHigashi_Wrapper.py:569-655): block pasting, symmetrisation, original cell ids,
	good and bad-QC batches - the same case the B200 runs (tests/wrapper_cases.py)."""
	import wrapper_cases
	wrapper_cases.case_only_partial_rwr_matches_oracle(_cpu_wrapper, tmp_path)


def test_wrapper_from_raw_files(fake, tmp_path):
	"""prep_dataset from raw/{chrom}_sparse_adj.npy (libfh_host.so ingest + block-CSR staging) -> run_model ->
	fetch_cell_embedding, against the oracle fed with the REFERENCE's packed tensors of the same raw files."""
	import wrapper_cases
	wrapper_cases.case_wrapper_from_raw_files(_cpu_wrapper, tmp_path)


# ---------------------------------------------------------------------------------------------------------------
def _free_port():
	s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _sharded_worker(rank, world, port, q, do_col, good, bs_cell, init_svd):
	import torch.distributed as dist
	os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
	dist.init_process_group("gloo", rank=rank, world_size=world)
	torch.set_num_threads(2)
	fake_abi.install()
	from fasthigashi_b200.sharding import shard_datasets, gather_cell_rows
	from fasthigashi_b200.parafac2_intergrative import Fast_Higashi_core
	mine = shard_datasets(load_small_dataset(good_qc_num=good, bs_cell=bs_cell), world, rank)
	core = Fast_Higashi_core(16, 12, [1000000], group=dist.group.WORLD, init_svd=init_svd).to("cpu")
	torch.manual_seed(0); np.random.seed(0)
	_, (A_list, B_list, D_list, V), _ = core.fit_transform(mine, size_ratio=0.3, n_iter_max=4, n_iter_parafac=1, do_conv=True,
	                                                      do_rwr=True, do_col=do_col, tol=0.0, verbose=False)
	V_all = gather_cell_rows(V, mine[0].num_cell, dist.group.WORLD)
	q.put((rank, list(core.n_i), list(core.re_trace), V_all.numpy(), [a.numpy() for a in A_list], [d.numpy() for d in D_list]))
	dist.destroy_process_group()


@pytest.mark.parametrize("do_col,good,bs_cell,init_svd,world", [(False, -1, 24, "host", 2), (True, 44, 22, "host", 2), (False, -1, 24, "device", 2),
                                                                (True, -1, 16, "host", 3)])
def test_two_rank_gloo_run_equals_single_process(do_col, good, bs_cell, init_svd, world):
	"""The whole cell-sharded run (init with features gathered to rank 0 and the MAX of the RWR step counts, all-reduced
	T1 / Gram / Y, per-bin polar problems partitioned over the ranks, factor broadcast, transform with bad-QC cells) on two
	gloo ranks against the same run in one process: same n_i, loss trace to 1e-6, identical factors on both ranks,
	embeddings of all cells in the unsharded order. `bs_cell` is chosen so that the good-cell batches of the single process
	are the two slabs: the reference's auto-stop in `init_params` is per CELL BATCH (partial_rwr.py:119-123), so the init
	features - and with them the whole run - depend on the batch composition (1e-4 on the loss with other batch sizes).
	init_svd='device': the init SVDs stay cell-sharded (dist_svd.py: only sketches and k x k Grams are all-reduced).
	world = 3: uneven shares of the per-bin polar problems (blocks of 32 / 26 / 6 ... rows over three ranks), one chromosome's
	CP-ALS per rank, 16-cell slabs."""
	import torch.multiprocessing as mp
	f, undo = fake_abi.install()
	try:
		core = _core(16, 12, [1000000], init_svd=init_svd)
		torch.manual_seed(0); np.random.seed(0)
		ds = load_small_dataset(good_qc_num=good, bs_cell=bs_cell)
		_, (A1, B1, D1, V1), _ = core.fit_transform(ds, size_ratio=0.3, n_iter_max=4, n_iter_parafac=1, do_conv=True, do_rwr=True,
		                                            do_col=do_col, tol=0.0, verbose=False)
		re1, n_i1 = list(core.re_trace), list(core.n_i)
	finally:
		undo()
	ctx = mp.get_context("spawn")
	q = ctx.Queue()
	port = _free_port()
	procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, q, do_col, good, bs_cell, init_svd)) for r in range(world)]
	for p in procs: p.start()
	res = sorted([q.get(timeout=180) for _ in procs], key=lambda r: r[0])
	for p in procs: p.join(timeout=60)
	for r in res:
		assert r[1] == n_i1
		assert np.allclose(r[2], re1, rtol=1e-6), (r[2], re1)
	for other in res[1:]:
		for a, b in zip(res[0][4] + res[0][5], other[4] + other[5]):
			assert np.array_equal(a, b)                                  # replicas stay bit-identical
		assert np.array_equal(res[0][3], other[3])
	V2 = res[0][3]
	assert V2.shape == tuple(V1.shape)
	E1 = O.embed_all(V1.numpy(), [d.numpy() for d in D1])
	E2 = O.embed_all(V2, res[0][5])
	for j in range(E1.shape[1]):
		assert abs(np.corrcoef(E1[:, j], E2[:, j])[0, 1]) > 0.9999


def _sharded_multires_worker(rank, world, port, q):
	import torch.distributed as dist
	os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
	dist.init_process_group("gloo", rank=rank, world_size=world)
	torch.set_num_threads(2)
	fake_abi.install()
	from fasthigashi_b200.sharding import shard_datasets, cell_slab
	from fasthigashi_b200.parafac2_intergrative import Fast_Higashi_core
	ds, g = load_multires_dataset()
	nchrom = len(g["chrom2size"])
	lo, hi = cell_slab(ds[0].num_cell, world, rank)
	st = _state(g, len(ds), nchrom)
	st = (st[0], st[1], st[2], st[3][lo:hi], [c[lo:hi] for c in st[4]], st[5], st[6])   # V rows and coverage rows of the slab
	core = Fast_Higashi_core(int(g["rank"]), int(g["off_diag"]), [int(r) for r in g["res"]], group=dist.group.WORLD).to("cpu")
	core.fit(shard_datasets(ds, world, rank), 0.3, 2, 1, True, True, False, 0.0, verbose=False, state=st)
	q.put((rank, list(core.re_trace), [core.projected_tensor_list[c].numpy() for c in core.chrom2size],
	       [a.numpy() for a in core.A_list] + [core.B_dict[c].numpy() for c in core.chrom2size] + [core.D_dict[c].numpy() for c in core.chrom2size]))
	dist.destroy_process_group()


def test_two_rank_gloo_multi_resolution_lockstep():
	"""The multi-resolution path cell-sharded over two gloo ranks (stacked bins of a chromosome's resolutions all-reduced as one
	projected tensor, one CP-ALS per chromosome on its owner rank, the stacked A rows split back and exchanged), in lock-step
	from the reference's state: the reference's losses <= 1e-4, its projected tensors, identical replicas."""
	import torch.multiprocessing as mp
	_, g = load_multires_dataset()
	ctx = mp.get_context("spawn")
	q = ctx.Queue()
	port = _free_port()
	procs = [ctx.Process(target=_sharded_multires_worker, args=(r, 2, port, q)) for r in range(2)]
	for p in procs: p.start()
	res = sorted([q.get(timeout=240) for _ in procs], key=lambda r: r[0])
	for p in procs: p.join(timeout=60)
	for r in res:
		assert np.max(np.abs(np.array(r[1]) - g["re"][:2]) / g["re"][:2]) < 1e-4, (r[1], g["re"][:2])
		for c, Y in enumerate(r[2]):
			assert rel_fro(Y, g["t1_Y%d" % c]) < 1e-3
	for a, b in zip(res[0][2] + res[0][3], res[1][2] + res[1][3]):
		assert np.array_equal(a, b)


def test_headline_job_script_runs_end_to_end(fake):
	"""scripts/headline_run.py (init + S sweeps + transform on per-rank synthetic slabs, the north star's full job) at a toy
	size: the JSON fields exist, the loss decreases, one RWR pass per sweep plus the one of transform's cache refresh."""
	import argparse
	sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
	import headline_run
	args = argparse.Namespace(cells_total=24, sweeps=3, geometry="pfc", rank=8, cache="sweep", init_svd="device")
	out = headline_run.run(args, "cpu", None, 0, 1, bins=[48, 36])
	assert out["cells_per_gpu"] == 24 and out["embedding_rows_local"] == 24 and len(out["rwr_steps"]) == 2
	assert out["re_last"][-1] < out["re_first"][1] and np.isfinite(out["job_s"])
	assert out["rwr_passes"] == 3


def _reference_datasets(**kw):
	from oracle import ref_shims
	from fasthigashi_b200 import synth
	if not ref_shims.reference_available():
		pytest.skip("the reference checkout is only present in the build container")
	mods = ref_shims.import_reference()
	# the generator call of tests/golden/make_golden.py: the tensors of data_small.npz
	chroms, _ = synth.synth_dataset([90, 70, 40], 48, 0.12, off_diag=12, seed=3, num_cluster=4)
	return ref_shims.build_reference_datasets(mods, chroms, off_diag=12, res=1000000, bs_bin=32, **kw)


def test_restaging_reference_chrom_datasets_is_exact():
	"""the reference's own `Chrom_Dataset` objects (pinned COO `Fake_Sparse` per bin block x cell
	batch, +1-offset int16 indices, good-QC batches then bad-QC batches) re-staged by `Chrom_Dataset.from_reference` give
	the same block-CSR, bit for bit, as building from the COO tensor directly."""
	from fasthigashi_b200.sparse_for_schic import Chrom_Dataset
	ref = _reference_datasets(bs_cell=20, good_qc_num=44)
	mine = load_small_dataset(good_qc_num=44, bs_cell=20)
	for r, m in zip(ref, mine):
		got = Chrom_Dataset.from_reference(r)
		assert (got.num_cell, got.total_cell_num, got.bs_bin, got.bs_cell, got.flank) == (44, 48, 32, 20, 12)
		assert [tuple(g) for g in got.geoms] == [tuple(g) for g in m.geoms]
		assert [(s.start, s.stop) for s in got.cell_slice_list] == [(s.start, s.stop) for s in m.cell_slice_list]
		for b in range(len(m.geoms)):
			assert torch.equal(got.rowptr[b], m.rowptr[b]) and torch.equal(got.col[b], m.col[b]) and torch.equal(got.val[b], m.val[b])


def test_core_accepts_the_reference_datasets(fake):
	"""The two-line patch of INTEGRATION.md section 2: `Fast_Higashi_core.fit_transform` fed with the REFERENCE's
	List[Chrom_Dataset] reproduces the reference's own run (core_col.npz) and writes `global_slice_bin` on its objects."""
	g = np.load(os.path.join(GOLDEN, "core_col.npz"))
	ref = _reference_datasets(bs_cell=int(g["bs_cell"]), good_qc_num=int(g["good_qc_num"]))
	core = _core(int(g["rank"]), 12, [1000000])
	torch.manual_seed(0); np.random.seed(0)
	nsweep = 4
	_, (A_list, B_list, D_list, V), proj = core.fit_transform(ref, size_ratio=0.3, n_iter_max=nsweep, n_iter_parafac=1, do_conv=True,
	                                                         do_rwr=True, do_col=True, tol=0.0, verbose=False)
	assert list(core.n_i) == list(g["n_i"])
	re = np.array(core.re_trace)
	assert np.max(np.abs(re - g["re"][:nsweep]) / g["re"][:nsweep]) < 1e-4
	assert tuple(V.shape) == (48, int(g["rank"])) and all(r.global_slice_bin == slice(0, r.num_bin) for r in ref)


def test_reference_wrapper_with_the_core_swapped_in(fake, tmp_path, monkeypatch):
	"""The literal two-line patch of INTEGRATION.md section 2 applied to the UNMODIFIED reference wrapper
	(`FastHigashi_Wrapper.Fast_Higashi_core = ours`): its own prep_dataset (reference ingest, reference Chrom_Datasets) ->
	run_model -> fetch_cell_embedding runs unchanged and gives the embeddings of the reference's own core (same seeds)."""
	import importlib
	import io
	import json
	import contextlib
	import pickle
	from oracle import ref_shims
	import wrapper_cases
	if not ref_shims.reference_available():
		pytest.skip("the reference checkout is only present in the build container")
	ref_shims.import_reference()
	W = importlib.import_module("fasthigashi.FastHigashi_Wrapper")
	from fasthigashi_b200.parafac2_intergrative import Fast_Higashi_core as Ours
	G = np.load(os.path.join(GOLDEN, "ingest_cases.npz"), allow_pickle=True)
	chroms = [str(c) for c in G["chroms"]]
	wrapper_cases.write_raw_files(G, tmp_path)
	pickle.dump({"batch": list(G["batch"])}, open(tmp_path / "label_info.pickle", "wb"))
	cfg = dict(chrom_list=chroms, temp_dir=str(tmp_path), data_dir=str(tmp_path), resolution=int(G["res"]), resolution_fh=[int(G["res"])])
	json.dump(cfg, open(tmp_path / "config.JSON", "w"))
	monkeypatch.chdir(tmp_path)                      # the reference constructor writes ./tmp1 ./tmp2 (:58-59)
	out = {}
	for who in ("reference", "ours"):
		(tmp_path / who).mkdir()
		if who == "ours":
			monkeypatch.setattr(W, "Fast_Higashi_core", Ours)
		with contextlib.redirect_stdout(io.StringIO()):
			w = W.FastHigashi(str(tmp_path / "config.JSON"), str(tmp_path / who), str(tmp_path / who), 12, True, True, True, False, False)
			w.prep_dataset()
			torch.manual_seed(0); np.random.seed(0)
			w.run_model(dim1=0.6, rank=8, n_iter_parafac=1, n_iter_max=4, tol=0.0)
			np.random.seed(1)
			out[who] = (w.fetch_cell_embedding(final_dim=4), [np.asarray(a) for a in w.A_list], w.meta_embedding)
		assert os.path.exists(tmp_path / who / ("results_all%s.pkl" % w.save_str))
	assert fake.calls["rwr_batched"] > 0 and fake.calls["cp_als"] > 0     # the second run went through the C-ABI entry points
	ea, eb = out["reference"][0]["embed_all"], out["ours"][0]["embed_all"]
	assert ea.shape == eb.shape
	for j in range(ea.shape[1]):
		assert abs(np.corrcoef(ea[:, j], eb[:, j])[0, 1]) > 0.999
	for a, b in zip(out["reference"][1], out["ours"][1]):
		assert a.shape == b.shape
	assert out["reference"][2].shape == out["ours"][2].shape


def test_chrom_dataset_fetch_matches_reference_fetch(fake):
	"""`Chrom_Dataset.fetch` / `fetch_bad` / `norm` (sparse_for_schic.py:585-620) against the dense blocks the reference's
	own fetch returned (tests/golden/rwr_cases.npz), bit for bit."""
	g = np.load(os.path.join(GOLDEN, "rwr_cases.npz"))
	ds = load_small_dataset(good_qc_num=44, bs_cell=20)
	for c in range(int(g["ncase"])):
		ci, b, cb, s, e = g["c%d_meta" % c]
		(x, t), kind = ds[ci].fetch(int(b), int(cb), save_context=dict(device="cpu"), transpose=True, do_conv=False)
		assert kind == "hic" and np.array_equal(x.numpy(), g["c%d_dense" % c])
		(xt, _), _ = ds[ci].fetch(int(b), int(cb), save_context=dict(device="cpu"))
		assert xt.shape == (x.shape[1], x.shape[2], x.shape[0]) and torch.equal(xt.permute(2, 0, 1), x)
	(xb, _), _ = ds[0].fetch_bad(1, 0, save_context=dict(device="cpu"), transpose=True)
	assert xb.shape[0] == 4 and torch.equal(xb, O.densify_block(ds[0], 1, 44, 48))
	d = np.load(os.path.join(GOLDEN, "data_small.npz"))
	good = d["chr1_idx"][2] < 44
	assert abs(ds[0].norm() - float(np.sqrt(np.square(d["chr1_val"][good].astype(np.float64)).sum()))) < 1e-6


def _dist_wrapper_worker(rank, world, port, q, tmp):
	import pathlib
	import torch.distributed as dist
	os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
	dist.init_process_group("gloo", rank=rank, world_size=world)
	torch.set_num_threads(2)
	fake_abi.install()
	G = np.load(os.path.join(GOLDEN, "ingest_cases.npz"), allow_pickle=True)
	chroms = [str(c) for c in G["chroms"]]
	w = _cpu_wrapper(pathlib.Path(tmp), 12, int(G["res"]), chroms).distribute(dist.group.WORLD)
	w.prep_dataset()
	torch.manual_seed(0); np.random.seed(0)
	w.run_model(dim1=0.6, rank=8, n_iter_parafac=1, n_iter_max=4, tol=0.0, init_svd="host")  # same start as the single process
	np.random.seed(1)
	emb = w.fetch_cell_embedding(final_dim=4)
	csr = [[(ds.rowptr[b].numpy(), ds.col[b].numpy(), ds.val[b].numpy()) for b in range(len(ds.geoms))] for ds in w.all_matrix]
	w.path2result_dir = tmp
	files = w.only_partial_rwr(out_format="npz")
	assert all(f.endswith("_rank%d.npz" % rank) for f in files)
	q.put((rank, w.reorder, w.good_qc_num, w.final_do_col, [(d.num_cell, d.total_cell_num, d.bs_bin, d.bs_cell) for d in w.all_matrix], csr,
	       emb["embed_all"], w.meta_embedding, os.path.exists(os.path.join(tmp, "results_all%s.pkl" % w.save_str))))
	dist.destroy_process_group()


def test_distributed_wrapper_on_two_gloo_ranks(fake, tmp_path):
	"""`FastHigashi.distribute(group)`: QC and ingest partitioned by chromosome (each raw file is read by ONE rank), the
	block-CSR of every chromosome scattered as cell slabs (== shard_datasets of the single-process datasets, bit for
	bit), sharded run_model with the embedding rows gathered in the unsharded order, files written by rank 0."""
	import pickle
	import torch.multiprocessing as mp
	import wrapper_cases
	from fasthigashi_b200.sharding import shard_datasets
	G = np.load(os.path.join(GOLDEN, "ingest_cases.npz"), allow_pickle=True)
	chroms = [str(c) for c in G["chroms"]]
	wrapper_cases.write_raw_files(G, tmp_path)
	single = _cpu_wrapper(tmp_path, 12, int(G["res"]), chroms)
	single.path2input_cache = single.path2result_dir = str(tmp_path / "single")
	os.makedirs(single.path2input_cache)
	single.prep_dataset()
	torch.manual_seed(0); np.random.seed(0)
	single.run_model(dim1=0.6, rank=8, n_iter_parafac=1, n_iter_max=4, tol=0.0)
	np.random.seed(1)
	emb1 = single.fetch_cell_embedding(final_dim=4)
	ctx = mp.get_context("spawn")
	q = ctx.Queue()
	port = _free_port()
	procs = [ctx.Process(target=_dist_wrapper_worker, args=(r, 2, port, q, str(tmp_path))) for r in range(2)]
	for p in procs: p.start()
	res = sorted([q.get(timeout=180) for _ in procs], key=lambda r: r[0])
	for p in procs: p.join(timeout=60)
	for r in res:
		rank = r[0]
		assert np.array_equal(r[1], single.reorder) and r[2] == single.good_qc_num and r[3] == single.final_do_col
		want = shard_datasets(single.all_matrix, 2, rank)
		assert r[4] == [(d.num_cell, d.total_cell_num, d.bs_bin, d.bs_cell) for d in want]
		for ds, got in zip(want, r[5]):
			for b in range(len(ds.geoms)):
				assert np.array_equal(got[b][0], ds.rowptr[b].numpy()) and np.array_equal(got[b][1], ds.col[b].numpy())
				assert np.array_equal(got[b][2], ds.val[b].numpy())
		assert r[6].shape == emb1["embed_all"].shape and r[7].shape == single.meta_embedding.shape
		for j in range(r[6].shape[1]):
			assert abs(np.corrcoef(r[6][:, j], emb1["embed_all"][:, j])[0, 1]) > 0.99
	assert np.array_equal(res[0][6], res[1][6])        # every rank holds the embeddings of ALL cells
	# only_partial_rwr in distributed mode: every rank wrote the maps of its own cells under their ORIGINAL ids
	single.path2result_dir = str(tmp_path / "single")
	ref_files = single.only_partial_rwr(out_format="npz")
	for ch, ref_file in zip(chroms, ref_files):
		ref = np.load(ref_file)
		parts = [np.load(os.path.join(str(tmp_path), "impute_prwr_%s_rank%d.npz" % (ch, r))) for r in range(2)]
		keys = [set(p.files) - {"shape"} for p in parts]
		assert not (keys[0] & keys[1]) and (keys[0] | keys[1]) == set(ref.files) - {"shape"}
		for p in parts:
			for k in set(p.files) - {"shape"}:
				# the auto-stop is per cell batch, so a slab may stop one step earlier or later than the full batch
				assert np.linalg.norm(p[k] - ref[k]) <= 0.05 * np.linalg.norm(ref[k])
	assert res[0][8]                                   # rank 0 wrote results_all*.pkl


def test_transform_with_other_flags_never_reuses_the_fitted_maps(fake):
	"""`transform(do_col=...)` with flags other than fit's (round-1 advisor item): the imputed maps of the last sweep and the
	kept Z = X^T V are dropped and every block is imputed again with the new flags - the embedding equals that of a core
	that was fitted to the same factors and transformed with those flags from scratch (cache='run' as well)."""
	g = np.load(os.path.join(GOLDEN, "core_col.npz"))
	good = int(g["good_qc_num"])
	out = {}
	for cache in ("sweep", "run"):
		core = _core(int(g["rank"]), 12, [1000000], cache=cache)
		ds = load_small_dataset(good_qc_num=good, bs_cell=int(g["bs_cell"]))
		core.fit(ds, 0.3, 2, 1, True, True, False, 0.0, verbose=False, state=_state(g, 3, 3))     # fitted with do_col=False
		fake.calls.clear()
		_, (_, _, _, V_same), _ = core.transform()                                                    # fit's flags
		same_calls = fake.calls["rwr_batched"]
		fake.calls.clear()
		_, (_, _, _, V_col), _ = core.transform(do_col=True)                                          # other flags: re-impute
		assert fake.calls["rwr_batched"] > same_calls or cache == "sweep"
		assert core._X_flags == (True, True, True) and not core._Z_valid
		out[cache] = (V_same.numpy().copy(), V_col.numpy().copy())
		assert rel_fro(out[cache][1], out[cache][0]) > 1e-3                                           # do_col changes the maps
	assert rel_fro(out["run"][1], out["sweep"][1]) < 1e-5 and rel_fro(out["run"][0], out["sweep"][0]) < 1e-5
	# from scratch: same factors, only the do_col transform
	core = _core(int(g["rank"]), 12, [1000000])
	ds = load_small_dataset(good_qc_num=good, bs_cell=int(g["bs_cell"]))
	core.fit(ds, 0.3, 2, 1, True, True, False, 0.0, verbose=False, state=_state(g, 3, 3))
	core.release()
	_, (_, _, _, V_ref), _ = core.transform(do_col=True)
	assert rel_fro(V_ref.numpy(), out["sweep"][1]) < 1e-5


def test_load_state_without_bad_cell_coverage_gives_them_a_zero_column_scale(fake):
	"""The reference stores 0 for "no coverage table of the bad-QC cells"; a do_col transform then reads rows of inf
	(column scale 1 / bin_cov = 0) for them instead of running past the end of the good cells' table (advisor item)."""
	g = np.load(os.path.join(GOLDEN, "core_col.npz"))
	good = int(g["good_qc_num"])
	assert good < 48
	st = list(_state(g, 3, 3))
	st[5] = [0, 0, 0]
	core = _core(int(g["rank"]), 12, [1000000])
	ds = load_small_dataset(good_qc_num=good, bs_cell=int(g["bs_cell"]))
	core.fit(ds, 0.3, 1, 1, True, True, True, 0.0, verbose=False, state=tuple(st))
	for ci, d in enumerate(core.schic):
		cov = core._cov_all[ci]
		assert tuple(cov.shape) == (d.total_cell_num, d.num_bin)
		assert torch.isinf(cov[d.num_cell:]).all() and torch.equal(cov[:d.num_cell], core.bin_cov_list[ci])
	_, (_, _, _, V), _ = core.transform()
	assert tuple(V.shape) == (48, int(g["rank"])) and torch.isfinite(V).all()


def test_inputs_consumed_hook_fires_behind_the_last_csr_read_of_a_sweep(fake, monkeypatch):
	"""`inputs_consumed_hook` (streaming input: bench.py's end-to-end leg uploads the next step's block-CSR from it): called
	once per sweep, after every RWR call of the sweep - the only readers of the block-CSR - and before the per-bin polar;
	phases C / P5 reuse the imputed X, so nothing reads the CSR after the hook (cache='sweep' and 'run')."""
	class _Event:
		def record(self, *a):
			self.recorded = True
	monkeypatch.setattr(torch.cuda, "Event", _Event)
	g = np.load(os.path.join(GOLDEN, "core_nocol.npz"))
	for cache in ("sweep", "run"):
		fake.calls.clear()
		core = _core(int(g["rank"]), 12, [1000000], cache=cache)
		core.prepare(load_small_dataset(), 0.3, True, True, False, state=_state(g, 3, 3))
		seen = []
		core.inputs_consumed_hook = lambda ev: seen.append((getattr(ev, "recorded", False), fake.calls.get("rwr_batched", 0),
		                                                    fake.calls.get("polar_isqrt_multi", 0)))
		after = []
		for t in range(3):
			core.sweep_once(1)
			after.append(fake.calls["rwr_batched"])
		assert len(seen) == 3 and all(s[0] for s in seen)
		assert [s[1] for s in seen] == after                      # no RWR call (no CSR read) after the hook within a sweep
		assert [s[2] for s in seen] == [0, 1, 2]                  # fired before the sweep's polar stage
		assert after[0] > 0 and (after[2] == 3 * after[0] if cache == "sweep" else after[2] == after[0])


def test_size_limits_are_checked_before_init(fake):
	"""The library's limits on the per-chromosome rank (per-bin polar: Gram side min(window, r) <= 160; inner CP-ALS: r <= 169)
	stop a run in _setup - before init_params' RWR passes - with a ValueError that names the chromosome and the way out; the
	reference has no such limit (rank 256, dim1 > 0.64 at 500 kb). Sizes inside the limits pass."""
	from types import SimpleNamespace as NS
	from fasthigashi_b200 import parafac2_intergrative as P
	core = _core(256, 12, [1000000])
	ds = load_small_dataset()
	with pytest.raises(ValueError, match="chr1.*r <= 169"):
		core._setup(ds, 0.3, [200, 200, 200])          # narrow windows (Gram side = window), but r beyond the CP-ALS limit
	core._setup(ds, 0.3, [165, 165, 165])              # Gram side = window width < 160, r <= 169: fine
	assert core.chrom2size["chr1"] == 165
	# a 500 kb chr1 block (115 rows, window 315) at dim1 = 0.7: r = int(499 * 0.7 * 0.5) = 174 -> Gram side 174
	core.schic = [NS(chrom="chr1", resolution=500000, geoms=[NS(w=215), NS(w=315)])]
	core.chrom2size = {"chr1": 161}
	with pytest.raises(ValueError, match="Gram side 161"):
		core._check_limits()
	core.chrom2size = {"chr1": P.MAX_POLAR_SIDE}
	core._check_limits()
