"""bench.py - cells/s per PARAFAC2 ALS sweep (incl. RWR) on B200, with the reference's CPU path
timed beside it (BASELINE.json metric; contract in the task statement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--cells C]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one full ALS sweep of the hot path over the whole resident workload: RWR imputation of
every (cell, chromosome) from the block-CSR, the P1/P3/P5 contractions, the per-bin and cell-mode
polar factors, the inner CP-ALS of every chromosome and the loss terms.
Workload at N=1: BASELINE.json configs[1] (PFC-shaped: 4,238 cells, 22 autosomes at 500 kb, 5,432
bins, rank 256, off_diag 100, dim1 0.6, do_conv=do_rwr=True, do_col=False as the wrapper decides at
density 0.05). N>1: the same per-GPU slab on every rank (weak scaling), cells sharded.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RES, OFF_DIAG, RANK, DIM1, DENSITY = 500000, 100, 256, 0.6, 0.05


def bs_bin_rule(n, res):
	rec = min(max(int(15000000 / res), 128), 256)  # FastHigashi_Wrapper.py:501
	return math.ceil(n / max(math.ceil(n / rec), 1))  # :507,512


def make_datasets(ncell, seed, device, bins):
	import fasthigashi_b200  # noqa: F401
	from fasthigashi_b200 import synth
	from fasthigashi_b200.sparse_for_schic import Sparse, Chrom_Dataset
	rng = np.random.default_rng(seed)
	cluster = rng.integers(0, 8, size=ncell)
	out = []
	for ci, n in enumerate(bins):
		idx, val = synth.synth_chrom(n, ncell, DENSITY, OFF_DIAG, seed * 1000 + ci, cluster, 8, device=device, cell_chunk=256)
		sp = Sparse.__new__(Sparse)
		sp.indices, sp.values, sp.shape, sp.ndim, sp.indptr = idx, val, np.array([n, n, ncell]), 3, None
		out.append(Chrom_Dataset(sp, bs_bin=bs_bin_rule(n, RES), bs_cell=ncell, compact=True, flank=OFF_DIAG,
		                         chrom="chr%d" % (ci + 1), resolution=RES, device=device))
		del idx, val, sp
	return out


def random_state(datasets, rank_R, seed, device="cpu", n_i=None):
	"""Random-init factors of the right shapes (the timed sweep does not depend on their values)."""
	g = torch.Generator().manual_seed(seed)
	sizes = [min(int(ds.num_bin * DIM1 * ds.resolution / 1000000), rank_R) for ds in datasets]
	ncell = datasets[0].num_cell
	A = [torch.randn(ds.num_bin, r, generator=g) * 1e-2 + 1 for ds, r in zip(datasets, sizes)]
	B = [torch.eye(r) + 1e-2 * torch.randn(r, generator=g) for r in sizes]
	D = [torch.randn(rank_R, r, generator=g) * 0.05 for r in sizes]
	V = torch.randn(ncell, rank_R, generator=g)
	V = torch.linalg.qr(V)[0] if ncell >= rank_R else torch.linalg.qr(V.T)[0].T
	cov = [torch.ones(ds.total_cell_num, ds.num_bin) for ds in datasets]
	return (A, B, D, V.contiguous(), cov, [0] * len(datasets), n_i if n_i is not None else [4] * len(datasets))


def probe_rwr_steps(datasets, ncell_probe=128):
	"""The auto-stop step count of init_params (parafac2_intergrative.py:131-140,265) on a cell
	sample: every timed sweep then runs exactly that many RWR steps, as the reference does."""
	from fasthigashi_b200.partial_rwr import rwr_block_csr, pad4
	n_i = []
	for ds in datasets:
		worst = 0
		nc = min(ncell_probe, ds.num_cell)
		for b, g in enumerate(ds.geoms):
			ldw = pad4(g.w)
			x = torch.empty(nc, g.nb * ldw, dtype=torch.float32, device=ds.val[b].device)
			worst = max(worst, rwr_block_csr(ds, b, 0, nc, x, g.nb * ldw, -1, True, True, False))
		n_i.append(worst)
	return n_i


class ClockSampler:
	Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
	     "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

	def __init__(self, gpu_index):
		self.rows, self.proc, self.gpu = [], None, gpu_index

	def start(self):
		try:
			self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
			                              "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
			self.t = threading.Thread(target=self._read, daemon=True)
			self.t.start()
		except Exception:
			self.proc = None

	def _read(self):
		for line in self.proc.stdout:
			self.rows.append([x.strip() for x in line.split(",")])

	def stop(self):
		if self.proc is None:
			return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
		self.proc.terminate()
		try:
			self.proc.wait(timeout=2)
		except Exception:
			pass
		sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
		mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
		reasons = set()
		for r in self.rows:
			if len(r) >= 9:
				for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
					if v.lower().startswith("active"):
						reasons.add(name)
		return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
		        "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
	p = os.path.join(ROOT, "MEASURED_PEAKS.json")
	if os.path.exists(p):
		d = json.load(open(p))
		return dict(hbm=d["hbm_gbs"], tensor=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
	return dict(hbm=6650.0, tensor=1400.0, src="fallback")


def algorithmic_work(datasets, rank_R):
	"""SURVEY.md 8d formulas, summed over (cell, bin-block) for one sweep."""
	rwr_bytes = contraction_flops = polar_flops = 0.0
	problems = 0
	for ds in datasets:
		r = min(int(ds.num_bin * DIM1 * ds.resolution / 1000000), rank_R)
		C = ds.num_cell
		for b, g in enumerate(ds.geoms):
			nnz = ds.val[b].numel() * (C / ds.total_cell_num)
			rwr_bytes += nnz * 6 + C * ((g.nb + 1) * 4 + 4 * g.nb * g.w)
			contraction_flops += C * (6.0 * g.nb * g.w * r + 4.0 * g.nb * r * rank_R)
			polar_flops += g.nb * (4.0 * g.w * r * r + 10.0 * r ** 3)
			problems += g.nb
	return dict(rwr_bytes=rwr_bytes, contraction_flops=contraction_flops, polar_flops=polar_flops, polar_problems=problems)


# --------------------------------------------------------------------------------------------
def oracle_rwr_steps(cpu_ds, sample=8):
	"""The auto-stop step count of init_params (parafac2_intergrative.py:131-140,265) evaluated by the ORACLE on a few cells
	per block (host only: the reference arm must not touch the product's CUDA library)."""
	from oracle import fh_oracle as O
	n_i = []
	for ds in cpu_ds:
		worst = 0
		nc = min(sample, ds.num_cell)
		for b, g in enumerate(ds.geoms):
			worst = max(worst, O.partial_rwr(O.densify_block(ds, b, 0, nc), g.s, g.e, True, True, False, None, -1)[1])
		n_i.append(worst)
	return n_i


def time_oracle(datasets, rank_R, n_i, sample_cells, steps, warmup, device="cpu"):
	"""The reference algorithm (oracle/fh_oracle.py = stock-torch restatement pinned to the reference's outputs; the reference
	itself is not on the GPU box) on a cell sample of the same geometry. One step = one iteration of the reference's outer loop
	(parafac2_intergrative.py:635-737): projections with TWO RWR passes (one cell batch, as the reference's batching rule gives
	on a 180 GB device, FastHigashi_Wrapper.py:500-517), V update, projected tensor, CP-ALS per chromosome, core norms.
	device="cpu": host cores (cpu_baseline / --impl reference); device="cuda": the same stock torch ops (cuBLAS / cuSOLVER) on
	the GPU - the same-box stock-PyTorch baseline of BASELINE.md 3.5. Returns measured seconds per step."""
	from oracle import fh_oracle as O
	ncpu = os.cpu_count() or 1
	if device == "cpu":
		torch.set_num_threads(ncpu)
	sample_cells = min(sample_cells, datasets[0].num_cell)
	ods = [ds.select_cells(0, sample_cells).to(device) for ds in datasets]
	core = O.OracleCore(rank_R, OFF_DIAG, [RES], device=device)
	state = random_state(ods, rank_R, 7, n_i=n_i)
	core.set_sizes(ods, DIM1)
	core.load_state(*state)
	fixed = [0.0]
	orig_polar, orig_cp = O.polar, O.cp_als
	sync = torch.cuda.synchronize if device != "cpu" else (lambda: None)

	def timed(fn):
		def w(*a, **k):
			sync()
			t = time.perf_counter()
			r = fn(*a, **k)
			sync()
			fixed[0] += time.perf_counter() - t
			return r
		return w
	O.polar, O.cp_als = timed(orig_polar), timed(orig_cp)
	rec = []
	try:
		for it in range(warmup + steps):
			fixed[0] = 0.0
			sync()
			t = time.perf_counter()
			core.sweep(ods, True, True, False, want_norm=(it == 0))
			for ci, ds in enumerate(ods):
				fac, _, _ = O.cp_als(core.projected[ds.chrom], [core.A_list[ci], core.B_dict[ds.chrom], core.D_dict[ds.chrom]], 1)
				core.A_list[ci], core.B_dict[ds.chrom], core.D_dict[ds.chrom] = fac
			core.core_norms(ods)
			sync()
			total = time.perf_counter() - t
			if it >= warmup:
				rec.append((total, fixed[0]))
	finally:
		O.polar, O.cp_als = orig_polar, orig_cp
	tot = [r[0] for r in rec]
	fx = float(np.median([r[1] for r in rec]))
	med = float(np.median(tot))
	return dict(total=med, sum=float(np.sum(tot)), fixed=fx, per_cell=(med - fx) / sample_cells, cores=ncpu, sample_cells=sample_cells)


def baseline_record(t, full_cells, kind, where):
	"""cells/s MEASURED on the sample (value) + the labelled linear extrapolation to the full cell count (not the headline)."""
	sec_full = t["fixed"] + full_cells * t["per_cell"]
	return {"value": t["sample_cells"] / t["total"], "unit": "cells/s", "cores": t["cores"] if where == "cpu" else 0, "kind": kind,
	        "sample": "%s: %d-cell sample of the same geometry (all 22 chromosomes, every bin block), measured %.2f s per sweep "
	                  "(2 RWR passes per sweep as the reference; cell-independent part - per-bin polar + CP-ALS - %.2f s of it)" % (
		                  where, t["sample_cells"], t["total"], t["fixed"]),
	        "seconds_per_sweep_sample": t["total"],
	        "extrapolated_full_workload": {"cells": full_cells, "seconds_per_sweep": sec_full, "cells_per_s": full_cells / sec_full,
	                                       "model": "fixed + cells x per_cell (cost linear in cells except the cell-independent part); "
	                                                "NOT the value above"}}


def build_roofline(per, kt, work, datasets, n_i, pk, dev, top):
	"""`roofline` of the dominant kernel + `roofline_all` for the three stage groups SURVEY.md 8d names.
	RWR: HBM bound per 8d (algorithmic bytes = block-CSR read + imputed panel written once), the tensor-pipe figure beside it
	(the stage is compute bound at fp32 parity: ~170 flop per algorithmic byte, three binary16 MMAs per product).
	Contractions: tensor pipe (useful fp32-equivalent flops counted once; 3xTF32 executes 3x that).
	Per-bin polar: fp64 pipe, against a DGEMM peak measured in this run.
	achieved = algorithmic work per launch / average launch duration from CUDA events on the launching stream (kt);
	traffic = dram__bytes_read + write per launch of that kernel from the committed ncu --set full capture (profiles/), or null."""
	roof_all = {}
	rwr_flops = 0.0
	for ds, k in zip(datasets, n_i):
		for g_ in ds.geoms:
			rwr_flops += ds.num_cell * (4.0 * g_.nb * g_.nb * g_.w + 2.0 * max(k - 1, 0) * g_.nb ** 3)
	traffic = {}
	tp = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
	if os.path.exists(tp):
		traffic = json.load(open(tp))

	def per_launch(name, work_total, unit_scale):
		k = kt.get(name)
		if not k or not k["launches_per_sweep"] or k["ms_per_sweep"] <= 0:
			return None
		return work_total / k["launches_per_sweep"] / (k["ms_per_sweep"] / k["launches_per_sweep"] / 1e3) / unit_scale

	if "rwr" in per:
		stage = work["rwr_bytes"] / (per["rwr"] / 1e3) / 1e9
		chain = per_launch("rwr_chain_kernel", work["rwr_bytes"], 1e9)
		t_stage = rwr_flops / (per["rwr"] / 1e3) / 1e12
		t_chain = per_launch("rwr_chain_kernel", rwr_flops, 1e12)
		roof_all["rwr"] = {"bound": "hbm", "kernel": "rwr_chain16_kernel (tcgen05 kind::f16, 3xFP16 operand split, TMA-fed: A A^T, transition matrix, RWR steps with Q in TMEM, Q A)",
		                   "achieved": chain if chain is not None else stage, "peak": pk["hbm"], "unit": "GB/s",
		                   "frac": (chain if chain is not None else stage) / pk["hbm"], "stage_achieved": stage, "stage_frac": stage / pk["hbm"],
		                   "stage": "densify_conv_kernel + rwr_chain_kernel",
		                   "tensor_side": {"achieved": t_chain if t_chain is not None else t_stage, "stage_achieved": t_stage, "peak": pk["tensor"],
		                                   "unit": "TFLOP/s", "frac": (t_chain if t_chain is not None else t_stage) / pk["tensor"],
		                                   "frac_of_3xfp16_ceiling": (t_chain if t_chain is not None else t_stage) / (pk["tensor"] / 3.0),
		                                   "note": "algorithmic fp32 flops (2 nb^2 w for A A^T, 2 nb^3 per step, 2 nb^2 w for Q A) against the measured dense "
		                                           "bf16 peak; fp32-parity maths executes three binary16 MMAs (hi hi + hi lo + lo hi) per product on 128-padded tiles "
		                                           "(~3.5x the algorithmic flops): the stage is bound by the SM (tensor pipe, shared-memory operand bandwidth, the "
		                                           "drain warps' fp32 work between the products), not by HBM"}}
	gem = sum(per.get(k, 0.0) for k in ("p1_mttkrp", "p3_project", "p5_tensor"))
	if gem > 0:
		a = work["contraction_flops"] / (gem / 1e3) / 1e12
		ak = per_launch("gemm_tc_kernel", work["contraction_flops"], 1e12)
		roof_all["contractions"] = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05 3xTF32, TMA, A operand in TMEM)", "achieved": a, "peak": pk["tensor"],
		                            "unit": "TFLOP/s", "frac": a / pk["tensor"], "frac_of_3xtf32_ceiling": a / (pk["tensor"] / 6.0), "kernel_only_achieved": ak,
		                            "note": "useful fp32-equivalent flops over the P1+P3+P5 stage times; fp32-parity maths: 3xTF32 ceiling is ~peak_tf32/3 = ~peak_bf16/6. "
		                                    "kernel_only_achieved divides by the gemm_tc_kernel launches alone (they also serve the small per-bin products)"}
	if "polar_bins" in per:
		# measured fp64 peak of this GPU: cuBLAS DGEMM (library call used ONLY as the yardstick, never on the path)
		a64 = torch.randn(4096, 4096, device=dev, dtype=torch.float64)
		torch.matmul(a64, a64)
		torch.cuda.synchronize()
		g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		g0.record()
		for _ in range(3):
			torch.matmul(a64, a64)
		g1.record()
		torch.cuda.synchronize()
		pk64 = 3 * 2.0 * 4096 ** 3 / (g0.elapsed_time(g1) / 1e3) / 1e12
		del a64
		a = work["polar_flops"] / (per["polar_bins"] / 1e3) / 1e12
		roof_all["polar_bins"] = {"bound": "fp64", "kernel": "chol_jacobi_rb_kernel (pivoted Cholesky + register-blocked one-sided Jacobi) + fp64 Gram / factor GEMMs",
		                          "achieved": a, "unit": "TFLOP/s fp64", "peak": pk64, "frac": a / pk64, "peak_source": "cuBLAS DGEMM 4096^3 timed in this run",
		                          "problems_per_s": work["polar_problems"] / (per["polar_bins"] / 1e3),
		                          "jacobi_kernel_ms_per_sweep": kt.get("chol_jacobi_rb_kernel", {}).get("ms_per_sweep")}
	# the dominant kernel: the largest of the per-kernel totals of one sweep
	kname = max(kt, key=lambda k: kt[k]["ms_per_sweep"]) if kt else "rwr_chain_kernel"
	dom = {"rwr_chain_kernel": "rwr", "densify_conv_kernel": "rwr", "gemm_tc_kernel": "contractions", "chol_jacobi_rb_kernel": "polar_bins"}[kname]
	roofline = dict(roof_all.get(dom, {}))
	roofline["dominant_kernel"] = kname
	roofline["dominant_stage"] = top
	tr = traffic.get(kname)
	roofline["traffic"] = tr["dram_bytes_per_launch"] if tr else None
	roofline["traffic_source"] = tr["source"] if tr else None
	roofline["peak_source"] = pk["src"]
	return roofline, roof_all


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument("--gpus", type=int, default=1)
	ap.add_argument("--steps", type=int, default=5)
	ap.add_argument("--warmup", type=int, default=3)
	ap.add_argument("--impl", default="b200")
	ap.add_argument("--cells", type=int, default=4238)
	ap.add_argument("--geometry", default="pfc", help="pfc (configs[1]) | hg19")
	ap.add_argument("--cache", default="sweep", help="sweep: RWR recomputed every sweep (metric); run: once per run")
	ap.add_argument("--cpu-sample-cells", type=int, default=0, help="cells of the CPU baseline sample (0: 512 for --impl reference, 64 for the cpu_baseline leg)")
	ap.add_argument("--torch-gpu-sample-cells", type=int, default=-1, help="cells of the same-box stock-PyTorch (cuBLAS/cuSOLVER) baseline; -1: the full workload, 0: skip")
	ap.add_argument("--no-cpu-baseline", action="store_true")
	ap.add_argument("--no-e2e", action="store_true")
	ap.add_argument("--tc", type=int, default=-1, help="1: tcgen05 3xTF32 GEMMs, 0: CUDA-core fp32 (default: library default)")
	args = ap.parse_args()

	world = int(os.environ.get("WORLD_SIZE", "1"))
	rank = int(os.environ.get("RANK", "0"))
	local_rank = int(os.environ.get("LOCAL_RANK", "0"))
	import fasthigashi_b200  # noqa: F401
	from fasthigashi_b200 import synth
	bins = synth.chrom_bins("pfc", RES) if args.geometry == "pfc" else synth.chrom_bins("hg19", RES)
	workload = "%s-shaped synthetic scHi-C: %d cells/GPU, 22 autosomes @500kb (%d bins), rank %d, off_diag %d, dim1 %.1f, density %.2f" % (
		"PFC" if args.geometry == "pfc" else "hg19", args.cells, sum(bins), RANK, OFF_DIAG, DIM1, DENSITY)
	config = {"workload": workload, "cells_per_gpu": args.cells, "rwr": "recomputed once per sweep" if args.cache == "sweep" else "cached per run",
	          "l2": "inputs (>= 20 GB imputed tensor per sweep) far exceed the 126 MB L2", "init": "random factors (timing does not depend on values)",
	          "sharding": "cells x%d" % world}

	if args.impl == "reference":
		# The reference's own algorithm on the box's host cores: the pinned stock-torch port (the reference is pure Python and
		# cannot travel; oracle/fh_oracle.py restates it and is pinned to its outputs). Host only - this arm neither imports
		# __graft_entry__ / _lib nor maps libfh_b200.so. Rank 0 alone; every step is one MEASURED sweep over a bounded cell
		# sample (>= 512 cells) of the same geometry; value = sample cells / measured seconds (no extrapolation).
		if rank != 0:
			return
		sample = args.cpu_sample_cells or 512
		ds = make_datasets(sample, 1000, "cpu", bins)
		n_i = oracle_rwr_steps(ds)
		# wall-clock guard ("the whole --steps K --warmup W run ends within a few minutes"): one probe sweep on the full sample;
		# if K + W such sweeps would not fit FH_REF_BUDGET_S (default 240 s) the sample is cut (never below 128 cells) and said so
		budget = float(os.environ.get("FH_REF_BUDGET_S", "240"))
		probe = time_oracle(ds, RANK, n_i, sample, 1, 0, "cpu")
		n_sweeps = args.steps + max(args.warmup - 1, 0)  # the probe is the first warm-up sweep
		if probe["total"] * n_sweeps > budget:
			per_cell = max(probe["per_cell"], 1e-9)
			fit = int((budget / n_sweeps - probe["fixed"]) / per_cell)
			sample = int(min(sample, max(128, fit)))
		t = time_oracle(ds, RANK, n_i, sample, args.steps, max(args.warmup - 1, 0) if sample == probe["sample_cells"] else max(args.warmup, 1), "cpu")
		cb = baseline_record(t, args.cells, "port", "cpu")
		config = dict(config, rwr="2 passes per sweep (the reference re-imputes for the projected tensor)", cells_per_step=sample)
		print(json.dumps({"impl": "reference", "metric": "cells/s per PARAFAC2 ALS sweep (incl. RWR)", "value": cb["value"], "unit": "cells/s",
		                  "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
		                  "ms_per_step": t["total"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
		                  "data": "synthetic", "config": config, "cpu_baseline": cb, "rwr_steps": n_i,
		                  "e2e": {"value": cb["value"], "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
		return

	# ---------------------------------------------------------------------------- b200 arm
	import __graft_entry__ as ge
	if rank == 0:
		ge.build()
	assert torch.cuda.is_available(), "bench.py --impl b200 needs a CUDA device (no CPU fallback)"
	torch.cuda.set_device(local_rank)
	dev = torch.device("cuda", local_rank)
	group = None
	if world > 1:
		import torch.distributed as dist
		dist.init_process_group("nccl", device_id=dev)
		group = dist.group.WORLD
		dist.barrier()
	from fasthigashi_b200 import _lib
	from fasthigashi_b200.parafac2_intergrative import Fast_Higashi_core
	_lib.lib()
	datasets = make_datasets(args.cells, 1000 + rank, dev, bins)
	n_i = probe_rwr_steps(datasets)
	if world > 1:
		t = torch.tensor(n_i, device=dev)
		dist.all_reduce(t, op=dist.ReduceOp.MAX)
		n_i = t.tolist()
	state = random_state(datasets, RANK, 7, n_i=n_i)
	if world > 1:  # identical replicated factors, rank-local V rows
		g = torch.Generator().manual_seed(100 + rank)
		state = state[:3] + (torch.linalg.qr(torch.randn(args.cells, RANK, generator=g))[0].contiguous(),) + state[4:]
	core = Fast_Higashi_core(RANK, OFF_DIAG, [RES], cache=args.cache, group=group,
	                         use_tc=None if args.tc < 0 else bool(args.tc)).to(dev)
	core.verbose = False
	core.prepare(datasets, DIM1, True, True, False, state=state)
	work = algorithmic_work(datasets, RANK)

	def sync_all():
		if world > 1:
			dist.barrier()
		torch.cuda.synchronize()

	for _ in range(args.warmup):
		core.sweep_once(1)
	sync_all()
	core.enable_timers(True)
	sampler = ClockSampler(local_rank)
	sampler.start()
	l0 = _lib.launch_count()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	sync_all()
	e0.record()
	for _ in range(args.steps):
		core.sweep_once(1)
	e1.record()
	sync_all()
	ms = e0.elapsed_time(e1)
	launches = _lib.launch_count() - l0
	clocks = sampler.stop()
	stages = core.collect_timers()
	core.enable_timers(False)
	if world > 1:
		t = torch.tensor([ms], device=dev, dtype=torch.float64)
		dist.all_reduce(t, op=dist.ReduceOp.MAX)
		ms = float(t.item())
	ms_per_step = ms / args.steps
	total_cells = args.cells * world
	value = total_cells / (ms_per_step / 1e3)

	# ---- end-to-end: block-CSR in pinned host memory -> H2D -> sweep -> loss D2H, every step
	e2e = None
	if not args.no_e2e:
		host, h2d = [], 0
		for ds in datasets:
			hs = []
			for arrs in (ds.rowptr, ds.col, ds.val):
				hs.append([a.cpu().pin_memory() for a in arrs])
				h2d += sum(a.numel() * a.element_size() for a in arrs)
			host.append(hs)
		n_e2e = max(2, min(args.steps, 10))
		copy_stream = torch.cuda.Stream(device=dev)

		def upload(after):
			"""One full upload of the block-CSR from pinned host memory on the copy stream, chromosome by chromosome (the
			sweep waits per chromosome, so its RWR pass starts on chr1 while the later chromosomes are still in flight).
			`after`: the event behind which the device arrays may be overwritten (None: everything on the compute stream)."""
			events = {}
			if after is None:
				copy_stream.wait_stream(torch.cuda.current_stream())
			else:
				copy_stream.wait_event(after)
			with torch.cuda.stream(copy_stream):
				for ci, (ds, hs) in enumerate(zip(datasets, host)):
					for dst, src in zip((ds.rowptr, ds.col, ds.val), hs):
						for d, s_ in zip(dst, src):
							d.copy_(s_, non_blocking=True)
					events[ci] = torch.cuda.Event()
					events[ci].record(copy_stream)
			return events

		# Every step uploads the whole block-CSR (n_e2e uploads inside the timed region, none before it). The upload of step
		# i + 1 is issued from the core's `inputs_consumed_hook`, i.e. as soon as step i's RWR pass - the only reader of the
		# block-CSR in a sweep - has been launched, behind an event on the compute stream: it runs under the polar / projection
		# / CP-ALS stages of step i instead of waiting for step i's final read-back.
		pipe = {"left": n_e2e - 1, "next": None}

		def on_consumed(ev):
			if pipe["left"] > 0:
				pipe["left"] -= 1
				pipe["next"] = upload(ev)

		sync_all()
		e0.record()
		nxt = upload(None)
		core.inputs_consumed_hook = on_consumed
		for _ in range(n_e2e):
			core.input_events = nxt
			pipe["next"] = None
			core.sweep_once(1)  # ends with the D2H read of the loss terms
			nxt = pipe["next"]
		core.inputs_consumed_hook = None
		core.input_events = None
		e1.record()
		sync_all()
		ms_e = e0.elapsed_time(e1) / n_e2e
		if world > 1:
			t = torch.tensor([ms_e], device=dev, dtype=torch.float64)
			dist.all_reduce(t, op=dist.ReduceOp.MAX)
			ms_e = float(t.item())
		e2e = {"value": total_cells / (ms_e / 1e3), "unit": "cells/s", "h2d_bytes_per_step": int(h2d),
		       "d2h_bytes_per_step": int((2 * len(datasets) + 1) * 8 + len(datasets) * 8), "ms_per_step": ms_e, "steps": n_e2e,
		       "pipeline": "one upload of the whole block-CSR per step, all inside the timed region; the upload of step i+1 starts when "
		                   "step i's RWR pass has consumed the block-CSR (Fast_Higashi_core.inputs_consumed_hook)"}
		del host

	# one more sweep with the library's per-kernel CUDA-event timing on (events on the launching stream around every
	# launch of the four hot kernels): average launch durations for the roofline, outside the timed region above
	kt = {}
	_lib.kernel_timing(True)
	core.sweep_once(1)
	torch.cuda.synchronize()
	for name, kind in (("densify_conv_kernel", _lib.TIME_DENSIFY), ("rwr_chain_kernel", _lib.TIME_RWR_CHAIN),
	                   ("gemm_tc_kernel", _lib.TIME_GEMM_TC), ("chol_jacobi_rb_kernel", _lib.TIME_POLAR_JACOBI)):
		ms_k, n_k = _lib.kernel_time(kind)
		kt[name] = {"ms_per_sweep": ms_k, "launches_per_sweep": n_k}
	_lib.kernel_timing(False)
	if world > 1:
		dist.barrier()
		dist.destroy_process_group()
	if rank != 0:
		return
	pk = peaks()
	per = {k: v / args.steps for k, v in stages.items()}
	top = max(per, key=per.get) if per else None
	roofline, roof_all = build_roofline(per, kt, work, datasets, n_i, pk, dev, top)
	out = {"metric": "cells/s per PARAFAC2 ALS sweep (incl. RWR)", "value": value, "unit": "cells/s", "n_gpus": world,
	       "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
	       "vs_baseline": None, "dtype": "f32 (fp64 inside the polar step)", "data": "synthetic", "config": config,
	       "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline,
	       "stages_ms_per_sweep": per, "kernels_ms_per_sweep": kt, "roofline_all": roof_all, "rwr_steps": n_i, "re_trace_tail": core.re_trace[-3:]}
	if not args.no_cpu_baseline and world == 1:
		# reported baselines, not targets: (1) the reference algorithm on the host cores, (2) the same stock torch ops on this
		# GPU (cuBLAS / cuSOLVER driven as the reference drives them: one cell batch on a 180 GB device, gesvda polar) - the
		# same-box number BASELINE.md 3.5 names. Both on bounded cell samples of the same geometry, measured, not extrapolated.
		core.release()
		torch.cuda.empty_cache()
		t = time_oracle(datasets, RANK, n_i, args.cpu_sample_cells or 256, 1, 0, "cpu")
		out["cpu_baseline"] = baseline_record(t, args.cells, "port", "cpu")
		if args.torch_gpu_sample_cells != 0:
			try:
				t = time_oracle(datasets, RANK, n_i, args.cells if args.torch_gpu_sample_cells < 0 else args.torch_gpu_sample_cells, 1, 1, str(dev))
				out["torch_gpu_baseline"] = baseline_record(t, args.cells, "port on cuda (stock torch: cuBLAS bmm/einsum, cuSOLVER gesvda/getrf)", "cuda")
			except Exception as e:  # a baseline must not take the product's line down
				out["torch_gpu_baseline"] = {"unavailable": repr(e)[:200]}
	print(json.dumps(out))


if __name__ == "__main__":
	main()
