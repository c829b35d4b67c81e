"""bench.py on the CPU: the roofline numerators (`algorithmic_work`) against SURVEY.md 8d's table for BASELINE config 2, the
block rule, and the JSON line of the reference arm (`--impl reference`: host only, the product library never mapped, rank 0
alone under a multi-rank launch). The b200 arm itself needs a GPU (tests/test_gpu_parity.py, the driver's bench run)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def config2_geometry():
	import bench
	from fasthigashi_b200 import synth
	bins = synth.chrom_bins("pfc", bench.RES)
	return bench, bins, bench.make_datasets(4, 1000, "cpu", bins)


def test_block_rule_and_geometry_of_config2(config2_geometry):
	"""FastHigashi_Wrapper.py:501-512 at 500 kb: chr1 (457 valid bins) splits into 4 blocks of <= 115 rows with windows
	215 / 315 / 315 / <= 315; 5,432 bins = polar problems per sweep."""
	bench, bins, ds = config2_geometry
	assert sum(bins) == 5432 and len(bins) == 22
	assert bench.bs_bin_rule(457, 500000) == 115 and bench.bs_bin_rule(499, 500000) == 125 and bench.bs_bin_rule(2493, 100000) == 147
	g = ds[0].geoms
	assert [x.nb for x in g] == [115, 115, 115, 112] and [x.w for x in g][:3] == [215, 315, 315] and g[3].w <= 315
	assert sum(x.nb for d in ds for x in d.geoms) == 5432


def test_algorithmic_work_matches_the_survey_table(config2_geometry):
	"""dense panel 4.77 MB per cell, contraction flops 1.19e9 per cell and sweep, 5,432 polar problems,
	r in 21 ... 144 with sum 1,622; RWR bytes = block-CSR (6 B per entry + row pointers) + the panel written once."""
	bench, bins, ds = config2_geometry
	w = bench.algorithmic_work(ds, bench.RANK)
	C = ds[0].num_cell
	panel = sum(4 * g.nb * g.w for d in ds for g in d.geoms)
	assert abs(panel / 4.77e6 - 1) < 0.01
	assert abs(w["contraction_flops"] / C / 1.19e9 - 1) < 0.01
	assert w["polar_problems"] == 5432
	r = [min(int(n * bench.DIM1 * bench.RES / 1000000), bench.RANK) for n in bins]
	assert (min(r), max(r), sum(r)) == (21, 144, 1622)
	nnz = sum(int(v.numel()) for d in ds for v in d.val)
	rowptr = sum(C * (g.nb + 1) * 4 for d in ds for g in d.geoms)
	assert w["rwr_bytes"] == pytest.approx(nnz * 6 + rowptr + C * panel, rel=1e-12)
	# the generator's density: 0.05 of n^2 per cell, both triangles stored (SURVEY.md 8d "Synthetic inputs")
	assert abs(nnz / C / sum(n * n for n in bins) / bench.DENSITY - 1) < 0.1


def _run_reference_arm(extra_env, *args):
	code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference'] + %r; "
	        "runpy.run_path(%r, run_name='__main__'); "
	        "maps = open('/proc/self/maps').read(); "
	        "print('PRODUCT_LIB_MAPPED' if 'libfh_b200' in maps else 'NO_PRODUCT_LIB')") % (list(args), os.path.join(ROOT, "bench.py"))
	env = dict(os.environ, **extra_env)
	return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=ROOT, timeout=900)


def test_reference_arm_line_and_isolation():
	"""`bench.py --impl reference`: one JSON line with the contract's keys, value = MEASURED sample cells / seconds, e2e = the line's
	own value with zero copy bytes, `kind: port`, the core count stated - and the process never maps libfh_b200.so."""
	p = _run_reference_arm({"FH_REF_BUDGET_S": "100000"}, "--cpu-sample-cells", "4", "--steps", "1", "--warmup", "1", "--gpus", "1")
	assert p.returncode == 0, p.stderr[-2000:]
	lines = [l for l in p.stdout.splitlines() if l.strip()]
	assert lines[-1] == "NO_PRODUCT_LIB"
	d = json.loads(lines[-2])
	for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
	          "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
		assert k in d, k
	assert d["impl"] == "reference" and d["unit"] == "cells/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
	assert d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["data"] == "synthetic"
	assert "workload" in d["config"] and "model" not in d["config"] and d["config"]["cells_per_step"] == 4
	assert "2 passes per sweep" in d["config"]["rwr"]
	cb = d["cpu_baseline"]
	assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and "4-cell sample" in cb["sample"]
	assert d["value"] == pytest.approx(4 / (d["ms_per_step"] / 1e3), rel=1e-9)      # measured on the sample, not extrapolated
	assert d["e2e"] == {"value": d["value"], "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
	assert len(d["rwr_steps"]) == 22 and all(1 <= k <= 60 for k in d["rwr_steps"])


def test_reference_arm_other_ranks_exit_without_work():
	"""Under torchrun (N > 1) rank 0 alone runs the reference arm; the other ranks print nothing and exit 0."""
	p = _run_reference_arm({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2", "--steps", "1", "--warmup", "1")
	assert p.returncode == 0, p.stderr[-2000:]
	assert [l for l in p.stdout.splitlines() if l.strip()] == ["NO_PRODUCT_LIB"]


def test_b200_arm_fails_loudly_without_a_gpu():
	"""No CPU fallback: on a box without a CUDA device the product arm stops with an assertion instead of timing anything."""
	import torch
	if torch.cuda.is_available():
		pytest.skip("a CUDA device is present")
	p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
	                   cwd=ROOT, timeout=900)
	assert p.returncode != 0 and "needs a CUDA device (no CPU fallback)" in p.stderr
	assert not [l for l in p.stdout.splitlines() if l.startswith("{")]


def test_roofline_object_of_the_dominant_kernel(config2_geometry):
	"""`roofline` describes the kernel with the largest per-kernel total of a sweep: bound / achieved / peak / unit / frac = achieved / peak,
	`traffic` = that kernel's dram bytes per launch from the committed ncu capture (profiles/r02_ncu_traffic.json), the RWR stage reported
	against the HBM peak with its tensor side beside it. Per-stage and per-kernel times are the final record's."""
	bench, bins, ds = config2_geometry
	work = bench.algorithmic_work(ds, bench.RANK)
	pk = bench.peaks()
	per = {"rwr": 39.0, "p1_mttkrp": 3.2, "p3_project": 19.0, "p5_tensor": 17.7, "cp_als": 5.3, "polar_cells": 1.6}
	kt = {"densify_conv_kernel": {"ms_per_sweep": 8.7, "launches_per_sweep": 99}, "rwr_chain_kernel": {"ms_per_sweep": 28.6, "launches_per_sweep": 99},
	      "gemm_tc_kernel": {"ms_per_sweep": 39.0, "launches_per_sweep": 422}, "chol_jacobi_rb_kernel": {"ms_per_sweep": 34.9, "launches_per_sweep": 5}}
	roof, roof_all = bench.build_roofline(per, kt, work, ds, [4] * len(ds), pk, "cpu", "rwr")
	assert roof["dominant_kernel"] == "gemm_tc_kernel" and roof["bound"] == "tensor" and roof["unit"] == "TFLOP/s"
	assert roof["peak"] == pk["tensor"] and roof["frac"] == pytest.approx(roof["achieved"] / roof["peak"])
	assert roof["achieved"] == pytest.approx(work["contraction_flops"] / (39.9e-3) / 1e12)
	traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
	assert roof["traffic"] == traffic["gemm_tc_kernel"]["dram_bytes_per_launch"] and "ncu" in roof["traffic_source"]
	r = roof_all["rwr"]
	assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] == pk["hbm"]
	assert r["achieved"] == pytest.approx(work["rwr_bytes"] / 28.6e-3 / 1e9) and r["stage_achieved"] == pytest.approx(work["rwr_bytes"] / 39.0e-3 / 1e9)
	assert r["frac"] == pytest.approx(r["achieved"] / pk["hbm"]) and r["tensor_side"]["unit"] == "TFLOP/s"
	kt["rwr_chain_kernel"]["ms_per_sweep"] = 50.0
	roof, _ = bench.build_roofline(per, kt, work, ds, [4] * len(ds), pk, "cpu", "rwr")
	assert roof["dominant_kernel"] == "rwr_chain_kernel" and roof["bound"] == "hbm" and roof["traffic"] == traffic["rwr_chain_kernel"]["dram_bytes_per_launch"]
