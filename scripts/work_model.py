"""Algorithmic work of one ALS sweep for every BASELINE config, from the geometry alone (SURVEY.md 8d formulas, the ones
`bench.py` uses), and the sweep time each implies at the per-unit rates MEASURED on one B200 in round 1
(profiles/r01_bench_n1_final.json: PFC-shaped config 2). A planning aid: everything below the measured row is a
projection and is labelled as one.

  python scripts/work_model.py            # prints a markdown table + one JSON line per config
"""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fasthigashi_b200  # noqa: E402,F401
from fasthigashi_b200 import synth  # noqa: E402
from fasthigashi_b200.sparse_for_schic import block_geometry  # noqa: E402

OFF_DIAG, DIM1 = 100, 0.6
# (name, geometry, resolution, cells, rank R, GPUs, density of stored entries, whole-chromosome blocks?)
CONFIGS = [("1 PR1 ref (CPU rule: whole-chromosome blocks)", "hg19", 1000000, 500, 64, 1, 0.10, True),
           ("2 PFC-shaped", "pfc", 500000, 4238, 256, 1, 0.05, False),
           ("3 20k cells", "hg19", 500000, 20000, 256, 8, 0.05, False),
           ("4 100k cells (headline)", "hg19", 500000, 100000, 256, 8, 0.05, False),
           ("5 10k cells @100 kb", "hg19", 100000, 10000, 256, 8, 0.004, False)]


def bs_bin_rule(n, res, whole):
	if whole:
		return n                                                        # FastHigashi_Wrapper.py:508-510 (CPU boxes)
	rec = min(max(int(15000000 / res), 128), 256)                       # :501
	return math.ceil(n / max(math.ceil(n / rec), 1))                    # :507,512


def work(kind, res, R, density, whole):
	"""Per-cell and per-sweep work of one sweep (one RWR pass): dict of sums over all (chromosome, bin block)."""
	out = dict(bins=0, blocks=0, units_per_cell=0, panel_bytes=0.0, rwr_bytes=0.0, rwr_dense_flops=0.0, contraction_flops=0.0,
	           polar_problems=0, polar_flops=0.0, t1_bytes=0.0, y_bytes=0.0, r_sum=0, max_nb=0, max_gram=0)
	for n in synth.chrom_bins(kind, res):
		r = min(int(n * DIM1 * res / 1000000), R)
		out["bins"] += n; out["r_sum"] += r
		for g in block_geometry(n, bs_bin_rule(n, res, whole), OFF_DIAG, True):
			nnz = density * n * n * (g.nb / n)                              # stored entries of the block, per cell
			out["blocks"] += 1; out["units_per_cell"] += 1
			out["panel_bytes"] += 4.0 * g.nb * g.w
			out["rwr_bytes"] += nnz * 6 + (g.nb + 1) * 4 + 4.0 * g.nb * g.w
			out["rwr_dense_flops"] += 4.0 * g.nb * g.nb * g.w + 2.0 * 4 * g.nb ** 3      # k = 4
			out["contraction_flops"] += 6.0 * g.nb * g.w * r + 4.0 * g.nb * r * R
			out["polar_problems"] += g.nb
			out["polar_flops"] += g.nb * (4.0 * g.w * r * r + 10.0 * min(g.w, r) ** 3)
			out["t1_bytes"] += 4.0 * g.nb * g.w * r
			out["max_nb"] = max(out["max_nb"], g.nb); out["max_gram"] = max(out["max_gram"], min(g.w, r))
		out["y_bytes"] += 4.0 * n * r * R
	return out


def main():
	meas = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_n1_final.json")))
	st = meas["stages_ms_per_sweep"]
	ref = work("pfc", 500000, 256, 0.05, False)
	cells_ref = 4238
	# measured rates of config 2 on one B200 (per unit of the work each stage scales with)
	rate = dict(rwr_s_per_dense_flop=st["rwr"] * 1e-3 / (ref["rwr_dense_flops"] * cells_ref),
	            contr_s_per_flop=(st["p1_mttkrp"] + st["p3_project"] + st["p5_tensor"]) * 1e-3 / (ref["contraction_flops"] * cells_ref),
	            polar_s_per_flop=st["polar_bins"] * 1e-3 / ref["polar_flops"],
	            fixed_s=(st["cp_als"] + st["polar_cells"]) * 1e-3)
	print("| config | GPUs | cells/GPU | bins | blocks | polar problems | panel MB/cell | resident GB/GPU | RWR GF/cell (dense-equiv.) | "
	      "contraction GF/cell | all-reduce MB/sweep (T1 + Y) | projected ms/sweep | projected cells/s | fused RWR kernel? |")
	print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
	for name, kind, res, cells, R, gpus, density, whole in CONFIGS:
		w = work(kind, res, R, density, whole)
		cpg = math.ceil(cells / gpus)
		t_rwr = rate["rwr_s_per_dense_flop"] * w["rwr_dense_flops"] * cpg
		t_con = rate["contr_s_per_flop"] * w["contraction_flops"] * cpg
		t_pol = rate["polar_s_per_flop"] * w["polar_flops"] / gpus          # problems partitioned over the ranks
		t_fix = rate["fixed_s"] * (w["bins"] / ref["bins"])
		t_ar = (w["t1_bytes"] + w["y_bytes"]) * 2 / 600e9 if gpus > 1 else 0.0    # ring-equivalent at ~600 GB/s effective
		total = t_rwr + t_con + t_pol + t_fix + t_ar
		fused = "yes" if w["max_nb"] <= 128 and not (density * (500000 / res) ** 2 <= 0.03) else \
			("no: nb %d > 128" % w["max_nb"] if w["max_nb"] > 128 else "no: do_col")
		row = dict(config=name, gpus=gpus, cells_per_gpu=cpg, bins=w["bins"], blocks=w["blocks"], polar_problems=w["polar_problems"],
		           panel_mb_per_cell=w["panel_bytes"] / 1e6, resident_gb_per_gpu=w["panel_bytes"] * cpg / 1e9,
		           rwr_dense_gflop_per_cell=w["rwr_dense_flops"] / 1e9, contraction_gflop_per_cell=w["contraction_flops"] / 1e9,
		           allreduce_mb=(w["t1_bytes"] + w["y_bytes"]) / 1e6, projected_ms=total * 1e3, projected_cells_per_s=cells / total,
		           split_ms=dict(rwr=t_rwr * 1e3, contractions=t_con * 1e3, polar_bins=t_pol * 1e3, fixed=t_fix * 1e3, allreduce=t_ar * 1e3),
		           max_gram_side=w["max_gram"], fused_rwr=fused,
		           basis="measured" if name.startswith("2") else "projection from config 2's per-unit rates")
		print("| %s | %d | %d | %d | %d | %d | %.2f | %.1f | %.2f | %.2f | %.0f | %.0f | %.0f | %s |" % (
			name, gpus, cpg, w["bins"], w["blocks"], w["polar_problems"], row["panel_mb_per_cell"], row["resident_gb_per_gpu"],
			row["rwr_dense_gflop_per_cell"], row["contraction_gflop_per_cell"], row["allreduce_mb"], row["projected_ms"],
			row["projected_cells_per_s"], fused))
		print(json.dumps(row), file=sys.stderr)


if __name__ == "__main__":
	main()
