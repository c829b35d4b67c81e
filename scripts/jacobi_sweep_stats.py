"""Jacobi sweep counts per bin on the bench workload (diagnostics)."""
import sys, os, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import fasthigashi_b200
from fasthigashi_b200 import synth
from fasthigashi_b200.parafac2_intergrative import Fast_Higashi_core
dev = torch.device("cuda:0")
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 512
bins = synth.chrom_bins("pfc", bench.RES)
ds = bench.make_datasets(cells, 1000, dev, bins)
state = bench.random_state(ds, bench.RANK, 7, n_i=bench.probe_rwr_steps(ds))
core = Fast_Higashi_core(bench.RANK, bench.OFF_DIAG, [bench.RES]).to(dev)
core.verbose = False
core.prepare(ds, bench.DIM1, True, True, False, state=state)
for sweep in range(4):
	core.sweep_once(1)
	tab = core._polar_table()
	ns = tab["nsweep"].cpu().numpy(); n = tab["n_host"]; slot = tab["slot_dev"].cpu().numpy()
	nsz = np.empty_like(ns); nsz[slot] = n  # size per slot
	print("sweep", sweep, "re %.4f" % core.re_trace[-1], "jacobi sweeps: mean %.1f max %d |" % (ns.mean(), ns.max()),
	      " ".join("n>%d: %.1f/%d" % (lo, ns[nsz > lo].mean(), ns[nsz > lo].max()) for lo in (117, 83, 58, 0)))
