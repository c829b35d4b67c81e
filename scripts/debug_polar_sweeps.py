"""Jacobi sweep counts per block (cold vs warm start) on the bench workload."""
import sys, os, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import fasthigashi_b200
from fasthigashi_b200 import _lib, synth
import fasthigashi_b200.project2orthogonal as P2O
from fasthigashi_b200.parafac2_intergrative import Fast_Higashi_core
import fasthigashi_b200.parafac2_intergrative as core_mod
dev = torch.device("cuda:0")
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 1
bins = synth.chrom_bins("pfc", bench.RES)
ds = bench.make_datasets(cells, 1000, dev, bins)
n_i = bench.probe_rwr_steps(ds)
state = bench.random_state(ds, bench.RANK, 7, n_i=n_i)
core = Fast_Higashi_core(bench.RANK, bench.OFF_DIAG, [bench.RES], ).to(dev)
core.verbose = False
core.prepare(ds, bench.DIM1, True, True, False, state=state)
log = []
orig = P2O.polar_batched
def wrapped(T, rows, cols, ld, out=None, want_sigma=False):
	U, ssum, sig, nsw = orig(T, rows, cols, ld, out=out, want_sigma=want_sigma, want_sweeps=True)
	log.append((min(rows, cols), T.shape[0], float(nsw), int(nsw)))
	return U, ssum, sig
core_mod.polar_batched = wrapped
for sweep in range(4):
	log.clear()
	t = torch.cuda.Event(enable_timing=True); t2 = torch.cuda.Event(enable_timing=True)
	t.record(); core.sweep_once(1); t2.record(); torch.cuda.synchronize()
	tot = sum(b * m for (_, b, m, _) in log) / sum(b for (_, b, _, _) in log)
	print("sweep", sweep, "ms %.0f" % t.elapsed_time(t2), "mean jacobi sweeps %.1f" % tot, "re %.4f" % core.re_trace[-1],
	      "| per block (n, mean, max):", " ".join("%d:%.0f/%d" % (n, m, mx) for (n, _, m, mx) in log[::4]))
