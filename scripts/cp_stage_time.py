# CP-ALS with several inner iterations: device early stop vs the reference fixture + timing of the cp_als stage at n_iter_parafac = 6
import sys, os, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import bench
import fasthigashi_b200
from fasthigashi_b200 import synth
from fasthigashi_b200.parafac2_intergrative import Fast_Higashi_core
bins = synth.chrom_bins("pfc", bench.RES)
ds = bench.make_datasets(512, 1000, "cuda:0", bins)
state = bench.random_state(ds, 256, 7, n_i=[4] * len(ds))
core = Fast_Higashi_core(256, 100, [bench.RES]).to("cuda:0")
core.verbose = False
core.prepare(ds, 0.6, True, True, False, state=state)
for k in (1, 1, 6, 6, 6):
	core.enable_timers(True)
	core.sweep_once(k)
	t = core.collect_timers()
	print("n_iter_parafac", k, "cp_als ms", round(t["cp_als"], 2), "re", core.re_trace[-1])
