"""Bisect of the multi-resolution lock-step run on the GPU (tests/golden/core_multires.npz): prints the loss trace against
the reference's under the switches given in the environment (FH_RWR_FUSED, FH_CP_STREAMS) and --tc 0/1; each stage of the
first sweep is also compared with the oracle's (T1-derived quantities: x_U per dataset, x_V, ||X||^2)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fasthigashi_b200  # noqa
from conftest import load_multires_dataset
from fasthigashi_b200.parafac2_intergrative import Fast_Higashi_core
from oracle import fh_oracle as O
tc = "--tc0" not in sys.argv
ds, g = load_multires_dataset()
res_list = [int(r) for r in g["res"]]
nchrom, nsweep = len(g["chrom2size"]), int(g["nsweep"])
state = ([g["t0_A%d" % i] for i in range(len(ds))], [g["t0_B%d" % c] for c in range(nchrom)],
         [g["t0_D%d" % c] for c in range(nchrom)], g["t0_V"], [g["bin_cov%d" % i] for i in range(len(ds))], [0] * len(ds), g["n_i"])
core = Fast_Higashi_core(int(g["rank"]), int(g["off_diag"]), res_list, use_tc=tc).to("cuda:0")
core.fit(ds, 0.3, nsweep, 1, True, True, False, 0.0, verbose=False, state=state)
re = np.array(core.re_trace)
print("tc", tc, "FUSED", os.environ.get("FH_RWR_FUSED"), "CPST", os.environ.get("FH_CP_STREAMS"))
print(" re  gpu", re)
print(" re  ref", g["re"])
print(" rel", np.abs(re - g["re"]) / g["re"])
ods, _ = load_multires_dataset()
oc = O.OracleCore(int(g["rank"]), int(g["off_diag"]), res_list)
oc.fit(ods, 0.3, nsweep, 1, True, True, False, 0.0, state=state)
print(" re  orc", np.array(oc.re_trace))
for k in range(min(3, len(core.loss_terms))):
	a = core.loss_terms[k]
	b = oc.loss_terms[k] if hasattr(oc, "loss_terms") else None
	print(" sweep", k, "gpu xnorm", a["xnorm"], "core", a["core"], "x_U", a["x_U"], "x_V", a["x_V"])
	if b is not None:
		print(" sweep", k, "orc xnorm", np.ravel(b["xnorm"]), "core", np.ravel(b["core"]), "x_U", np.ravel(b["x_U"]), "x_V", b["x_V"])
