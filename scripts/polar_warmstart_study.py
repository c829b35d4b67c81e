"""CPU study for round 2 (no GPU needed): how many one-sided Jacobi sweeps does the per-bin polar step
(csrc/fh_polar.cu: fp64 Gram -> pivoted Cholesky -> Jacobi on the columns of L) need when it is
WARM-STARTED from the previous ALS sweep's eigenvectors, i.e. run on G' = E_prev G E_prev^T instead
of G?  The per-bin matrices temp_i change slowly from one ALS sweep to the next, so G' is nearly
diagonal and Jacobi converges quadratically from the start.

The numpy Jacobi below uses the kernel's rules: round-robin pairs, rotation skipped when
gamma^2 <= 1e-17 min(alpha, beta)^2, stop after a sweep whose largest gamma^2/(alpha beta) <= 1e-11.
Data: the oracle's ALS (oracle/fh_oracle.py) on a synthetic dataset; `polar` is intercepted to record
every bin's temp_i per sweep.

  python scripts/polar_warmstart_study.py [--cells 160] [--sweeps 10]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fasthigashi_b200  # noqa: E402,F401
from fasthigashi_b200 import synth  # noqa: E402
from fasthigashi_b200.sparse_for_schic import Sparse, Chrom_Dataset  # noqa: E402
from oracle import fh_oracle as O  # noqa: E402


def pivoted_cholesky_upper(G):
	n = G.shape[0]
	G = G.copy()
	perm = np.arange(n)
	R = np.zeros_like(G)
	d0 = None
	for k in range(n):
		p = k + int(np.argmax(np.diag(G)[k:]))
		piv = G[p, p]
		if d0 is None:
			d0 = piv
		if piv <= d0 * 1e-14 or not piv > 0:
			rt = np.sqrt(max(d0 * 1e-14, 1e-300))
			R[k:, k:] = np.eye(n - k) * rt
			break
		if p != k:
			G[[k, p], :] = G[[p, k], :]; G[:, [k, p]] = G[:, [p, k]]
			R[:, [k, p]] = R[:, [p, k]]
			perm[[k, p]] = perm[[p, k]]
		R[k, k] = np.sqrt(G[k, k])
		R[k, k + 1:] = G[k, k + 1:] / R[k, k]
		G[k + 1:, k + 1:] -= np.outer(R[k, k + 1:], R[k, k + 1:])
	return R, perm


def jacobi_rows(R, max_sweeps=30, skip=1e-17):
	"""One-sided Jacobi on the ROWS of R (in place); returns (sweeps, rotations applied)."""
	n = R.shape[0]
	m = n + (n & 1)
	mm = m - 1
	rot = 0
	for sweep in range(max_sweeps):
		worst = 0.0
		for step in range(mm):
			t = np.arange(m // 2)
			p = np.where(t == 0, mm, (step + t) % mm)
			q = np.where(t == 0, step, (step - t + mm) % mm)
			ok = (p < n) & (q < n)
			p, q = p[ok], q[ok]
			a, b = R[p], R[q]
			al, be, ga = (a * a).sum(1), (b * b).sum(1), (a * b).sum(1)
			mn = np.minimum(al, be)
			act = (ga * ga > skip * mn * mn) & (ga != 0)
			if not act.any():
				continue
			worst = max(worst, float(np.max((ga * ga / (al * be))[act])))
			zeta = (be - al) / (2 * np.where(act, ga, 1.0))
			tt = np.sign(zeta) / (np.abs(zeta) + np.sqrt(1 + zeta * zeta))
			tt = np.where(zeta == 0, 1.0, tt)
			tt = np.where(act, tt, 0.0)
			cs = 1 / np.sqrt(1 + tt * tt)
			sn = tt * cs
			R[p] = cs[:, None] * a - sn[:, None] * b
			R[q] = sn[:, None] * a + cs[:, None] * b
			rot += int(act.sum())
		if worst <= 1e-11:
			return sweep + 1, rot
	return max_sweeps, rot


def isqrt_via_jacobi(G, E_prev=None):
	"""G^{-1/2}, the eigenvector rows E (for the next warm start), sweeps, rotations."""
	Gw = G if E_prev is None else E_prev @ G @ E_prev.T
	Gw = (Gw + Gw.T) / 2
	R, perm = pivoted_cholesky_upper(Gw)
	sweeps, rot = jacobi_rows(R)
	W = np.zeros_like(R)
	W[:, perm] = R                     # rows w_j in the index order of Gw
	lam = np.maximum((W * W).sum(1), 1e-300)
	E = W / np.sqrt(lam)[:, None]      # orthonormal rows: eigenvectors of Gw
	M = (E.T * lam ** -0.5) @ E
	if E_prev is not None:
		M = E_prev.T @ M @ E_prev
		E = E @ E_prev
	return M, E, sweeps, rot


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument("--cells", type=int, default=160)
	ap.add_argument("--sweeps", type=int, default=10)
	ap.add_argument("--rank", type=int, default=48)
	args = ap.parse_args()
	torch.set_num_threads(os.cpu_count() or 1)
	bins, off, res = [110, 80], 30, 1000000
	chroms, _ = synth.synth_dataset(bins, args.cells, 0.12, off_diag=off, seed=4, num_cluster=5)
	ds = [Chrom_Dataset(Sparse(c["indices"], c["values"], c["shape"]), bs_bin=64, bs_cell=args.cells, compact=True, flank=off,
	                    chrom=c["chrom"], resolution=res) for c in chroms]
	record = []
	orig = O.polar

	def spy(matrix, rank=None):
		if matrix.dim() == 3:
			record[-1].append(matrix.detach().double().numpy().copy())
		return orig(matrix, rank)
	O.polar = spy
	core = O.OracleCore(args.rank, off, [res])
	core.set_sizes(ds, 0.6)
	torch.manual_seed(0); np.random.seed(0)
	core.init_params(ds, True, True, False)
	for it in range(args.sweeps):
		record.append([])
		core.sweep(ds, True, True, False, want_norm=(it == 0))
		for ci, d in enumerate(ds):
			fac, _, _ = O.cp_als(core.projected[d.chrom], [core.A_list[ci], core.B_dict[d.chrom], core.D_dict[d.chrom]], 1)
			core.A_list[ci], core.B_dict[d.chrom], core.D_dict[d.chrom] = fac
	O.polar = orig
	prev = {}
	for it, blocks in enumerate(record):
		cold_s, warm_s, cold_r, warm_r, err = [], [], [], [], []
		for key, blk in enumerate(blocks):
			for i in range(0, blk.shape[0], 7):      # every 7th bin keeps the study short
				Ti = blk[i]
				G = Ti.T @ Ti if Ti.shape[0] >= Ti.shape[1] else Ti @ Ti.T
				M0, E0, s0, r0 = isqrt_via_jacobi(G)
				cold_s.append(s0); cold_r.append(r0)
				k = (key, i)
				if k in prev:
					M1, E1, s1, r1 = isqrt_via_jacobi(G, prev[k])
					warm_s.append(s1); warm_r.append(r1)
					U0 = Ti @ M0 if Ti.shape[0] >= Ti.shape[1] else M0 @ Ti
					U1 = Ti @ M1 if Ti.shape[0] >= Ti.shape[1] else M1 @ Ti
					err.append(float(np.linalg.norm(U1 - U0) / np.linalg.norm(U0)))
					prev[k] = E1
				else:
					prev[k] = E0
		row = {"als_sweep": it, "problems": len(cold_s), "n": int(G.shape[0]), "cold_sweeps_mean": float(np.mean(cold_s)),
		       "cold_rotations_mean": float(np.mean(cold_r))}
		if warm_s:
			row.update({"warm_sweeps_mean": float(np.mean(warm_s)), "warm_sweeps_max": int(np.max(warm_s)),
			            "warm_rotations_mean": float(np.mean(warm_r)), "warm_vs_cold_polar_factor_rel_diff_max": float(np.max(err))})
		print(json.dumps(row), flush=True)


if __name__ == "__main__":
	main()
