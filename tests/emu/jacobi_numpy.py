"""numpy restatement of the scalar polar pipeline of csrc/fh_polar.cu (pivoted Cholesky, one-sided Jacobi with the
kernel's skip / stop rules) - the yardstick for the block variant's host emulation. TEST INFRASTRUCTURE ONLY."""
import numpy as np


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
			R[k:, k:] = np.eye(n - k) * np.sqrt(max(d0 * 1e-14, 1e-300))
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
	n = R.shape[0]
	m = n + (n & 1)
	mm = m - 1
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
		if worst <= 1e-11:
			return sweep + 1
	return max_sweeps


def polar_from_rows(T, R, perm):
	"""U = T M with M = sum_j w_j w_j^T lambda_j^{-3/2}, w_j = rows of the orthogonalised factor in the original order."""
	W = np.zeros_like(R)
	W[:, perm] = R
	lam = np.maximum((W * W).sum(1), 1e-300)
	lam = np.maximum(lam, lam.max() * 1e-17)
	return T @ ((W.T * lam ** -1.5) @ W)
