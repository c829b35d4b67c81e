"""Synthetic scHi-C tensors with the geometry of BASELINE.json's configs.

Per cell and chromosome a banded, symmetric sparse contact map is drawn: every in-band upper
entry (row, row+d), d <= off_diag, is present with probability ~ alpha/(d+1)^gamma, modulated by a
cluster-specific TAD structure so that the cell embedding is non-trivial. Counts go through the
same value pipeline the reference applies before the tensor exists
(`normalize_by_coverage` preprocessing.py:137-142, `log1p` + clip at mean+15 sigma
FastHigashi_Wrapper.py:358-364). alpha is solved so the expected density (nnz / n^2, both
triangles stored, FastHigashi_Wrapper.py:531 `upper_sim=False`) hits the target.

Everything is torch so the same generator fills a B200 in seconds for the bench and runs on CPU
for the tests. It is data plumbing, not part of the measured path.
"""
import math
import numpy as np
import torch

HG19_LENGTHS = [249250621, 243199373, 198022430, 191154276, 180915260, 171115067, 159138663,
                146364022, 141213431, 135534747, 135006516, 133851895, 115169878, 107349540,
                102531392, 90354753, 81195210, 78077248, 59128983, 63025520, 48129895, 51304566]
# valid-bin counts of the Lee et al. PFC data at 500 kb (PFC tutorial.ipynb cell 6)
PFC_VALID_BINS = [457, 481, 391, 378, 357, 338, 314, 288, 248, 267, 265, 263, 193, 177, 166,
                  160, 158, 151, 114, 121, 74, 71]


def chrom_bins(kind, res):
	if kind == "pfc":
		assert res == 500000
		return list(PFC_VALID_BINS)
	return [int(math.ceil(l / res)) for l in HG19_LENGTHS]


def _solve_alpha(n, off_diag, gamma, density, pmax, weight):
	# expected stored entries per cell = sum_d cnt(d) * mult(d) * weight(d) * min(pmax, alpha/(d+1)^gamma)
	d = np.arange(0, min(off_diag, n - 1) + 1)
	cnt = (n - d).astype(np.float64)
	mult = np.where(d == 0, 1.0, 2.0)
	target = density * n * n
	lo, hi = 0.0, 1e3
	for _ in range(80):
		mid = 0.5 * (lo + hi)
		e = (cnt * mult * weight * np.minimum(pmax, mid / (d + 1.0) ** gamma)).sum()
		if e < target: lo = mid
		else: hi = mid
	return 0.5 * (lo + hi)


@torch.no_grad()
def synth_chrom(n, num_cell, density, off_diag, seed, cluster_of_cell, num_cluster,
                device="cpu", cell_chunk=512, gamma=0.75, cross_tad=0.35):
	"""One chromosome: returns (indices int32 (3, nnz) [row, col, cell], values fp32 (nnz,)),
	sorted by (cell, row, col). Both triangles stored."""
	dev = torch.device(device)
	g = torch.Generator(device=dev)
	g.manual_seed(int(seed))
	D = min(off_diag, n - 1)
	pmax = 0.9
	rows = torch.arange(n, device=dev)
	# cluster-specific TAD ids per bin
	gcpu = torch.Generator().manual_seed(int(seed) + 7919)
	tad_id = torch.empty(num_cluster, n, dtype=torch.int32)
	for k in range(num_cluster):
		nb = max(2, n // 12)
		cuts = torch.sort(torch.randperm(n - 1, generator=gcpu)[:nb] + 1).values
		ids = torch.zeros(n, dtype=torch.int32)
		ids[cuts] = 1
		tad_id[k] = torch.cumsum(ids, 0)
	# fraction of same-TAD pairs per distance -> expected thinning, then solve alpha for the density
	same_frac = np.ones(D + 1)
	tn = tad_id.numpy()
	for d in range(1, D + 1):
		same_frac[d] = float((tn[:, :-d] == tn[:, d:]).mean())
	weight = same_frac + (1.0 - same_frac) * cross_tad
	alpha = _solve_alpha(n, off_diag, gamma, density, pmax, weight)
	dist = torch.arange(D + 1, device=dev, dtype=torch.float32)
	base_p = (alpha / (dist + 1.0) ** gamma).clamp_(max=pmax)  # (D+1,)
	tad_id = tad_id.to(dev)
	cluster_of_cell = torch.as_tensor(cluster_of_cell, device=dev)
	out_idx, out_val = [], []
	for c0 in range(0, num_cell, cell_chunk):
		c1 = min(c0 + cell_chunk, num_cell)
		cc = c1 - c0
		tid = tad_id[cluster_of_cell[c0:c1].long()]  # (cc, n)
		col = rows[:, None] + torch.arange(D + 1, device=dev)[None, :]  # (n, D+1)
		valid = col < n
		colc = col.clamp(max=n - 1)
		same = tid[:, :, None] == torch.gather(
			tid[:, None, :].expand(cc, n, n), 2, colc[None].expand(cc, n, D + 1))
		p = base_p[None, None, :] * torch.where(same, 1.0, cross_tad)
		# per-cell depth jitter
		depth = torch.exp(0.35 * torch.randn(cc, 1, 1, generator=g, device=dev))
		p = (p * depth).clamp_(max=0.95)
		hit = (torch.rand(cc, n, D + 1, generator=g, device=dev) < p) & valid[None]
		cnt = 1.0 + torch.poisson((2.0 * p).clamp_(max=3.0), generator=g)
		cidx, ridx, didx = hit.nonzero(as_tuple=True)
		v = cnt[cidx, ridx, didx]
		r = ridx
		c = ridx + didx
		off = didx > 0
		# symmetrise: both triangles
		rr = torch.cat([r, c[off]])
		cl = torch.cat([c, r[off]])
		ce = torch.cat([cidx, cidx[off]])
		vv = torch.cat([v, v[off]])
		# coverage normalise per cell (scale = n), log1p
		tot = torch.zeros(cc, device=dev, dtype=torch.float64).index_add_(0, ce, vv.double())
		vv = torch.log1p(vv * (n / (tot[ce] + 1e-15)).float())
		key = (ce.long() * n + rr.long()) * n + cl.long()
		order = torch.argsort(key)
		out_idx.append(torch.stack([rr[order], cl[order], ce[order] + c0]).int())
		out_val.append(vv[order].float())
	idx = torch.cat(out_idx, 1)
	val = torch.cat(out_val)
	m, s = val.mean(), val.std(unbiased=False)
	val = val.clamp_(max=float(m + 15 * s))
	return idx, val


def synth_dataset(bins, num_cell, density, off_diag=100, seed=0, num_cluster=6, device="cpu",
                  cell_chunk=512):
	"""All chromosomes. Returns list of dict(chrom, n, indices, values, shape) and the cluster
	label of every cell."""
	rng = np.random.default_rng(seed)
	cluster = rng.integers(0, num_cluster, size=num_cell)
	out = []
	for ci, n in enumerate(bins):
		idx, val = synth_chrom(n, num_cell, density, off_diag, seed * 1000 + ci, cluster,
		                       num_cluster, device=device, cell_chunk=cell_chunk)
		out.append(dict(chrom="chr%d" % (ci + 1), n=n, indices=idx, values=val,
		                shape=(n, n, num_cell)))
	return out, cluster
