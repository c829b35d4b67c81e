"""CPU study for round 2 (no GPU needed): operand-split schemes for the fp32-parity tensor-core products of the RWR chain
(S2 = A A^T, Q <- 1/2 Q P + 1/2 I, X = Q A) - what the 3xTF32 split costs in accuracy against cheaper candidates.

  3xTF32   a = hi + lo with hi = trunc_tf32(a), lo = tf32(a - hi); products hi*hi + hi*lo + lo*hi   (today's kernels)
  3xFP16   the same split into two binary16 values after a power-of-two scaling of each operand matrix; kind::f16
           MMAs run at twice the TF32 rate and the operands take half the shared-memory bytes
  3xBF16   hi/lo bf16 (16 mantissa bits in total): shown as the negative control
  2xFP16   (a_hi + a_lo) * b_hi: only one operand split; 1xFP16: a_hi * b_hi (binary16 rounds to nearest, so the error of
           non-negative operands - everything in the RWR chain is >= 0 - averages out instead of adding up)
  1xTF32   a single TF32 product (the tensor core truncates: a coherent bias)

Operands are split exactly as a kernel would; every product is then accumulated in fp64 here, so the numbers isolate the
operand-representation error (the tensor core's truncating fp32 accumulation is a separate, measured effect).
Prints one JSON line per scheme: relative Frobenius error of the imputed panel X against an fp64 evaluation of the same chain
(north-star tolerance: 1e-5)."""
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


def trunc_bits(x, keep):
	"""fp32 -> keep the top `keep` explicit mantissa bits (truncate), as the tensor core reads a tf32 operand (keep=10)."""
	u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
	return (u & np.uint32(0xFFFFFFFF << (23 - keep) & 0xFFFFFFFF)).view(np.float32)


def split_tf32(a):
	hi = trunc_bits(a, 10)
	lo = trunc_bits((a - hi).astype(np.float32), 10)
	return hi.astype(np.float64), lo.astype(np.float64)


def split_f16(a):
	"""Scale by a power of two so that max|a| sits near 2^14, split into two binary16 values (round to nearest)."""
	m = float(np.max(np.abs(a)))
	s = 2.0 ** (14 - int(np.ceil(np.log2(m)))) if m > 0 else 1.0
	with np.errstate(over="raise"):
		hi = (a * s).astype(np.float16)
		lo = ((a * s).astype(np.float32) - hi.astype(np.float32)).astype(np.float16)
	return hi.astype(np.float64) / s, lo.astype(np.float64) / s


def to_bf16(x):
	u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
	r = ((u >> 16) & 1) + np.uint32(0x7FFF)
	return ((u + r) & np.uint32(0xFFFF0000)).view(np.float32)


def split_bf16(a):
	hi = to_bf16(a)
	lo = to_bf16((a - hi).astype(np.float32))
	return hi.astype(np.float64), lo.astype(np.float64)


def mm3(split, a, b):
	ah, al = split(a.astype(np.float32))
	bh, bl = split(b.astype(np.float32))
	return (ah @ bh + (ah @ bl + al @ bh)).astype(np.float32)


def mm2_f16(a, b):
	"""Two products: a (both halves) times the hi half of b only."""
	ah, al = split_f16(a.astype(np.float32))
	bh, _ = split_f16(b.astype(np.float32))
	return (ah @ bh + al @ bh).astype(np.float32)


def mm1_f16(a, b):
	return (split_f16(a.astype(np.float32))[0] @ split_f16(b.astype(np.float32))[0]).astype(np.float32)


def mm1(a, b):
	return (split_tf32(a.astype(np.float32))[0] @ split_tf32(b.astype(np.float32))[0]).astype(np.float32)


def chain(A, s, e, k, mm):
	"""partial_rwr.py:84-138 for one cell (do_col off) with the three tensor-core products going through `mm`; everything
	else in fp32 like the kernel's drain warps (fp64 when mm is None)."""
	dt = np.float64 if mm is None else np.float32
	mm = mm or (lambda a, b: a @ b)
	A = A.astype(dt)
	nb = A.shape[0]
	S2 = np.array(mm(A, A.T.copy()), dtype=dt)
	np.fill_diagonal(S2, 0)
	S1 = A[:, s:e]
	L = dt(0.75) * S1 / (S1.sum(0, keepdims=True) + dt(1e-15)) + dt(0.25) * S2 / (S2.sum(0, keepdims=True) + dt(1e-15))
	P = L / (L.sum(0, keepdims=True) + dt(1e-15))
	Q = np.eye(nb, dtype=dt)
	for step in range(k):
		Q = (dt(0.5) * (np.array(mm(Q, P), dtype=dt) if step > 0 else P) + dt(0.5) * np.eye(nb, dtype=dt)).astype(dt)
	return np.array(mm(Q, A), dtype=dt)


def main():
	n, ncell, off = 230, 12, 100
	idx, val = synth.synth_chrom(n, ncell, 0.05, off, 5, np.arange(ncell) % 4, 4)
	ds = Chrom_Dataset(Sparse(idx, val, (n, n, ncell), copy=False), bs_bin=115, bs_cell=ncell, compact=True, flank=off)
	schemes = {"3xTF32": lambda a, b: mm3(split_tf32, a, b), "3xFP16_scaled": lambda a, b: mm3(split_f16, a, b),
	           "3xBF16": lambda a, b: mm3(split_bf16, a, b), "2xFP16_scaled": mm2_f16, "1xFP16_scaled": mm1_f16, "1xTF32": mm1, "fp32_numpy": lambda a, b: a @ b}
	errs = {k: [] for k in schemes}
	for b, g in enumerate(ds.geoms):
		x = O.densify_block(ds, b, 0, ncell)
		conv = torch.nn.functional.avg_pool2d(x[:, None], 3, 1, padding=1, ceil_mode=True)[:, 0].clamp_(min=1e-8).numpy()
		for c in range(ncell):
			ref = chain(conv[c], g.s, g.e, 4, None)
			for name, mm in schemes.items():
				got = chain(conv[c], g.s, g.e, 4, mm)
				errs[name].append(float(np.linalg.norm(got - ref) / np.linalg.norm(ref)))
	# a long-K contraction like P1 / P3 / P5 (K = cells): operand 1 with the dynamic range of imputed values (log-normal over
	# four decades), operand 2 a dense factor; every scheme scales per MATRIX (one power of two), not per row
	rng = np.random.default_rng(0)
	Xl = np.exp(rng.normal(-3.0, 2.3, size=(256, 4096))).astype(np.float32)
	Cf = (rng.standard_normal((4096, 144)) * np.exp(rng.normal(0, 1.5, size=(1, 144)))).astype(np.float32)
	ref = Xl.astype(np.float64) @ Cf.astype(np.float64)
	for name, mm in schemes.items():
		got = mm(Xl, Cf)
		print(json.dumps({"scheme": name, "case": "contraction 256 x 4096 x 144", "rel_fro": float(np.linalg.norm(got - ref) / np.linalg.norm(ref)),
		                  "max_rel_col": float(np.max(np.linalg.norm(got - ref, axis=0) / np.linalg.norm(ref, axis=0)))}))
	for name, e in errs.items():
		print(json.dumps({"scheme": name, "rel_fro_max": max(e), "rel_fro_median": float(np.median(e)), "panels": len(e),
		                  "geometry": "nb=115, w=215/230, k=4, density 0.05"}))


if __name__ == "__main__":
	main()
