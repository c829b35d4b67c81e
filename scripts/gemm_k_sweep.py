"""Per-tile cost of the tcgen05 3xTF32 GEMM as a function of K and operand layout (batched 128x128 tiles).
usage: python scripts/gemm_k_sweep.py [batch]   (FH_TC_DEBUG=1|2|4 switches off stores / split / MMA)"""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fasthigashi_b200
from fasthigashi_b200 import _lib
dev = torch.device("cuda:0")
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 2072  # 14 tiles per SM
M = N = 128


def run(K, b_mn, reps=10):
	A = torch.randn(batch, M, K, device=dev)
	B = torch.randn(batch, K, N, device=dev) if b_mn else torch.randn(batch, N, K, device=dev)
	Cm = torch.empty(batch, M, N, device=dev)
	sb = (N, 1) if b_mn else (1, K)
	def go():
		_lib.gemm(A, B, Cm, M, N, K, (K, 1), sb, N, batch=batch, batch_strides=(M * K, K * N, M * N), dtype=_lib.GEMM_TF32X3)
	for _ in range(3):
		go()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	torch.cuda.synchronize()
	e0.record()
	for _ in range(reps):
		go()
	e1.record()
	torch.cuda.synchronize()
	us = e0.elapsed_time(e1) * 1e3 / reps
	tiles_per_sm = batch / 148.0
	return us, us / tiles_per_sm


print("FH_TC_DEBUG", os.environ.get("FH_TC_DEBUG", "0"), "batch", batch)
for b_mn in (False, True):
	for K in (32, 64, 128, 256, 512, 1024):
		us, per = run(K, b_mn)
		print("B %s K %4d: %8.1f us  %6.2f us/tile  %5.2f us/kblock  %6.1f TF/s fp32-equiv" %
		      ("MN" if b_mn else "K ", K, us, per, per / (K / 32), 2.0 * M * N * K * batch / us * 1e-6))
