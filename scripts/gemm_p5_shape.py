"""The P5 contraction Z = X^T V at chr1-block size (M = nb*ldw = 36,340, N = 256, K = 4,238 cells; X MN-major, read once
from HBM, V shared by every tile) under the FH_TC_DEBUG knobs: where the tcgen05 GEMM spends its time on a real shape."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fasthigashi_b200
from fasthigashi_b200 import _lib
dev = torch.device("cuda:0")
cells, P, R = 4238, 115 * 316, 256
X = torch.randn(cells, P, device=dev)
V = torch.randn(cells, R, device=dev)
Z = torch.empty(P, R, device=dev)
W = torch.randn(P, R, device=dev)
MT = torch.zeros(cells, R, device=dev)


def timeit(fn, reps=5):
	for _ in range(2):
		fn()
	e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
	torch.cuda.synchronize(); e0.record()
	for _ in range(reps):
		fn()
	e1.record(); torch.cuda.synchronize()
	return e0.elapsed_time(e1) * 1e3 / reps


p5 = timeit(lambda: _lib.gemm(X, V, Z, P, R, cells, (1, P), (R, 1), R, dtype=_lib.GEMM_TF32X3))
p3 = timeit(lambda: _lib.gemm(X, W, MT, cells, R, P, (P, 1), (R, 1), R, beta=1.0, dtype=_lib.GEMM_TF32X3))
fl = 2.0 * P * R * cells
print("FH_TC_DEBUG %s: P5 %.0f us (%.1f TF/s fp32-equiv, %.2f us per k-block tile)  P3 %.0f us (%.1f TF/s)" % (
	os.environ.get("FH_TC_DEBUG", "0"), p5, fl / p5 * 1e-6, p5 * 148 / ((P + 127) // 128 * 2 * ((cells + 31) // 32)), p3, fl / p3 * 1e-6))
