"""Host-side cost of one ALS sweep: the product's own orchestration (`Fast_Higashi_core.sweep_once`) run on this machine's
CPU against a NULL stand-in of the C ABI - every `fh_*` entry returns FH_OK at once - at the bin geometry of the bench
(BASELINE config 2) and a small cell count. What is timed is therefore only what the host has to do per sweep before
the GPU can be kept busy: Python control flow, descriptor structs, ctypes calls, torch views / allocations (torch-CPU ops
on small tensors stand in for what are asynchronous launches on the device, so this is an upper bound).
If this number approaches the device time of a sweep (178 ms at 4,238 cells), the sweep is launch-bound and belongs in a
CUDA graph. No GPU, no kernels: says nothing about results (the factors are garbage after a null sweep).

    python scripts/host_overhead.py [--cells 64] [--sweeps 3]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class NullLib:
	"""Every C-ABI entry point: count the call, report success. Size queries get a small positive answer."""
	def __init__(self):
		self.calls = {}

	def fh_last_error(self):
		return b""

	def __getattr__(self, name):
		if not name.startswith("fh_"):
			raise AttributeError(name)
		size_query = name.endswith("_bytes") or name.endswith("_ws") or "workspace" in name

		def call(*a, **k):
			self.calls[name] = self.calls.get(name, 0) + 1
			return 1 << 20 if size_query else 0
		call.restype = None
		call.argtypes = None
		setattr(self, name, call)
		return call


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument("--cells", type=int, default=64)
	ap.add_argument("--sweeps", type=int, default=3)
	ap.add_argument("--geometry", default="pfc")
	args = ap.parse_args()
	import bench
	import fasthigashi_b200  # noqa: F401
	from emu import fake_abi
	from fasthigashi_b200 import _lib
	from fasthigashi_b200.parafac2_intergrative import Fast_Higashi_core
	from fasthigashi_b200 import synth
	bins = synth.chrom_bins(args.geometry, bench.RES)
	datasets = bench.make_datasets(args.cells, 1000, "cpu", bins)
	fake, undo = fake_abi.install()
	null = NullLib()
	_lib.lib = lambda: null
	try:
		state = bench.random_state(datasets, bench.RANK, 7, n_i=[4] * len(datasets))
		core = Fast_Higashi_core(bench.RANK, bench.OFF_DIAG, [bench.RES], cache="sweep", group=None, use_tc=True)
		core.verbose = False
		core.prepare(datasets, bench.DIM1, True, True, False, state=state)
		core.sweep_once(1)
		null.calls.clear()
		# zero-fills of operand buffers are real memsets of host memory here and asynchronous memset launches on the device:
		# timed apart, with their bytes
		fill = {"s": 0.0, "bytes": 0, "n": 0}
		zeros, zero_ = torch.zeros, torch.Tensor.zero_

		def timed_zeros(*a, **k):
			t = time.perf_counter()
			out = zeros(*a, **k)
			fill["s"] += time.perf_counter() - t; fill["bytes"] += out.numel() * out.element_size(); fill["n"] += 1
			return out

		def timed_zero_(self):
			t = time.perf_counter()
			out = zero_(self)
			fill["s"] += time.perf_counter() - t; fill["bytes"] += self.numel() * self.element_size(); fill["n"] += 1
			return out
		torch.zeros, torch.Tensor.zero_ = timed_zeros, timed_zero_
		try:
			t0 = time.perf_counter()
			for _ in range(args.sweeps):
				core.sweep_once(1)
			dt = (time.perf_counter() - t0) / args.sweeps
		finally:
			torch.zeros, torch.Tensor.zero_ = zeros, zero_
	finally:
		undo()
	ncall = sum(null.calls.values()) / args.sweeps
	top = sorted(null.calls.items(), key=lambda kv: -kv[1])[:8]
	print(json.dumps({"what": "host orchestration of one sweep against a null C ABI (CPU only)", "cells": args.cells,
	                  "geometry": args.geometry, "host_ms_per_sweep": dt * 1e3,
	                  "of_which_zero_fill_ms": fill["s"] * 1e3 / args.sweeps, "zero_fill_mb_per_sweep": fill["bytes"] / 1e6 / args.sweeps,
	                  "zero_fills_per_sweep": fill["n"] / args.sweeps,
	                  "host_ms_per_sweep_without_fills": (dt - fill["s"] / args.sweeps) * 1e3, "abi_calls_per_sweep": ncall,
	                  "us_per_call_without_fills": (dt - fill["s"] / args.sweeps) * 1e6 / max(ncall, 1), "top_calls_per_sweep": {k: v / args.sweeps for k, v in top},
	                  "threads": torch.get_num_threads()}))


if __name__ == "__main__":
	main()
