"""The whole job of BASELINE configs 3 / 4 on N GPUs of one box: synthetic hg19-shaped scHi-C at 500 kb, cells sharded in
slabs, `init_params` (device RWR with auto-stop, cell-sharded init SVDs) + S ALS sweeps (tol 0: no early stop) + `transform`.
`bench.py` measures the sweep (the metric); this script measures the full run the north star describes (RWR + 60 sweeps).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \\
      scripts/headline_run.py --cells-total 100000 --sweeps 60          # config 4 (65 GB of imputed panels per GPU)
  python scripts/headline_run.py --cells-total 4238 --sweeps 10 --geometry pfc                       # one GPU

Rank 0 prints ONE JSON line: seconds for data generation (not part of the job), init, sweeps (median and total), transform,
cells/s per sweep, the loss trace and the per-stage device times of the last sweeps. Every rank generates its own slab on its
GPU (seed 1000 + rank), so nothing is read from disk and nothing crosses the host."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def run(args, dev, group, rank, world, bins=None):
	import fasthigashi_b200  # noqa: F401
	from fasthigashi_b200 import synth
	from fasthigashi_b200.sharding import cell_slab
	from fasthigashi_b200.parafac2_intergrative import Fast_Higashi_core
	bins = bins if bins is not None else synth.chrom_bins(args.geometry, bench.RES)
	lo, hi = cell_slab(args.cells_total, world, rank)
	sync = (lambda: torch.cuda.synchronize()) if torch.device(dev).type == "cuda" else (lambda: None)
	t0 = time.perf_counter()
	datasets = bench.make_datasets(hi - lo, 1000 + rank, dev, bins)
	sync()
	t_data = time.perf_counter() - t0
	core = Fast_Higashi_core(args.rank, bench.OFF_DIAG, [bench.RES], cache=args.cache, group=group, init_svd=args.init_svd).to(dev)
	core.verbose = False
	torch.manual_seed(0); np.random.seed(0)
	t0 = time.perf_counter()
	core.prepare(datasets, bench.DIM1, True, True, False)          # sizes + init_params
	sync()
	t_init = time.perf_counter() - t0
	sweep_s = []
	n_iter_parafac = 1
	for it in range(args.sweeps):
		if it % 10 == 0 and it > 0 and n_iter_parafac < 10:        # the reference's schedule (parafac2_intergrative.py:636-637)
			n_iter_parafac += 1
		if it == max(args.sweeps - 3, 0):
			core.enable_timers(True)
		t0 = time.perf_counter()
		core.sweep_once(n_iter_parafac)
		sync()
		sweep_s.append(time.perf_counter() - t0)
	stages = core.collect_timers() if getattr(core, "timers", None) is not None else {}
	core.enable_timers(False)
	core._export()
	t0 = time.perf_counter()
	_, (A_list, B_list, D_list, V), _ = core.transform()
	sync()
	t_transform = time.perf_counter() - t0
	times = torch.tensor([t_init, float(np.sum(sweep_s)), float(np.median(sweep_s)), t_transform], dtype=torch.float64, device=dev)
	if group is not None:
		import torch.distributed as dist
		dist.all_reduce(times, op=dist.ReduceOp.MAX, group=group)
	t_init, t_sweeps, t_median, t_transform = [float(x) for x in times.cpu()]
	timed = max(min(3, args.sweeps), 1)
	out = {"job": "init + %d ALS sweeps + transform" % args.sweeps, "n_gpus": world, "cells_total": args.cells_total,
	       "cells_per_gpu": hi - lo, "geometry": args.geometry, "bins": int(sum(bins)), "rank": args.rank, "init_svd": args.init_svd,
	       "cache": args.cache, "data_generation_s": t_data, "init_s": t_init, "sweeps_s": t_sweeps, "sweep_median_s": t_median,
	       "transform_s": t_transform, "job_s": t_init + t_sweeps + t_transform,
	       "cells_per_s_per_sweep": args.cells_total / t_median if t_median > 0 else None,
	       "rwr_steps": [int(x) for x in core.n_i], "rwr_passes": int(core.n_rwr_passes), "re_first": core.re_trace[:3],
	       "re_last": core.re_trace[-3:], "stages_ms_per_sweep_last": {k: v / timed for k, v in stages.items()},
	       "embedding_rows_local": int(V.shape[0])}
	return out


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument("--cells-total", type=int, default=100000)
	ap.add_argument("--sweeps", type=int, default=60)
	ap.add_argument("--geometry", default="hg19", help="hg19 (configs 3, 4) | pfc (config 2)")
	ap.add_argument("--rank", type=int, default=bench.RANK)
	ap.add_argument("--cache", default="sweep", help="sweep: one RWR pass per sweep | run: imputed panels kept for the whole run")
	ap.add_argument("--init-svd", default="device", help="device: cell-sharded init SVDs | host: the reference's sklearn SVD on rank 0")
	args = ap.parse_args()
	world = int(os.environ.get("WORLD_SIZE", "1"))
	rank = int(os.environ.get("RANK", "0"))
	local_rank = int(os.environ.get("LOCAL_RANK", "0"))
	assert torch.cuda.is_available(), "needs CUDA devices (no CPU path)"
	torch.cuda.set_device(local_rank)
	dev = torch.device("cuda", local_rank)
	group = None
	if world > 1:
		import torch.distributed as dist
		dist.init_process_group("nccl", device_id=dev)
		group = dist.group.WORLD
	import __graft_entry__ as ge
	if rank == 0:
		ge.build()
	if group is not None:
		dist.barrier()
	out = run(args, dev, group, rank, world)
	if rank == 0:
		print(json.dumps(out))
	if group is not None:
		dist.destroy_process_group()


if __name__ == "__main__":
	main()
