"""torchrun --nproc-per-node N scripts/multigpu_check.py : the cell-sharded run must reproduce the
single-GPU run (same state, same data): loss per sweep and final factors."""
import os, sys
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fasthigashi_b200
from conftest import load_small_dataset, GOLDEN
from fasthigashi_b200.parafac2_intergrative import Fast_Higashi_core
from fasthigashi_b200.sharding import shard_datasets, cell_slab
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
g = np.load(os.path.join(GOLDEN, "core_nocol.npz"))
lo, hi = cell_slab(48, world, rank)
state_full = ([g["t0_A%d" % i] for i in range(3)], [g["t0_B%d" % i] for i in range(3)], [g["t0_D%d" % i] for i in range(3)],
              g["t0_V"], [g["bin_cov%d" % i] for i in range(3)], [0, 0, 0], g["n_i"])
state_loc = state_full[:3] + (g["t0_V"][lo:hi],) + ([g["bin_cov%d" % i][lo:hi] for i in range(3)], [0, 0, 0], g["n_i"])
full = load_small_dataset()
core = Fast_Higashi_core(int(g["rank"]), 12, [1000000], group=dist.group.WORLD).to(dev)
core.fit(shard_datasets(full, world, rank), 0.3, 5, 1, True, True, False, 0.0, verbose=False, state=state_loc)
ref = Fast_Higashi_core(int(g["rank"]), 12, [1000000]).to(dev)
ref.fit(load_small_dataset(), 0.3, 5, 1, True, True, False, 0.0, verbose=False, state=state_full)
a, b = np.array(core.re_trace), np.array(ref.re_trace)
err = float(np.max(np.abs(a - b) / b))
dA = max(float((x - y).abs().max() / y.abs().max()) for x, y in zip(core.A_dev, ref.A_dev))
dV = float((core.meta_embedding - ref.meta_embedding[lo:hi]).abs().max())
print("rank %d/%d cells [%d,%d): re sharded %s | single %s | max rel %.2e | dA %.2e dV %.2e" % (rank, world, lo, hi, np.round(a, 6), np.round(b, 6), err, dA, dV))
assert err < 1e-5 and dA < 1e-3 and dV < 1e-3
dist.barrier()
if rank == 0: print("MULTIGPU_CHECK_OK")
dist.destroy_process_group()
