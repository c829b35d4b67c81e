"""Host-side logic of the N>1 path on CPU with the gloo backend (world_size 2): slab bounds, per-rank
block-CSR views, and the collective plumbing of the core (`_allreduce`) - no compute kernels involved."""
import os
import socket
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from conftest import load_small_dataset


def _free_port():
	s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
	os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
	dist.init_process_group("gloo", rank=rank, world_size=world)
	from fasthigashi_b200.sharding import cell_slab, shard_datasets
	from fasthigashi_b200.parafac2_intergrative import Fast_Higashi_core
	full = load_small_dataset()
	mine = shard_datasets(full, world, rank)
	lo, hi = cell_slab(full[0].num_cell, world, rank)
	nnz = torch.tensor([float(sum(d.nnz() for d in mine))])
	core = Fast_Higashi_core(8, 12, [1000000], group=dist.group.WORLD)
	core._allreduce(nnz)                                   # the core's own collective helper, SUM
	mx = torch.tensor([hi - lo]); core._allreduce(mx, dist.ReduceOp.MAX)
	# every (cell, row) of the slab keeps exactly its entries
	ok = True
	for d_full, d_mine in zip(full, mine):
		for b, g in enumerate(d_full.geoms):
			rp, col, val = d_full.cell_range_csr(b, lo, hi)
			ok &= torch.equal(rp, d_mine.rowptr[b]) and torch.equal(col, d_mine.col[b]) and torch.equal(val, d_mine.val[b])
	# good + bad QC cells (45 good of 48, so both slabs are uneven): every rank gets its good slab then its bad slab, and
	# gather_cell_rows restores the unsharded order
	from fasthigashi_b200.sharding import gather_cell_rows
	fullb = load_small_dataset(good_qc_num=45)
	mineb = shard_datasets(fullb, world, rank)
	glo, ghi = cell_slab(45, world, rank)
	blo, bhi = cell_slab(3, world, rank)
	okb = True
	for d_full, d_mine in zip(fullb, mineb):
		okb &= d_mine.num_cell == ghi - glo and d_mine.total_cell_num == (ghi - glo) + (bhi - blo)
		for b, g in enumerate(d_full.geoms):
			for (a0, a1), off in (((glo, ghi), 0), ((45 + blo, 45 + bhi), ghi - glo)):
				rp, col, val = d_full.cell_range_csr(b, a0, a1)
				rp2, col2, val2 = d_mine.cell_range_csr(b, off, off + (a1 - a0))
				okb &= torch.equal(rp, rp2) and torch.equal(col, col2) and torch.equal(val, val2)
	ids = torch.cat([torch.arange(glo, ghi), 45 + torch.arange(blo, bhi)]).float()[:, None] * torch.ones(1, 3)
	back = gather_cell_rows(ids, ghi - glo, dist.group.WORLD)
	okb &= torch.equal(back, torch.arange(48).float()[:, None] * torch.ones(1, 3))
	q.put((rank, lo, hi, float(nnz), int(mx), bool(ok and okb), [d.num_cell for d in mine]))
	dist.destroy_process_group()


def test_cell_slabs_cover_everything():
	from fasthigashi_b200.sharding import cell_slab
	for total in [1, 7, 48, 4238, 100000]:
		for world in [1, 2, 3, 8]:
			slabs = [cell_slab(total, world, r) for r in range(world)]
			assert slabs[0][0] == 0 and slabs[-1][1] == total
			assert all(a[1] == b[0] for a, b in zip(slabs[:-1], slabs[1:]))
			assert max(h - l for l, h in slabs) - min(h - l for l, h in slabs) <= 1


def test_two_rank_gloo_sharding():
	ctx = mp.get_context("spawn")
	q = ctx.Queue()
	port = _free_port()
	procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
	for p in procs: p.start()
	res = sorted(q.get(timeout=120) for _ in procs)
	for p in procs: p.join(timeout=60)
	total_nnz = float(sum(d.nnz() for d in load_small_dataset()))
	assert [r[1:3] for r in res] == [(0, 24), (24, 48)]
	assert all(r[3] == total_nnz for r in res)            # SUM all-reduce saw both slabs
	assert all(r[4] == 24 for r in res) and all(r[5] for r in res)
	assert all(r[6] == [24, 24, 24] for r in res)


def test_polar_bin_ranges_are_a_balanced_cover():
	"""Per-bin polar problems across ranks: every bin of every block exactly once, contiguous per block (one GEMM batch per
	rank and block), Jacobi cost (~n^3) balanced to a few per cent at the PFC problem mix (blocks of <= 128 rows)."""
	import math
	from fasthigashi_b200.sharding import polar_bin_range
	from fasthigashi_b200 import synth
	blocks = []  # (rows, problem size) of every bin block: the wrapper's block rule at 500 kb, dim1 0.6
	for n in synth.PFC_VALID_BINS:
		nblk = max(math.ceil(n / 128), 1)
		bs = math.ceil(n / nblk)
		r = min(int(n * 0.6 * 0.5), 256)
		blocks += [(min(bs, n - i * bs), r) for i in range(nblk)]
	for world in (1, 2, 3, 8):
		cost = []
		for rank in range(world):
			c = 0.0
			for nb, r in blocks:
				lo, hi = polar_bin_range(nb, world, rank)
				assert 0 <= lo <= hi <= nb
				if rank + 1 < world:
					assert polar_bin_range(nb, world, rank + 1)[0] == hi
				else:
					assert hi == nb
				c += (hi - lo) * float(r) ** 3
			cost.append(c)
		assert polar_bin_range(blocks[0][0], world, 0)[0] == 0
		assert max(cost) / min(cost) < 1.08
