"""Cell-slab sharding for one-process-per-GPU runs: rank g owns a contiguous
slab of cells for ALL chromosomes; A, B, D, U, Y are replicated; T1, Y, the R x R Gram of
SVD_term^T and two scalars are all-reduced every sweep; the per-bin polar problems (which depend only
on all-reduced data) are partitioned across ranks."""
import numpy as np


def cell_slab(total_cells, world_size, rank):
	"""[lo, hi) of the cells owned by `rank`: slabs differ by at most one cell."""
	base, rem = divmod(int(total_cells), int(world_size))
	lo = rank * base + min(rank, rem)
	return lo, lo + base + (1 if rank < rem else 0)


def shard_datasets(datasets, world_size, rank):
	"""Per-rank view of a list of Chrom_Dataset objects holding ALL cells (good-QC cells first).
	Good and bad cells are sharded separately so every rank keeps the good-first order."""
	out = []
	for ds in datasets:
		if ds.total_cell_num != ds.num_cell:
			raise NotImplementedError("shard before adding bad-QC cells (or pass rank-local tensors)")
		lo, hi = cell_slab(ds.num_cell, world_size, rank)
		out.append(ds.select_cells(lo, hi))
	return out


def polar_partition(sizes, world_size, rank):
	"""Problems of the per-bin polar step owned by `rank`. `sizes`: Gram side of every bin (any order).
	Returns (order, mine): `order` = all problem indices sorted by decreasing size (stable), `mine` = the
	indices this rank factorises, i.e. entries rank, rank + world, ... of `order` - still sorted, and
	balanced to within one problem per size class because neighbours in `order` have (nearly) equal cost."""
	sizes = np.asarray(sizes)
	order = np.argsort(-sizes, kind="stable")
	return order, order[int(rank)::int(world_size)]
