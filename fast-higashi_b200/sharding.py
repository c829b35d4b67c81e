"""Cell-slab sharding for one-process-per-GPU runs: rank g owns a contiguous
slab of cells for ALL chromosomes; A, B, D, U, Y are replicated; T1, Y, the R x R Gram of
SVD_term^T and two scalars are all-reduced every sweep; the per-bin polar problems (which depend only
on all-reduced data) are partitioned across ranks and their inverse square roots exchanged."""


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
		lo, hi = cell_slab(ds.num_cell, world_size, rank)
		nbad = ds.total_cell_num - ds.num_cell
		if nbad == 0:
			out.append(ds.select_cells(lo, hi))
			continue
		blo, bhi = cell_slab(nbad, world_size, rank)
		out.append(ds.select_cell_ranges([(lo, hi), (ds.num_cell + blo, ds.num_cell + bhi)], good_qc_num=hi - lo))
	return out


def gather_cell_rows(local, num_good_local, group=None):
	"""Inverse of `shard_datasets` for per-cell results (e.g. the rows of meta_embedding returned by `transform`):
	`local` (cells_local, ...) holds this rank's good cells then its bad cells; returns the rows of ALL cells in the
	unsharded order (all good cells, then all bad cells) on every rank."""
	import torch
	if group is None:
		return local
	import torch.distributed as dist
	world = dist.get_world_size(group)
	meta = [None] * world
	dist.all_gather_object(meta, (int(local.shape[0]), int(num_good_local)), group=group)
	rows = max(m[0] for m in meta)  # slabs differ by at most one cell: pad to a common size for all_gather
	padded = torch.zeros((rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
	padded[:local.shape[0]] = local
	parts = [torch.empty_like(padded) for _ in meta]
	dist.all_gather(parts, padded, group=group)
	good = [p[:m[1]] for p, m in zip(parts, meta)]
	bad = [p[m[1]:m[0]] for p, m in zip(parts, meta)]
	return torch.cat(good + bad, 0)


def polar_bin_range(num_bins_in_block, world_size, rank):
	"""Bins [lo, hi) of one bin block whose per-bin polar problems `rank` solves. All bins of a block have the same problem
	size, so a contiguous even split of every block balances the ranks AND keeps each rank's share of a block one
	contiguous batch for the Gram / factor GEMMs around the eigen-solver (the whole stage is partitioned, not only the
	Jacobi kernel)."""
	return cell_slab(num_bins_in_block, world_size, rank)


def scatter_dataset(ds, owner, group, device):
	"""`ds` (a Chrom_Dataset holding ALL cells, on the host) lives on rank `owner` only (None elsewhere); every rank gets
	its slab - exactly `shard_datasets([ds], world, rank)[0]` - on `device`. Used by the chromosome-partitioned ingest of
	`FastHigashi.prep_dataset` in distributed mode. Returns (slab, meta) with meta = dict(nnz, shape) of the whole tensor."""
	import torch.distributed as dist
	from .sparse_for_schic import Chrom_Dataset
	world, rank = dist.get_world_size(group), dist.get_rank(group)
	src = dist.get_global_rank(group, owner) if hasattr(dist, "get_global_rank") else owner
	payloads = None
	if rank == owner:
		meta = dict(nnz=ds.nnz(), shape=(ds.num_bin, ds.num_bin, ds.total_cell_num))
		payloads = [(shard_datasets([ds], world, j)[0].to_payload(), meta) for j in range(world)]
	got = [None]
	dist.scatter_object_list(got, payloads, src=src, group=group)
	payload, meta = got[0]
	return Chrom_Dataset.from_payload(payload, device), meta
