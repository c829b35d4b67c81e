/* fh_host.h - C ABI of libfh_host.so: the host-side (CPU, OpenMP) ingest stage that feeds the
 * B200 hot path. Plain pointers and sizes only; no torch / numpy / C++ types.
 *
 * Replaces, on the caller side of the hot path, these pieces of the reference
 * (all in fasthigashi/FastHigashi_Wrapper.py unless noted):
 *   fh_host_qc_chrom    <- get_qc                          :428-458 (per-chromosome part)
 *   fh_host_pack_chrom  <- pack_training_data_one_process  :221-366 with
 *                          preprocessing.filter_bin :474-489, normalize_per_batch :232-292,
 *                          norm2 :195-215, normalize_by_coverage :137-142
 *   fh_host_block_csr_* <- sparse_for_schic.py:356-510 (Chrom_Dataset.__init__: the bin-block x cell-batch
 *                          split into Fake_Sparse objects, :322-353), producing the device layout instead
 * The reference walks a Python list of scipy matrices (one object per cell, about ten passes);
 * here one call handles a chromosome with the cells spread over the host threads.
 *
 * Input of both calls: the per-cell CSR matrices of `raw/{chrom}_sparse_adj.npy` as arrays of
 * per-cell base pointers (no copy, no concatenation on the Python side).
 * Every function returns 0 on success, a negative code on error; fh_host_last_error() has the text.
 */
#ifndef FH_HOST_H
#define FH_HOST_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { FH_HOST_OK = 0, FH_HOST_EINVAL = -1, FH_HOST_ENOMEM = -2, FH_HOST_ENAN = -3, FH_HOST_EWINDOW = -4, FH_HOST_EDUP = -5 };
/* element types of the scipy arrays */
enum { FH_HOST_I32 = 0, FH_HOST_I64 = 1, FH_HOST_F32 = 2, FH_HOST_F64 = 3 };

typedef struct {
	int64_t num_cell;
	int32_t n_row, n_col;           /* shape of every cell's matrix */
	const void* const* indptr;      /* [num_cell] -> (n_row + 1) offsets, rebased per cell */
	const void* const* indices;     /* [num_cell] -> column ids */
	const void* const* data;        /* [num_cell] -> values */
	int32_t index_type;             /* FH_HOST_I32 | FH_HOST_I64 (indptr and indices) */
	int32_t data_type;              /* FH_HOST_F32 | FH_HOST_F64 | FH_HOST_I32 | FH_HOST_I64 */
} fh_host_cells;

typedef struct {
	int32_t off_diag;               /* keep |col - row| <= off_diag (after merging) */
	int32_t merge_row, merge_col;   /* bin coarsening factors (>= 1); colliding contacts of a cell are summed */
	const uint8_t* dead_bin;        /* [max(n_row, n_col)] blacklist mask over RAW bins, or NULL */
	const int32_t* batch_of_cell;   /* [num_cell] batch index in [0, num_batch), or NULL (no batches) */
	int32_t num_batch;
	int32_t batch_norm;             /* apply the per-batch normalisation (needs batch_of_cell) */
	int32_t num_threads;            /* <= 0: all */
} fh_host_pack_opts;

const char* fh_host_last_error(void);
int fh_host_version(void);

/* get_qc, one chromosome: contacts[c] = (nnz + #positive diagonal entries) / 2, reads[c] = sum of the
 * cell's values, *n_bin = number of bins whose pooled row has more than 0.1 * n_row * scale non-zeros. */
int fh_host_qc_chrom(const fh_host_cells* cells, int32_t scale, double* contacts, double* reads,
                     int64_t* n_bin, int32_t num_threads);

/* pack, phase 1: everything up to log1p + clip. The result stays in an opaque handle; *nnz and
 * *num_bins (valid bins = side of the output tensor) tell the caller what to allocate. */
int fh_host_pack_chrom(const fh_host_cells* cells, const fh_host_pack_opts* opts, void** handle,
                       int64_t* nnz, int32_t* num_bins);
/* phase 2: copy out. indices: int32 [3 * nnz] as three rows (row, col, cell), values: fp32 [nnz];
 * entries are cell-major, row-major inside a cell (the reference's order). */
int fh_host_pack_fetch(void* handle, int32_t* indices, float* values);
void fh_host_pack_free(void* handle);

/* Block-CSR staging: the chromosome tensor as COO (row, col, cell; any order) -> for every bin block b ONE CSR over
 * (cell, local row): rowptr int32 [num_cell * nb[b] + 1], window-local columns int16 (col - col0[b]), values fp32,
 * columns ascending inside a row. This is the layout fh_rwr_batched / fh_densify read (include/fh_b200.h); the
 * reference builds one pinned COO object per (bin block, cell batch) instead (sparse_for_schic.py:440-499). */
typedef struct {
	int32_t num_bin;                /* rows (= columns) of the chromosome tensor */
	int32_t bs_bin;                 /* rows per bin block; block of row r = r / bs_bin */
	int32_t num_block;
	const int32_t* nb;              /* [num_block] rows of block b (bs_bin except the last) */
	const int32_t* col0;            /* [num_block] first global column of the block's window */
	const int32_t* w;               /* [num_block] window width (<= 32767) */
	int64_t num_cell;               /* cells (good-QC first, then bad-QC: the caller's order is kept) */
} fh_host_block_geom;

/* pass 1: rowptr[b] (caller-allocated, [num_cell * nb[b] + 1]) and nnz_block[b]. index_type: FH_HOST_I32 | FH_HOST_I64.
 * FH_HOST_EWINDOW when a contact lies outside its block's window (|col - row| > flank: filter with off_diag first). */
int fh_host_block_csr_count(const void* row, const void* col, const void* cell, int32_t index_type, int64_t nnz,
                            const fh_host_block_geom* geom, int32_t* const* rowptr, int64_t* nnz_block, int32_t num_threads);
/* pass 2: col_out[b] int16 [nnz_block[b]], val_out[b] fp32 [nnz_block[b]] (caller-allocated). rowptr is pass 1's output
 * for the same input; it is used as scratch during the call and restored. FH_HOST_EDUP when two entries share
 * (row, col, cell). The result does not depend on the thread count or schedule. */
int fh_host_block_csr_fill(const void* row, const void* col, const void* cell, int32_t index_type, const float* val,
                           int64_t nnz, const fh_host_block_geom* geom, int32_t* const* rowptr, int16_t* const* col_out,
                           float* const* val_out, int32_t num_threads);

#ifdef __cplusplus
}
#endif
#endif
