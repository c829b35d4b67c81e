// fh_host.cpp - host-side ingest stage (C ABI in include/fh_host.h), OpenMP over cells.
//
// One call per chromosome turns the per-cell CSR matrices of raw/{chrom}_sparse_adj.npy into the
// normalised COO tensor the block-CSR staging consumes. Behaviour follows the reference's
// FastHigashi_Wrapper.py:221-366 (pack_training_data_one_process) and :428-458 (get_qc); the
// structure does not: cells are independent until the pooled ("bulk") matrix is needed, so phase A
// filters / coarsens every cell in parallel into its own compact entry list, the bulk is reduced
// from per-thread partial matrices, and the per-contact rescaling, log1p and clip run in parallel
// over the flat output.
#include "../../include/fh_host.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
	return code;
}

struct Entry {
	int32_t r, c;
	double v;
};

inline int64_t load_index(const void* p, int64_t i, int type) {
	return type == FH_HOST_I64 ? ((const int64_t*)p)[i] : (int64_t)((const int32_t*)p)[i];
}

inline double load_value(const void* p, int64_t i, int type) {
	switch (type) {
	case FH_HOST_F32: return (double)((const float*)p)[i];
	case FH_HOST_F64: return ((const double*)p)[i];
	case FH_HOST_I32: return (double)((const int32_t*)p)[i];
	default: return (double)((const int64_t*)p)[i];
	}
}

int check_cells(const fh_host_cells* c) {
	if (!c) return fail(FH_HOST_EINVAL, "cells == NULL");
	if (c->num_cell <= 0 || c->n_row <= 0 || c->n_col <= 0) return fail(FH_HOST_EINVAL, "empty input (cells %lld, shape %d x %d)", (long long)c->num_cell, c->n_row, c->n_col);
	if (!c->indptr || !c->indices || !c->data) return fail(FH_HOST_EINVAL, "NULL pointer table");
	if (c->index_type != FH_HOST_I32 && c->index_type != FH_HOST_I64) return fail(FH_HOST_EINVAL, "index_type must be I32 or I64");
	if (c->data_type < FH_HOST_I32 || c->data_type > FH_HOST_F64) return fail(FH_HOST_EINVAL, "bad data_type");
	return FH_HOST_OK;
}

int thread_count(int requested) {
#ifdef _OPENMP
	int t = requested > 0 ? requested : omp_get_max_threads();
	return std::max(1, t);
#else
	(void)requested;
	return 1;
#endif
}

// Dense pooled matrix of the selected cells: per-thread partial sums over a static split of the
// cells, added in thread order (deterministic for a given thread count; exact for integer counts).
void pooled_matrix(const std::vector<std::vector<Entry>>& cells, const int32_t* batch_of_cell, int batch, int n0, int n1,
                   int threads, std::vector<double>& out) {
	const size_t sz = (size_t)n0 * n1;
	out.assign(sz, 0.0);
	const size_t budget = (size_t)2 << 30;  // bytes of per-thread partials
	int T = (int)std::max<size_t>(1, std::min<size_t>((size_t)threads, budget / (sz * sizeof(double) + 1)));
	std::vector<std::vector<double>> part((size_t)T);
	const int64_t C = (int64_t)cells.size();
#pragma omp parallel num_threads(T)
	{
#ifdef _OPENMP
		const int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
		const int t = 0, nt = 1;
#endif
		std::vector<double>& p = part[(size_t)t];
		p.assign(sz, 0.0);
		const int64_t lo = C * t / nt, hi = C * (t + 1) / nt;
		for (int64_t c = lo; c < hi; ++c) {
			if (batch_of_cell && batch_of_cell[c] != batch) continue;
			for (const Entry& e : cells[(size_t)c]) p[(size_t)e.r * n1 + e.c] += e.v;
		}
	}
	for (int t = 0; t < T; ++t) {
		if (part[(size_t)t].empty()) continue;
		const double* p = part[(size_t)t].data();
#pragma omp parallel for num_threads(threads) schedule(static)
		for (int64_t i = 0; i < (int64_t)sz; ++i) out[(size_t)i] += p[i];
	}
}

// sum of the k-th diagonal of m scaled by 1 / ((sqrt(cov[j]) + eps)(sqrt(cov[i]) + eps)), both triangles for k > 0
// (preprocessing.py:236-251); total = sum of the whole scaled matrix.
void distance_profile(const std::vector<double>& m, const std::vector<double>& cov, int n0, int n1, int len,
                      std::vector<double>& prof, double& total) {
	prof.assign((size_t)len, 0.0);
	std::vector<double> inv_r((size_t)n0), inv_c((size_t)n1);
	for (int i = 0; i < n0; ++i) inv_r[(size_t)i] = 1.0 / (std::sqrt(cov[(size_t)i]) + 1e-15);
	for (int j = 0; j < n1; ++j) inv_c[(size_t)j] = 1.0 / (std::sqrt(j < n0 ? cov[(size_t)j] : 0.0) + 1e-15);
	total = 0.0;
	for (int i = 0; i < n0; ++i) {
		double row = 0.0;
		for (int j = 0; j < n1; ++j) {
			const double x = m[(size_t)i * n1 + j] * inv_c[(size_t)j] * inv_r[(size_t)i];
			row += x;
			const int k = j - i;
			if (k >= 0 && k < len) prof[(size_t)k] += (k == 0 ? x : 2.0 * x);
		}
		total += row;
	}
}

// FH_HOST_TRACE=1 prints the wall time of each phase of fh_host_pack_chrom to stderr
struct Tracer {
	bool on;
	double t0;
	Tracer() : on(getenv("FH_HOST_TRACE") && getenv("FH_HOST_TRACE")[0] == '1'), t0(now()) {}
	static double now() {
#ifdef _OPENMP
		return omp_get_wtime();
#else
		return 0.0;
#endif
	}
	void mark(const char* what) {
		if (!on) return;
		const double t = now();
		fprintf(stderr, "[fh_host] %-28s %8.1f ms\n", what, (t - t0) * 1e3);
		t0 = t;
	}
};

// result of phase 1: the normalised, bin-mapped entries of every cell (values before log1p)
struct Packed {
	std::vector<std::vector<Entry>> ent;
	std::vector<int64_t> offset;
	int32_t num_bins = 0;
	int threads = 1;
};

}  // namespace

extern "C" const char* fh_host_last_error(void) { return g_err; }
extern "C" int fh_host_version(void) { return 100; }

extern "C" int fh_host_qc_chrom(const fh_host_cells* cells, int32_t scale, double* contacts, double* reads,
                                int64_t* n_bin, int32_t num_threads) {
	int rc = check_cells(cells);
	if (rc) return rc;
	if (!contacts || !reads || !n_bin) return fail(FH_HOST_EINVAL, "NULL output");
	const int n0 = cells->n_row, n1 = cells->n_col;
	const int threads = thread_count(num_threads);
	const int64_t C = cells->num_cell;
	const size_t sz = (size_t)n0 * n1;
	// pooled "is there any positive contact" map; get_qc only needs bulk > 0 per (row, col)
	int T = (int)std::max<size_t>(1, std::min<size_t>((size_t)threads, ((size_t)2 << 30) / (sz * sizeof(double) + 1)));
	std::vector<std::vector<double>> part((size_t)T);
	int bad = 0;
#pragma omp parallel num_threads(T)
	{
#ifdef _OPENMP
		const int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
		const int t = 0, nt = 1;
#endif
		std::vector<double>& p = part[(size_t)t];
		p.assign(sz, 0.0);
		std::vector<double> diag((size_t)n0);
		const int64_t lo = C * t / nt, hi = C * (t + 1) / nt;
		for (int64_t c = lo; c < hi; ++c) {
			const void* ip = cells->indptr[c];
			const void* ix = cells->indices[c];
			const void* dv = cells->data[c];
			std::fill(diag.begin(), diag.end(), 0.0);
			double total = 0.0;
			const int64_t nnz = load_index(ip, n0, cells->index_type);
			for (int r = 0; r < n0; ++r) {
				const int64_t a = load_index(ip, r, cells->index_type), b = load_index(ip, r + 1, cells->index_type);
				for (int64_t k = a; k < b; ++k) {
					const int64_t col = load_index(ix, k, cells->index_type);
					if (col < 0 || col >= n1) {
#pragma omp atomic write
						bad = 1;
						continue;
					}
					const double v = load_value(dv, k, cells->data_type);
					total += v;
					p[(size_t)r * n1 + col] += v;
					if (col == r) diag[(size_t)r] += v;
				}
			}
			int64_t dpos = 0;
			for (int r = 0; r < n0; ++r) dpos += diag[(size_t)r] > 0;
			contacts[c] = (double)(nnz + dpos) / 2.0;
			reads[c] = total;
		}
	}
	if (bad) return fail(FH_HOST_EINVAL, "column index outside [0, %d)", n1);
	int64_t nb = 0;
	for (int i = 0; i < n0; ++i) {
		int64_t cov = 0;
		for (int j = 0; j < n1; ++j) {
			double s = 0.0;
			for (int t = 0; t < T; ++t)
				if (!part[(size_t)t].empty()) s += part[(size_t)t][(size_t)i * n1 + j];
			cov += s > 0;
		}
		nb += (double)cov > 0.1 * (double)n0 * (double)scale;
	}
	*n_bin = nb;
	return FH_HOST_OK;
}

extern "C" int fh_host_pack_chrom(const fh_host_cells* cells, const fh_host_pack_opts* o, void** handle,
                                  int64_t* nnz_out, int32_t* num_bins_out) {
	int rc = check_cells(cells);
	if (rc) return rc;
	if (!o || !handle || !nnz_out || !num_bins_out) return fail(FH_HOST_EINVAL, "NULL argument");
	if (o->merge_row < 1 || o->merge_col < 1 || o->off_diag < 0) return fail(FH_HOST_EINVAL, "bad merge factors / off_diag");
	if (o->batch_of_cell && o->num_batch < 1) return fail(FH_HOST_EINVAL, "num_batch < 1");
	*handle = nullptr;
	const int threads = thread_count(o->num_threads);
	const int64_t C = cells->num_cell;
	const int raw0 = cells->n_row, raw1 = cells->n_col;
	const bool merging = o->merge_row > 1 || o->merge_col > 1;
	// the reference resizes to ceil(shape / [merge_col, merge_row]) (:261); square matrices with equal factors in practice
	const int n0 = merging ? (raw0 + o->merge_col - 1) / o->merge_col : raw0;
	const int n1 = merging ? (raw1 + o->merge_row - 1) / o->merge_row : raw1;
	if (o->batch_of_cell)
		for (int64_t c = 0; c < C; ++c)
			if (o->batch_of_cell[c] < 0 || o->batch_of_cell[c] >= o->num_batch) return fail(FH_HOST_EINVAL, "batch id of cell %lld out of range", (long long)c);
	try {
		Tracer trace;
		// ---- phase A: per-cell blacklist, coarsening (+ summing what collides), band filter
		Packed* P = new Packed();
		std::unique_ptr<Packed> guard(P);
		std::vector<std::vector<Entry>>& ent = P->ent;
		ent.resize((size_t)C);
		int bad = 0;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 16)
		for (int64_t c = 0; c < C; ++c) {
			const void* ip = cells->indptr[c];
			const void* ix = cells->indices[c];
			const void* dv = cells->data[c];
			std::vector<Entry>& e = ent[(size_t)c];
			e.reserve((size_t)load_index(ip, raw0, cells->index_type));
			for (int r = 0; r < raw0; ++r) {
				const int64_t a = load_index(ip, r, cells->index_type), b = load_index(ip, r + 1, cells->index_type);
				for (int64_t k = a; k < b; ++k) {
					const int64_t col = load_index(ix, k, cells->index_type);
					if (col < 0 || col >= raw1) {
#pragma omp atomic write
						bad = 1;
						continue;
					}
					const double v = load_value(dv, k, cells->data_type);
					if (o->dead_bin && (o->dead_bin[r] || o->dead_bin[col] || v == 0.0)) continue;
					const int rr = r / o->merge_row, cc = (int)(col / o->merge_col);
					if (!merging && std::abs(cc - rr) > o->off_diag) continue;
					e.push_back(Entry{rr, cc, v});
				}
			}
			if (merging) {
				std::stable_sort(e.begin(), e.end(), [](const Entry& x, const Entry& y) { return x.r != y.r ? x.r < y.r : x.c < y.c; });
				size_t w = 0;
				for (size_t i = 0; i < e.size();) {
					Entry acc = e[i];
					size_t j = i + 1;
					for (; j < e.size() && e[j].r == acc.r && e[j].c == acc.c; ++j) acc.v += e[j].v;
					if (std::abs(acc.c - acc.r) <= o->off_diag) e[w++] = acc;
					i = j;
				}
				e.resize(w);
			}
		}
		if (bad) return fail(FH_HOST_EINVAL, "column index outside [0, %d)", raw1);

		trace.mark("filter/coarsen per cell");
		// ---- pooled matrix, valid bins (preprocessing.py:474-489: any coverage)
		std::vector<double> bulk;
		std::vector<std::vector<double>> batch_bulk;
		if (o->batch_of_cell) {
			batch_bulk.resize((size_t)o->num_batch);
			bulk.assign((size_t)n0 * n1, 0.0);
			for (int b = 0; b < o->num_batch; ++b) {
				pooled_matrix(ent, o->batch_of_cell, b, n0, n1, threads, batch_bulk[(size_t)b]);
				for (size_t i = 0; i < bulk.size(); ++i) bulk[i] += batch_bulk[(size_t)b][i];
			}
		} else {
			pooled_matrix(ent, nullptr, 0, n0, n1, threads, bulk);
		}
		std::vector<double> bk_cov((size_t)n0, 0.0);
		for (int i = 0; i < n0; ++i) {
			double s = 0.0;
			for (int j = 0; j < n1; ++j) s += bulk[(size_t)i * n1 + j];
			bk_cov[(size_t)i] = s;
		}
		std::vector<int32_t> map_bin((size_t)std::max(n0, n1), -1);
		int32_t num_bins = 0;
		for (int i = 0; i < n0; ++i)
			if (bk_cov[(size_t)i] / (double)C > 0.0) map_bin[(size_t)i] = num_bins++;

		trace.mark("pooled matrix");
		// ---- per-batch normalisation tables (preprocessing.py:232-292)
		const bool bnorm = o->batch_of_cell && o->batch_norm;
		const int len = o->off_diag + 1;
		std::vector<std::vector<double>> info, bcov;
		if (bnorm) {
			std::vector<double> bulk_ratio, prof;
			double total = 0.0;
			distance_profile(bulk, bk_cov, n0, n1, len, bulk_ratio, total);
			for (double& x : bulk_ratio) x /= total;
			info.resize((size_t)o->num_batch);
			bcov.resize((size_t)o->num_batch);
			for (int b = 0; b < o->num_batch; ++b) {
				std::vector<double>& cov = bcov[(size_t)b];
				cov.assign((size_t)n0, 0.0);
				for (int i = 0; i < n0; ++i) {
					double s = 0.0;
					for (int j = 0; j < n1; ++j) s += batch_bulk[(size_t)b][(size_t)i * n1 + j];
					cov[(size_t)i] = s;
				}
				distance_profile(batch_bulk[(size_t)b], cov, n0, n1, len, prof, total);
				info[(size_t)b].resize((size_t)len);
				for (int k = 0; k < len; ++k) info[(size_t)b][(size_t)k] = prof[(size_t)k] / (total + 1e-15) / (bulk_ratio[(size_t)k] + 1e-15);
			}
		}
		std::vector<double>().swap(bulk);
		batch_bulk.clear();

		trace.mark("batch tables");
		// ---- per-cell: batch rescale, coverage normalisation (preprocessing.py:137-142), valid-bin map
		std::vector<int64_t>& offset = P->offset;
		offset.assign((size_t)C + 1, 0);
		int nan_seen = 0;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 16)
		for (int64_t c = 0; c < C; ++c) {
			std::vector<Entry>& e = ent[(size_t)c];
			if (bnorm) {
				const int b = o->batch_of_cell[c];
				const std::vector<double>& cov = bcov[(size_t)b];
				for (Entry& x : e) {
					const double cr = cov[(size_t)x.r], cc = x.c < n0 ? cov[(size_t)x.c] : 0.0;
					const double kr = bk_cov[(size_t)x.r], kc = x.c < n0 ? bk_cov[(size_t)x.c] : 0.0;
					double v = x.v / (std::sqrt(cr) * std::sqrt(cc)) * (std::sqrt(kr) * std::sqrt(kc));
					x.v = v / (info[(size_t)b][(size_t)std::abs(x.r - x.c)] + 1e-15);
				}
			}
			double total = 0.0;
			for (const Entry& x : e) total += x.v;
			const double scale = (double)n0 / (total + 1e-15);
			size_t w = 0;
			for (size_t i = 0; i < e.size(); ++i) {
				const int32_t r = map_bin[(size_t)e[i].r], cc = map_bin[(size_t)e[i].c];
				const double v = e[i].v * scale;
				if (v != v) {
#pragma omp atomic write
					nan_seen = 1;
				}
				if (r < 0 || cc < 0) continue;
				e[w++] = Entry{r, cc, v};
			}
			e.resize(w);
			offset[(size_t)c + 1] = (int64_t)w;
		}
		if (nan_seen) return fail(FH_HOST_ENAN, "NaN after normalisation");
		for (int64_t c = 0; c < C; ++c) offset[(size_t)c + 1] += offset[(size_t)c];
		const int64_t nnz = offset[(size_t)C];

		trace.mark("normalise + map bins");
		P->num_bins = num_bins;
		P->threads = threads;
		*handle = guard.release();
		*nnz_out = nnz;
		*num_bins_out = num_bins;
		return FH_HOST_OK;
	} catch (const std::bad_alloc&) {
		return fail(FH_HOST_ENOMEM, "out of host memory");
	}
}

// phase 2 writes straight into the caller's arrays: flatten, log1p, clip at mean + 15 sigma (:353-361)
extern "C" int fh_host_pack_fetch(void* handle, int32_t* indices, float* values) {
	if (!handle) return fail(FH_HOST_EINVAL, "NULL handle");
	const Packed* P = (const Packed*)handle;
	const int64_t C = (int64_t)P->ent.size();
	const int64_t nnz = P->offset.empty() ? 0 : P->offset[(size_t)C];
	if (nnz == 0) return FH_HOST_OK;
	if (!indices || !values) return fail(FH_HOST_EINVAL, "NULL output");
	Tracer trace;
	int32_t* row = indices;
	int32_t* col = indices + nnz;
	int32_t* cell = indices + 2 * nnz;
	const int threads = P->threads;
	double sum = 0.0;
#pragma omp parallel for num_threads(threads) schedule(static) reduction(+ : sum)
	for (int64_t c = 0; c < C; ++c) {
		int64_t k = P->offset[(size_t)c];
		for (const Entry& x : P->ent[(size_t)c]) {
			row[k] = x.r; col[k] = x.c; cell[k] = (int32_t)c;
			const float f = std::log1p((float)x.v);  // fp32 log1p like numpy on the fp32 array (:353)
			values[k] = f;
			sum += (double)f;
			++k;
		}
	}
	const double mean = sum / (double)nnz;
	double var = 0.0;
#pragma omp parallel for num_threads(threads) schedule(static) reduction(+ : var)
	for (int64_t k = 0; k < nnz; ++k) {
		const double d = (double)values[k] - mean;
		var += d * d;
	}
	const float cap = (float)(mean + 15.0 * std::sqrt(var / (double)nnz));
#pragma omp parallel for num_threads(threads) schedule(static)
	for (int64_t k = 0; k < nnz; ++k)
		if (values[k] > cap) values[k] = cap;
	trace.mark("flatten + log1p + clip");
	return FH_HOST_OK;
}

extern "C" void fh_host_pack_free(void* handle) { delete (Packed*)handle; }

// ---------------------------------------------------------------------------------------------------------
// Block-CSR staging (include/fh_host.h): COO (row, col, cell) -> per bin-block CSR over (cell, local row) with
// window-local int16 columns. Two linear passes with per-row atomic counters / cursors instead of a sort of
// nnz 64-bit keys; the entries of a row are then put in ascending column order (rows are short, and already in
// order when the input is the pack stage's), which makes the result independent of the thread schedule.
namespace {

struct StageError {
	int kind = 0;          // 1 = outside window, 2 = index out of range, 3 = duplicate
	int64_t at = -1;       // entry (kinds 1, 2) or flattened (block, row) id (kind 3)
	void set(int k, int64_t where) {
#pragma omp critical(fh_stage_error)
		if (kind == 0 || where < at) { kind = k; at = where; }
	}
};

int check_geom(const fh_host_block_geom* g) {
	if (!g) return fail(FH_HOST_EINVAL, "geometry == NULL");
	if (g->num_bin <= 0 || g->bs_bin <= 0 || g->num_cell <= 0 || g->num_block <= 0 || !g->nb || !g->col0 || !g->w)
		return fail(FH_HOST_EINVAL, "bad block geometry (num_bin %d, bs_bin %d, blocks %d, cells %lld)", g->num_bin, g->bs_bin,
		            g->num_block, (long long)g->num_cell);
	if ((int64_t)g->num_block * g->bs_bin < g->num_bin) return fail(FH_HOST_EINVAL, "blocks do not cover the %d bins", g->num_bin);
	for (int b = 0; b < g->num_block; ++b) {
		if (g->nb[b] <= 0 || g->nb[b] > g->bs_bin || g->w[b] <= 0 || g->col0[b] < 0)
			return fail(FH_HOST_EINVAL, "bad geometry of block %d", b);
		if (g->w[b] > 32767) return fail(FH_HOST_EINVAL, "window of block %d is %d columns wide: int16 column ids hold 32767", b, g->w[b]);
		if ((int64_t)g->num_cell * g->nb[b] + 1 > INT32_MAX) return fail(FH_HOST_EINVAL, "cells x rows of block %d exceeds int32", b);
	}
	return FH_HOST_OK;
}

}  // namespace

extern "C" int fh_host_block_csr_count(const void* row, const void* col, const void* cell, int32_t index_type, int64_t nnz,
                                       const fh_host_block_geom* g, int32_t* const* rowptr, int64_t* nnz_block,
                                       int32_t num_threads) {
	if (int rc = check_geom(g)) return rc;
	if (index_type != FH_HOST_I32 && index_type != FH_HOST_I64) return fail(FH_HOST_EINVAL, "index_type must be I32 or I64");
	if (nnz < 0 || (nnz > 0 && (!row || !col || !cell)) || !rowptr || !nnz_block) return fail(FH_HOST_EINVAL, "NULL argument");
	const int threads = thread_count(num_threads);
	Tracer trace;
	const int B = g->num_block;
	for (int b = 0; b < B; ++b) {
		if (!rowptr[b]) return fail(FH_HOST_EINVAL, "rowptr[%d] == NULL", b);
		const int64_t rows = g->num_cell * g->nb[b] + 1;
		int32_t* rp = rowptr[b];
#pragma omp parallel for num_threads(threads) schedule(static)
		for (int64_t i = 0; i < rows; ++i) rp[i] = 0;
	}
	StageError err;
#pragma omp parallel for num_threads(threads) schedule(static)
	for (int64_t k = 0; k < nnz; ++k) {
		const int64_t r = load_index(row, k, index_type), c = load_index(col, k, index_type), z = load_index(cell, k, index_type);
		if (r < 0 || r >= g->num_bin || z < 0 || z >= g->num_cell || c < 0) { err.set(2, k); continue; }
		const int b = (int)(r / g->bs_bin);
		const int64_t lc = c - g->col0[b];
		if (lc < 0 || lc >= g->w[b]) { err.set(1, k); continue; }
		int32_t* slot = rowptr[b] + (z * g->nb[b] + (r - (int64_t)b * g->bs_bin)) + 1;
		__atomic_fetch_add(slot, 1, __ATOMIC_RELAXED);
	}
	if (err.kind == 2) return fail(FH_HOST_EINVAL, "entry %lld: index outside the (%d, *, %lld) tensor", (long long)err.at, g->num_bin, (long long)g->num_cell);
	if (err.kind == 1) return fail(FH_HOST_EWINDOW, "entry %lld lies outside the window of its bin block", (long long)err.at);
	trace.mark("block-CSR count");
	int bad_block = -1;
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
	for (int b = 0; b < B; ++b) {
		int32_t* rp = rowptr[b];
		const int64_t rows = g->num_cell * g->nb[b];
		int64_t run = 0;
		for (int64_t i = 1; i <= rows; ++i) {
			run += rp[i];
			rp[i] = (int32_t)run;   // checked below: a block holds fewer than 2^31 entries
		}
		nnz_block[b] = run;
		if (run > INT32_MAX) {
#pragma omp critical(fh_stage_error)
			bad_block = b;
		}
	}
	if (bad_block >= 0) return fail(FH_HOST_EINVAL, "block %d holds %lld entries: exceeds int32", bad_block, (long long)nnz_block[bad_block]);
	trace.mark("block-CSR prefix sums");
	return FH_HOST_OK;
}

extern "C" int fh_host_block_csr_fill(const void* row, const void* col, const void* cell, int32_t index_type, const float* val,
                                      int64_t nnz, const fh_host_block_geom* g, int32_t* const* rowptr, int16_t* const* col_out,
                                      float* const* val_out, int32_t num_threads) {
	if (int rc = check_geom(g)) return rc;
	if (index_type != FH_HOST_I32 && index_type != FH_HOST_I64) return fail(FH_HOST_EINVAL, "index_type must be I32 or I64");
	if (nnz < 0 || (nnz > 0 && (!row || !col || !cell || !val)) || !rowptr || !col_out || !val_out) return fail(FH_HOST_EINVAL, "NULL argument");
	const int threads = thread_count(num_threads);
	Tracer trace;
	const int B = g->num_block;
	for (int b = 0; b < B; ++b) {
		const int64_t rows = g->num_cell * g->nb[b];
		if (!rowptr[b]) return fail(FH_HOST_EINVAL, "rowptr[%d] == NULL", b);
		if (rowptr[b][rows] > 0 && (!col_out[b] || !val_out[b])) return fail(FH_HOST_EINVAL, "NULL output of block %d", b);
	}
	// scatter: rowptr[b][i] is the cursor of row i; afterwards it equals the original rowptr[b][i + 1]
	StageError err;
#pragma omp parallel for num_threads(threads) schedule(static)
	for (int64_t k = 0; k < nnz; ++k) {
		const int64_t r = load_index(row, k, index_type), c = load_index(col, k, index_type), z = load_index(cell, k, index_type);
		if (r < 0 || r >= g->num_bin || z < 0 || z >= g->num_cell) { err.set(2, k); continue; }
		const int b = (int)(r / g->bs_bin);
		const int64_t lc = c - g->col0[b];
		if (lc < 0 || lc >= g->w[b]) { err.set(1, k); continue; }
		const int64_t rid = z * g->nb[b] + (r - (int64_t)b * g->bs_bin);
		const int32_t pos = __atomic_fetch_add(rowptr[b] + rid, 1, __ATOMIC_RELAXED);
		// the last pointer of a block is never a cursor: it bounds every write even if the pointers are not this input's
		if (pos < 0 || pos >= rowptr[b][g->num_cell * g->nb[b]]) { err.set(5, k); continue; }
		col_out[b][pos] = (int16_t)lc;
		val_out[b][pos] = val[k];
	}
	// restore the row pointers (cursor i ended at the start of row i + 1) and verify that every row was filled exactly
	int mismatch = 0;
	for (int b = 0; b < B; ++b) {
		int32_t* rp = rowptr[b];
		const int64_t rows = g->num_cell * g->nb[b];
		const int32_t total = rp[rows];
		// rp[rows] was never a cursor: total is intact. Shift right by one.
		int32_t prev = 0;
		for (int64_t i = 0; i < rows; ++i) {
			const int32_t ended = rp[i];
			rp[i] = prev;
			prev = ended;
		}
		if (prev != total) mismatch = 1;
	}
	if (err.kind == 2) return fail(FH_HOST_EINVAL, "entry %lld: index outside the tensor", (long long)err.at);
	if (err.kind == 1) return fail(FH_HOST_EWINDOW, "entry %lld lies outside the window of its bin block", (long long)err.at);
	if (mismatch || err.kind == 5) return fail(FH_HOST_EINVAL, "row pointers do not match the entries (not the output of fh_host_block_csr_count for this input)");
	trace.mark("block-CSR scatter");
	// ascending columns inside every row; equal neighbours are duplicates of one (row, col, cell)
	for (int b = 0; b < B; ++b) {
		const int32_t* rp = rowptr[b];
		const int64_t rows = g->num_cell * g->nb[b];
		int16_t* cb = col_out[b];
		float* vb = val_out[b];
#pragma omp parallel for num_threads(threads) schedule(static, 4096)
		for (int64_t i = 0; i < rows; ++i) {
			const int32_t lo = rp[i], hi = rp[i + 1];
			if (hi < lo) { err.set(4, i); continue; }
			bool sorted = true;
			for (int32_t p = lo + 1; p < hi; ++p)
				if (cb[p] <= cb[p - 1]) { sorted = false; break; }
			if (sorted) continue;
			for (int32_t p = lo + 1; p < hi; ++p) {   // insertion sort: rows hold at most a window of entries
				const int16_t c = cb[p];
				const float v = vb[p];
				int32_t q = p;
				while (q > lo && cb[q - 1] > c) { cb[q] = cb[q - 1]; vb[q] = vb[q - 1]; --q; }
				cb[q] = c; vb[q] = v;
			}
			for (int32_t p = lo + 1; p < hi; ++p)
				if (cb[p] == cb[p - 1]) { err.set(3, (int64_t)b * ((int64_t)1 << 40) + i); break; }
		}
	}
	if (err.kind == 4) return fail(FH_HOST_EINVAL, "row pointers are not monotone");
	if (err.kind == 3) {
		const int b = (int)(err.at >> 40);
		const int64_t i = err.at & (((int64_t)1 << 40) - 1);
		return fail(FH_HOST_EDUP, "duplicate (row, col, cell) entries: block %d, cell %lld, local row %lld", b, (long long)(i / g->nb[b]),
		            (long long)(i % g->nb[b]));
	}
	trace.mark("block-CSR row order");
	return FH_HOST_OK;
}
