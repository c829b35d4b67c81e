// Error state, version, launch counter and optional per-kernel CUDA-event timing of libfh_b200.so.
#include <stdarg.h>
#include <atomic>
#include <mutex>
#include <vector>
#include "fh_common.cuh"
#include "../../include/fh_b200.h"

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};
static std::atomic<long long> g_tc_fallbacks{0};

extern "C" void fh_set_error(const char* fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}
extern "C" void fh_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" const char* fh_last_error(void) { return g_err; }
extern "C" int fh_version(void) { return 100; }
extern "C" long long fh_launch_count(void) { return g_launches.load(); }
extern "C" void fh_count_tc_fallback(void) { g_tc_fallbacks.fetch_add(1, std::memory_order_relaxed); }
extern "C" long long fh_tc_fallback_count(void) { return g_tc_fallbacks.load(); }

// ---- per-kernel timing (bench.py's roofline: average launch duration of a kernel measured with CUDA events on the
// stream it is launched on). Off by default: the launch sites then only test one flag.
namespace {
struct TimeRec { int id; cudaEvent_t a, b; };
std::vector<TimeRec> g_recs;
std::mutex g_rec_mu;
std::atomic<int> g_timing{0};
}  // namespace
extern "C" void fh_timing_enable(int on) {
	std::lock_guard<std::mutex> lk(g_rec_mu);
	for (auto& r : g_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
	g_recs.clear();
	g_timing.store(on ? 1 : 0);
}
extern "C" int fh_time_begin(int id, void* stream) {
	if (!g_timing.load(std::memory_order_relaxed)) return -1;
	TimeRec r; r.id = id;
	if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return -1;
	cudaEventRecord(r.a, (cudaStream_t)stream);
	std::lock_guard<std::mutex> lk(g_rec_mu);
	g_recs.push_back(r);
	return (int)g_recs.size() - 1;
}
extern "C" void fh_time_end(int idx, void* stream) {
	if (idx < 0) return;
	std::lock_guard<std::mutex> lk(g_rec_mu);
	if (idx < (int)g_recs.size()) cudaEventRecord(g_recs[idx].b, (cudaStream_t)stream);
}
extern "C" int fh_timing_read(int id, double* total_ms, long long* launches) {
	std::lock_guard<std::mutex> lk(g_rec_mu);
	double t = 0.0; long long n = 0;
	for (auto& r : g_recs) {
		if (r.id != id) continue;
		FH_CUDA(cudaEventSynchronize(r.b));
		float ms = 0.f;
		FH_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
		t += ms; ++n;
	}
	if (total_ms) *total_ms = t;
	if (launches) *launches = n;
	return FH_OK;
}
