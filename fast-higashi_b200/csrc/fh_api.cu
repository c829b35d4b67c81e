// Error state, version and launch counter of libfh_b200.so.
#include <stdarg.h>
#include <atomic>
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
