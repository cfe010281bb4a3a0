// Error reporting and launch accounting for libmeshdqn_b200.so.
#include <stdarg.h>

#include <atomic>

#include "mdq_common.cuh"

namespace mdq {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace mdq

extern "C" {
const char *mdq_last_error(void) { return mdq::g_err; }
int mdq_version(void) { return 100; }
int64_t mdq_launch_count(void) { return mdq::g_launches.load(std::memory_order_relaxed); }
void mdq_launch_count_add(int64_t n) { mdq::g_launches.fetch_add(n, std::memory_order_relaxed); }
}
