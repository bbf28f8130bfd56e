// lz_core.cu -- error reporting, version and launch accounting for libliuzhou_b200.so.
#include <stdarg.h>

#include "lz_common.cuh"

namespace lzb {

static thread_local char g_error[512] = "";
std::atomic<uint64_t> g_launch_count{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

}  // namespace lzb

extern "C" const char* lzb_last_error(void) { return lzb::g_error; }
extern "C" const char* lzb_version(void) { return "liuzhou_b200 0.1 sm_100a"; }
extern "C" uint64_t lzb_launch_count(void) { return lzb::g_launch_count.load(std::memory_order_relaxed); }
