// runtime.cu — error reporting, launch accounting, version.  No device memory is ever owned here.
#include "common.cuh"
#include <atomic>

namespace tamgcn {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return -1;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: launch failed: %s", what, cudaGetErrorString(e));
        return -2;
    }
    return 0;
}

}  // namespace tamgcn

extern "C" int tamgcn_version(void) { return 100; }
extern "C" const char* tamgcn_last_error(void) { return tamgcn::g_err; }
extern "C" int64_t tamgcn_launch_count(void) { return (int64_t)tamgcn::g_launches.load(); }
