// runtime.cu — error reporting, launch accounting, version.  No device memory is ever owned here.
#include "common.cuh"
#include <atomic>
#include <mutex>

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

// process-wide (set by the step engine around its phases; read by launches from autograd's device threads)
static std::atomic<int> g_wgrad_share{100}, g_main_share{100};
int wgrad_sm_share() { return g_wgrad_share.load(std::memory_order_relaxed); }
int main_sm_share() { return g_main_share.load(std::memory_order_relaxed); }

int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev < 0 || dev >= TG_MAX_DEVICES) ? 0 : dev;
}

int num_sms() {
    static std::atomic<int> n[TG_MAX_DEVICES];
    const int dev = current_device();
    int v = n[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        if (v <= 0) v = 148;
        n[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

void raise_smem_limit(const void* kernel, SmemLimit& lim, int dev, size_t bytes) {
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if ((int)bytes > lim.cur[dev].load(std::memory_order_relaxed)) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        lim.cur[dev].store((int)bytes, std::memory_order_release);
    }
}

}  // namespace tamgcn

extern "C" int tamgcn_set_wgrad_sm_share(int percent) {
    const int old = tamgcn::g_wgrad_share.load();
    if (percent >= 1 && percent <= 100) tamgcn::g_wgrad_share.store(percent);
    return old;
}
extern "C" int tamgcn_set_main_sm_share(int percent) {
    const int old = tamgcn::g_main_share.load();
    if (percent >= 1 && percent <= 100) tamgcn::g_main_share.store(percent);
    return old;
}
extern "C" int tamgcn_version(void) { return 100; }
extern "C" const char* tamgcn_last_error(void) { return tamgcn::g_err; }
extern "C" int64_t tamgcn_launch_count(void) { return (int64_t)tamgcn::g_launches.load(); }
