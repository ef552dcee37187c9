// Library-wide state of librdfc_b200.so: error string, launch counter, device queries.
#include <atomic>

#include "common.cuh"

namespace rdfc {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

char *err_buf() { return g_err; }

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

}  // namespace rdfc

extern "C" int rdfc_abi_version(void) { return RDFC_ABI_VERSION; }
extern "C" const char *rdfc_last_error(void) { return rdfc::err_buf(); }
extern "C" uint64_t rdfc_launch_count(void) { return rdfc::g_launches.load(); }
