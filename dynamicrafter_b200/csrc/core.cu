// Error reporting and device checks of the C ABI.
#include <cstdlib>
#include <cstdarg>
#include <cstdio>

#include "dcb_internal.h"

namespace dcb {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev;
}

// per-device caches: a process may drive several GPUs (cuda:0 for one model, cuda:1 for another)
int num_sms() {
    static int n[kMaxDevices] = {};
    const int dev = current_device() & (kMaxDevices - 1);
    if (n[dev] == 0) {
        if (cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, current_device()) != cudaSuccess || n[dev] <= 0) n[dev] = 148;
    }
    return n[dev];
}

// DCB_PDL: bit mask of the kernel classes launched with programmatic stream serialization
// (1 tap_gemm, 2 layer norm, 4 cooperative group norm, 8 attention, 16 group-norm finalize + streaming apply); default
// set by measurement, see DESIGN.md.
bool pdl_enabled(int cls) {
    static int v = -1;
    if (v < 0) { const char* e = getenv("DCB_PDL"); v = (e == nullptr) ? 17 : atoi(e); }
    return (v & cls) != 0;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(DCB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return DCB_OK;
}

}  // namespace dcb

extern "C" const char* dcb_last_error(void) { return dcb::g_err; }
extern "C" int dcb_abi_version(void) { return 1; }

extern "C" int dcb_check_device(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
        return dcb::set_error(DCB_ERR_CUDA, "no CUDA device");
    if (major != 10) return dcb::set_error(DCB_ERR_DEVICE, "dcb200 needs an sm_100 device, found sm_%d0", major);
    return DCB_OK;
}
