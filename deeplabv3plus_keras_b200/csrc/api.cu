// Error plumbing and library identity for the C-ABI (include/dlv3p.h).
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace dlv3p {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
        return DLV3P_ERR_CUDA;
    }
    return DLV3P_OK;
}

static int g_pdl = 1;
bool pdl_enabled() { return g_pdl != 0; }

}  // namespace dlv3p

extern "C" const char* dlv3p_last_error(void) { return dlv3p::g_err; }
extern "C" int dlv3p_version(void) { return 200; }
extern "C" int dlv3p_set_pdl(int enabled) { const int was = dlv3p::g_pdl; dlv3p::g_pdl = enabled ? 1 : 0; return was; }
extern "C" int dlv3p_device_arch(void) {
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
        dlv3p::set_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
        return DLV3P_ERR_CUDA;
    }
    return prop.major * 10 + prop.minor;
}
