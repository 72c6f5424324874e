// C-ABI plumbing shared by all translation units: error reporting, launch accounting, device probe.
#include <stdio.h>
#include <string.h>

#include "ssf_common.cuh"

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;

int ssf_set_error(cudaError_t e) {
    snprintf(g_err, sizeof(g_err), "CUDA error %d: %s", (int)e, cudaGetErrorString(e));
    return SSF_ERR_CUDA;
}

int ssf_arg_error(const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return SSF_ERR_ARG;
}

void ssf_count_launch() { ++g_launches; }

extern "C" const char* ssf_last_error(void) { return g_err; }

extern "C" unsigned long long ssf_launch_count(void) { return g_launches; }

extern "C" int ssf_abi_version(void) { return 1; }

// Fails (non-zero) unless the current device is an sm_100 part: there is no fallback path.
extern "C" int ssf_require_device(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return ssf_set_error(e);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return ssf_set_error(e);
    if (prop.major != 10) {
        snprintf(g_err, sizeof(g_err), "ssf_b200 kernels are built for sm_100a only; device is sm_%d%d", prop.major, prop.minor);
        return SSF_ERR_ARG;
    }
    return SSF_OK;
}
