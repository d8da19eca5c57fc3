// C-ABI plumbing shared by all entry points: error string, version, device query.
#include "common.cuh"
#include <string.h>

static thread_local char g_err[1024] = "";

extern "C" {

void osb_set_error(const char* msg) {
    strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
    g_err[sizeof(g_err) - 1] = 0;
}

const char* osb_last_error(void) { return g_err; }

int osb_abi_version(void) { return 1; }

static long long g_launches = 0;
void osb_count_launch(void) { ++g_launches; }
// number of kernels this library has launched in this process (every launch site counts itself)
long long osb_launch_count(void) { return g_launches; }

// Fills sm_count / cc_major / cc_minor of `device`; fails loudly when there is no CUDA device.
int osb_device_info(int device, int* sm_count, int* cc_major, int* cc_minor) {
    cudaDeviceProp prop;
    OSB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return OSB_OK;
}

}  // extern "C"
