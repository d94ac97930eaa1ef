// Error reporting and device queries of libsg_b200.so.
#include <stdarg.h>

#include "sg_common.cuh"

namespace sg {

static thread_local char g_err[512] = "";

char *err_buf() { return g_err; }

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace sg

extern "C" {

const char *sg_last_error(void) { return sg::err_buf(); }

int sg_abi_version(void) { return 3; }

int sg_device_info(int *sm_count, int *smem_optin_bytes, int *l2_bytes) {
    int dev = 0;
    SG_CUDA_TRY(cudaGetDevice(&dev));
    if (sm_count) SG_CUDA_TRY(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
    if (smem_optin_bytes)
        SG_CUDA_TRY(cudaDeviceGetAttribute(smem_optin_bytes, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (l2_bytes) SG_CUDA_TRY(cudaDeviceGetAttribute(l2_bytes, cudaDevAttrL2CacheSize, dev));
    return SG_OK;
}

}  // extern "C"
