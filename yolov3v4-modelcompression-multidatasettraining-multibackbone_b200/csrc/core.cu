// Library-wide bits: status strings, last-error slot, device query.
#include "b200yolo.h"
#include "common.cuh"

static thread_local int g_last_cuda_error = 0;

extern "C" void b2y_set_last_cuda_error(int e) { g_last_cuda_error = e; }
extern "C" int b2y_last_cuda_error(void) { return g_last_cuda_error; }
extern "C" int b2y_abi_version(void) { return B2Y_ABI_VERSION; }

extern "C" const char* b2y_strerror(int status) {
    switch (status) {
        case B2Y_OK: return "ok";
        case B2Y_ERR_INVALID: return "invalid argument (shape/alignment/null pointer)";
        case B2Y_ERR_CUDA: return "CUDA runtime error (see b2y_last_cuda_error)";
        case B2Y_ERR_UNSUPPORTED: return "unsupported configuration for the sm_100a kernels";
        case B2Y_ERR_DRIVER: return "CUDA driver entry point unavailable or tensor-map encode failed";
        default: return "unknown status";
    }
}

extern "C" int b2y_device_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return n;
}
