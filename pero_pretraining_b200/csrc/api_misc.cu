// Version, error strings and the device check of the C ABI.
#include <cuda_runtime.h>
#include "../../include/pero_b200.h"

extern "C" {

int pero_version(void) { return 100; }

const char* pero_strerror(int code) {
    switch (code) {
        case PERO_OK: return "success";
        case PERO_ERR_BAD_SHAPE: return "bad shape or size argument";
        case PERO_ERR_BAD_ALIGN: return "pointer or pitch not aligned as required (256 B for blobs/workspaces, 16 B for operands)";
        case PERO_ERR_WORKSPACE: return "workspace or blob smaller than the *_bytes() query";
        case PERO_ERR_ARCH: return "device is not sm_100 (B200); this library has no other code path";
        case PERO_ERR_NULL: return "required pointer is NULL";
        case PERO_ERR_DRIVER: return "CUDA driver entry point unavailable or tensor-map encode failed";
        case PERO_ERR_UNSUPPORTED: return "unsupported argument combination";
        default: break;
    }
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    return "unknown pero error";
}

int pero_check_device(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return (int)e;
    return major == 10 ? PERO_OK : PERO_ERR_ARCH;
}

}  // extern "C"
