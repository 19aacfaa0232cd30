// Library-level entry points: version, error strings, device probe.
#include "common.cuh"

extern "C" int pvsg_version(void) { return PVSG_VERSION; }

extern "C" const char* pvsg_error_string(int code) {
    switch (code) {
        case PVSG_OK: return "ok";
        case PVSG_ERR_INVALID_ARG: return "invalid argument (null pointer, non-positive size or misaligned stride)";
        case PVSG_ERR_UNSUPPORTED: return "unsupported shape for this kernel";
        case PVSG_ERR_LAUNCH: return "CUDA launch failed";
        case PVSG_ERR_NO_DEVICE: return "no CUDA device";
        default: return "unknown pvsg error";
    }
}

extern "C" int pvsg_device_info(int device, int* sm_count, int* cc_major, int* cc_minor) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
        cudaGetLastError();
        return PVSG_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PVSG_ERR_NO_DEVICE;
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return PVSG_OK;
}
