// Library-level entry points: version, error strings, device probe.
#include <new>

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

// ------------------------------------------------------------------------------------------ handle
// A handle pins a device, configures every kernel's opt-in attributes on it up front and owns one grow-only scratch
// allocation that callers may use for the `ws` arguments (sized by the pvsg_<op>_workspace_bytes companions).  The op
// entry points themselves stay stateless (plain pointers, sizes and a stream), so a handle is optional.
struct pvsg_handle {
    int device;
    int sm_count;
    int64_t smem_optin;
    void* ws;
    int64_t ws_bytes;
};

namespace {
struct DeviceScope {
    int prev = -1;
    bool ok;
    explicit DeviceScope(int dev) { ok = cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess; }
    ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
};
}  // namespace

extern "C" int pvsg_create(int device, pvsg_handle** out) {
    PVSG_CHECK_ARG(out);
    *out = nullptr;
    int sms = 0, major = 0, minor = 0;
    if (const int rc = pvsg_device_info(device, &sms, &major, &minor)) return rc;
    if (major != 10) return PVSG_ERR_UNSUPPORTED;        // the library holds sm_100a code only
    DeviceScope scope(device);
    if (!scope.ok) return PVSG_ERR_NO_DEVICE;
    int optin = 0;
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    using namespace pvsg_internal;
    int (*const units[])() = {configure_gemm_tc,  configure_gemm_skinny, configure_attention_t5, configure_attention_mma,
                              configure_msda_tile, configure_panoptic,   configure_overlap,      configure_swin,
                              configure_relation};
    for (auto f : units)
        if (const int rc = f()) return rc;
    auto* h = new (std::nothrow) pvsg_handle{device, sms, (int64_t)optin, nullptr, 0};
    if (!h) return PVSG_ERR_LAUNCH;
    *out = h;
    return PVSG_OK;
}

extern "C" int pvsg_destroy(pvsg_handle* h) {
    if (!h) return PVSG_OK;
    int rc = PVSG_OK;
    if (h->ws) {
        DeviceScope scope(h->device);
        if (!scope.ok || cudaFree(h->ws) != cudaSuccess) rc = PVSG_ERR_LAUNCH;
    }
    delete h;
    return rc;
}

extern "C" int pvsg_handle_info(const pvsg_handle* h, int* device, int* sm_count, int64_t* smem_optin_bytes,
                                int64_t* workspace_bytes) {
    PVSG_CHECK_ARG(h);
    if (device) *device = h->device;
    if (sm_count) *sm_count = h->sm_count;
    if (smem_optin_bytes) *smem_optin_bytes = h->smem_optin;
    if (workspace_bytes) *workspace_bytes = h->ws_bytes;
    return PVSG_OK;
}

extern "C" int pvsg_workspace(pvsg_handle* h, int64_t bytes, void** ptr) {
    PVSG_CHECK_ARG(h && ptr && bytes >= 0);
    if (bytes > h->ws_bytes) {
        DeviceScope scope(h->device);
        if (!scope.ok) return PVSG_ERR_NO_DEVICE;
        // growing synchronises the device (cudaFree), like any allocator refill: size it once before a capture
        if (h->ws && cudaFree(h->ws) != cudaSuccess) return PVSG_ERR_LAUNCH;
        h->ws = nullptr;
        h->ws_bytes = 0;
        const int64_t want = (bytes + 255) & ~(int64_t)255;
        if (cudaMalloc(&h->ws, (size_t)want) != cudaSuccess) { cudaGetLastError(); return PVSG_ERR_LAUNCH; }
        h->ws_bytes = want;
    }
    *ptr = h->ws;
    return PVSG_OK;
}
