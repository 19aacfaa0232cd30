// Shared helpers for libpvsg_sm100.so kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/pvsg.h"

#define PVSG_CHECK_ARG(cond) \
    do {                     \
        if (!(cond)) return PVSG_ERR_INVALID_ARG; \
    } while (0)

static inline int pvsg_launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? PVSG_OK : PVSG_ERR_LAUNCH;
}

// Function attributes (dynamic shared-memory opt-in, carve-out) are PER DEVICE: a process driving several GPUs
// must configure each kernel once on every device it launches on.  `flags` is a zero-initialised static array.
constexpr int PVSG_MAX_DEVICES = 64;
static inline bool pvsg_first_use_on_device(bool (&flags)[PVSG_MAX_DEVICES]) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= PVSG_MAX_DEVICES) dev = 0;
    const bool first = !flags[dev];
    flags[dev] = true;
    return first;
}

// Every translation unit with opt-in kernel attributes exposes configure_<unit>(): sets them for the CURRENT device
// (idempotent).  Launchers call them lazily; pvsg_create calls all of them up front so that the first launch of a
// kernel may happen inside a stream capture (cudaFuncSetAttribute is not a stream operation, but doing it eagerly keeps
// captures free of first-use side effects and surfaces an unsupported device at create time).
namespace pvsg_internal {
int configure_attention_mma();
int configure_attention_t5();
int configure_gemm_tc();
int configure_gemm_skinny();
int configure_msda_tile();
int configure_overlap();
int configure_panoptic();
int configure_relation();
int configure_swin();
}  // namespace pvsg_internal

static inline int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Source index / weight of one output coordinate for bilinear, align_corners=False
// (ATen area_pixel_compute_source_index + guard): src = scale*(dst+0.5)-0.5 clamped at 0.
__device__ __forceinline__ void bilinear_coord(int dst, float scale, int in_size, int& i0, int& i1,
                                               float& w0, float& w1) {
    float src = scale * (dst + 0.5f) - 0.5f;
    src = src < 0.f ? 0.f : src;
    i0 = (int)src;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    w1 = src - (float)i0;
    w0 = 1.f - w1;
}

// exact GELU, torch nn.GELU(approximate='none'): 0.5 x (1 + erf(x / sqrt 2))
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
