// LayerNorm, GroupNorm (token-major), broadcast add.
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

// One warp per row; the row lives in registers (C = 32 * 4 * V floats).
template <int V>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta,
                                                        float* __restrict__ y, int64_t rows, float eps,
                                                        uint16_t* __restrict__ y_hi = nullptr,
                                                        uint16_t* __restrict__ y_lo = nullptr,
                                                        const float* __restrict__ add = nullptr,
                                                        uint16_t* __restrict__ s_hi = nullptr,
                                                        uint16_t* __restrict__ s_lo = nullptr) {
    constexpr int C = 128 * V;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + row * C);
    float4 v[V];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        v[i] = xr[lane + 32 * i];
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + eps);
    float4* yr = reinterpret_cast<float4*>(y + row * C);
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + lane + 32 * i);
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + lane + 32 * i);
        float4 o;
        o.x = (v[i].x - mean) * rstd * g.x + b.x;
        o.y = (v[i].y - mean) * rstd * g.y + b.y;
        o.z = (v[i].z - mean) * rstd * g.z + b.z;
        o.w = (v[i].w - mean) * rstd * g.w + b.w;
        yr[lane + 32 * i] = o;
        if (y_hi) {  // split-bf16 planes of the same values for the tcgen05 engine
            const float a[4] = {o.x, o.y, o.z, o.w};
            uint16_t h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const __nv_bfloat16 hh = __float2bfloat16_rn(a[e]);
                h[e] = __bfloat16_as_ushort(hh);
                l[e] = __bfloat16_as_ushort(__float2bfloat16_rn(a[e] - __bfloat162float(hh)));
            }
            const int64_t q = row * (C / 4) + lane + 32 * i;
            reinterpret_cast<uint2*>(y_hi)[q] = make_uint2(h[0] | ((uint32_t)h[1] << 16), h[2] | ((uint32_t)h[3] << 16));
            reinterpret_cast<uint2*>(y_lo)[q] = make_uint2(l[0] | ((uint32_t)l[1] << 16), l[2] | ((uint32_t)l[3] << 16));
        }
        if (s_hi) {  // planes of y + add (the query = x + pos operand of the next attention's projections)
            const int64_t q = row * (C / 4) + lane + 32 * i;
            const float4 p4 = __ldg(reinterpret_cast<const float4*>(add) + q);
            const float a[4] = {o.x + p4.x, o.y + p4.y, o.z + p4.z, o.w + p4.w};
            uint16_t h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const __nv_bfloat16 hh = __float2bfloat16_rn(a[e]);
                h[e] = __bfloat16_as_ushort(hh);
                l[e] = __bfloat16_as_ushort(__float2bfloat16_rn(a[e] - __bfloat162float(hh)));
            }
            reinterpret_cast<uint2*>(s_hi)[q] = make_uint2(h[0] | ((uint32_t)h[1] << 16), h[2] | ((uint32_t)h[3] << 16));
            reinterpret_cast<uint2*>(s_lo)[q] = make_uint2(l[0] | ((uint32_t)l[1] << 16), l[2] | ((uint32_t)l[3] << 16));
        }
    }
}

// GroupNorm pass 1: each CTA reduces a slab of pixels; thread t owns channel quad
// (t % (C/4)); partial (sum, sumsq) per group are pushed with double atomics.
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ x, double* __restrict__ stats,
                                                       int64_t HW, int C, int groups, int pix_per_cta) {
    const int b = blockIdx.y;
    const int cq = C >> 2;                 // float4 per pixel
    const int lanes_per_pix = cq;          // threads covering one pixel
    const int pix_par = blockDim.x / lanes_per_pix;
    const int q = threadIdx.x % lanes_per_pix;
    const int pl = threadIdx.x / lanes_per_pix;
    const int64_t p0 = (int64_t)blockIdx.x * pix_per_cta;
    const int64_t p1 = min(p0 + (int64_t)pix_per_cta, HW);
    float s = 0.f, ss = 0.f;
    if (pl < pix_par) {
        const float4* xb = reinterpret_cast<const float4*>(x + (int64_t)b * HW * C);
        for (int64_t p = p0 + pl; p < p1; p += pix_par) {
            float4 v = __ldg(xb + p * cq + q);
            s += (v.x + v.y) + (v.z + v.w);
            ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        }
    }
    // channels per group = C / groups; quads per group = cpg / 4 (>= 1 required)
    const int qpg = (C / groups) >> 2;
    __shared__ double sh[2][64];
    if (threadIdx.x < 2 * 64) (&sh[0][0])[threadIdx.x] = 0.0;
    __syncthreads();
    if (pl < pix_par) {
        const int g = q / qpg;
        atomicAdd(&sh[0][g], (double)s);
        atomicAdd(&sh[1][g], (double)ss);
    }
    __syncthreads();
    if (threadIdx.x < groups) {
        atomicAdd(&stats[((int64_t)b * groups + threadIdx.x) * 2 + 0], sh[0][threadIdx.x]);
        atomicAdd(&stats[((int64_t)b * groups + threadIdx.x) * 2 + 1], sh[1][threadIdx.x]);
    }
}

__global__ void __launch_bounds__(256) gn_apply_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, const double* __restrict__ stats,
                                                       float* __restrict__ y, int64_t HW, int C, int groups,
                                                       float eps, int act, int64_t total_quads,
                                                       uint16_t* __restrict__ y_hi = nullptr,
                                                       uint16_t* __restrict__ y_lo = nullptr) {
    const int cq = C >> 2;
    const int qpg = (C / groups) >> 2;
    const double cnt = (double)HW * (C / groups);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_quads;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(i % cq);
        const int64_t b = i / (cq * HW);
        const int g = q / qpg;
        const double m = stats[(b * groups + g) * 2 + 0] / cnt;
        const double var = stats[(b * groups + g) * 2 + 1] / cnt - m * m;
        const float mean = (float)m;
        const float rstd = rsqrtf(fmaxf((float)var, 0.f) + eps);
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + q);
        const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + q);
        float4 o;
        o.x = (v.x - mean) * rstd * ga.x + be.x;
        o.y = (v.y - mean) * rstd * ga.y + be.y;
        o.z = (v.z - mean) * rstd * ga.z + be.z;
        o.w = (v.w - mean) * rstd * ga.w + be.w;
        if (act == PVSG_ACT_RELU) {
            o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
        }
        if (y) reinterpret_cast<float4*>(y)[i] = o;
        if (y_hi) {
            const float a[4] = {o.x, o.y, o.z, o.w};
            uint16_t h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const __nv_bfloat16 hh = __float2bfloat16_rn(a[e]);
                h[e] = __bfloat16_as_ushort(hh);
                l[e] = __bfloat16_as_ushort(__float2bfloat16_rn(a[e] - __bfloat162float(hh)));
            }
            reinterpret_cast<uint2*>(y_hi)[i] = make_uint2(h[0] | ((uint32_t)h[1] << 16), h[2] | ((uint32_t)h[3] << 16));
            reinterpret_cast<uint2*>(y_lo)[i] = make_uint2(l[0] | ((uint32_t)l[1] << 16), l[2] | ((uint32_t)l[3] << 16));
        }
    }
}

__global__ void __launch_bounds__(256) add_rowvec_kernel(const float* __restrict__ x, const float* __restrict__ v,
                                                         float* __restrict__ y, int64_t total_quads, int cq) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_quads;
         i += (int64_t)gridDim.x * blockDim.x) {
        float4 a = __ldg(reinterpret_cast<const float4*>(x) + i);
        const float4 b = __ldg(reinterpret_cast<const float4*>(v) + (i % cq));
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        reinterpret_cast<float4*>(y)[i] = a;
    }
}

}  // namespace

static int layernorm_launch(const float* x, const float* gamma, const float* beta, float* y, int64_t rows, int C,
                            float eps, uint16_t* hi, uint16_t* lo, void* stream, const float* add = nullptr,
                            uint16_t* s_hi = nullptr, uint16_t* s_lo = nullptr) {
    const int wpb = 8;
    dim3 grid((unsigned)((rows + wpb - 1) / wpb));
    cudaStream_t st = as_stream(stream);
    switch (C) {
        case 128: layernorm_kernel<1><<<grid, wpb * 32, 0, st>>>(x, gamma, beta, y, rows, eps, hi, lo, add, s_hi, s_lo); break;
        case 256: layernorm_kernel<2><<<grid, wpb * 32, 0, st>>>(x, gamma, beta, y, rows, eps, hi, lo, add, s_hi, s_lo); break;
        case 512: layernorm_kernel<4><<<grid, wpb * 32, 0, st>>>(x, gamma, beta, y, rows, eps, hi, lo, add, s_hi, s_lo); break;
        case 1024: layernorm_kernel<8><<<grid, wpb * 32, 0, st>>>(x, gamma, beta, y, rows, eps, hi, lo, add, s_hi, s_lo); break;
        default: return PVSG_ERR_UNSUPPORTED;
    }
    return pvsg_launch_status();
}

extern "C" int pvsg_layernorm(const float* x, const float* gamma, const float* beta, float* y,
                              int64_t rows, int C, float eps, void* stream) {
    PVSG_CHECK_ARG(x && gamma && beta && y && rows > 0);
    return layernorm_launch(x, gamma, beta, y, rows, C, eps, nullptr, nullptr, stream);
}

extern "C" int pvsg_layernorm_split(const float* x, const float* gamma, const float* beta, float* y, void* y_hi,
                                    void* y_lo, int64_t rows, int C, float eps, void* stream) {
    PVSG_CHECK_ARG(x && gamma && beta && y && y_hi && y_lo && rows > 0);
    return layernorm_launch(x, gamma, beta, y, rows, C, eps, reinterpret_cast<uint16_t*>(y_hi),
                            reinterpret_cast<uint16_t*>(y_lo), stream);
}

extern "C" int pvsg_layernorm_split2(const float* x, const float* gamma, const float* beta, float* y, void* y_hi,
                                     void* y_lo, const float* add, void* s_hi, void* s_lo, int64_t rows, int C,
                                     float eps, void* stream) {
    PVSG_CHECK_ARG(x && gamma && beta && y && y_hi && y_lo && add && s_hi && s_lo && rows > 0);
    return layernorm_launch(x, gamma, beta, y, rows, C, eps, reinterpret_cast<uint16_t*>(y_hi),
                            reinterpret_cast<uint16_t*>(y_lo), stream, add, reinterpret_cast<uint16_t*>(s_hi),
                            reinterpret_cast<uint16_t*>(s_lo));
}

static int groupnorm_launch(const float* x, const float* gamma, const float* beta, float* y, uint16_t* y_hi,
                            uint16_t* y_lo, double* stats, int B, int64_t HW, int C, int groups, float eps, int act,
                            void* stream) {
    PVSG_CHECK_ARG(x && gamma && beta && (y || y_hi) && stats && B > 0 && HW > 0 && C > 0 && groups > 0);
    PVSG_CHECK_ARG(C % groups == 0 && (C / groups) % 4 == 0 && groups <= 64 && C / 4 <= 256);
    cudaStream_t st = as_stream(stream);
    if (cudaMemsetAsync(stats, 0, sizeof(double) * 2 * (size_t)B * groups, st) != cudaSuccess)
        return PVSG_ERR_LAUNCH;
    const int pix_per_cta = 128;
    dim3 g1((unsigned)((HW + pix_per_cta - 1) / pix_per_cta), (unsigned)B);
    gn_stats_kernel<<<g1, 256, 0, st>>>(x, stats, HW, C, groups, pix_per_cta);
    const int64_t total = (int64_t)B * HW * (C / 4);
    const unsigned g2 = (unsigned)imin64((total + 255) / 256, 148 * 16);
    gn_apply_kernel<<<g2, 256, 0, st>>>(x, gamma, beta, stats, y, HW, C, groups, eps, act, total, y_hi, y_lo);
    return pvsg_launch_status();
}

extern "C" int pvsg_groupnorm_nhwc(const float* x, const float* gamma, const float* beta, float* y,
                                   double* stats, int B, int64_t HW, int C, int groups, float eps,
                                   int act, void* stream) {
    return groupnorm_launch(x, gamma, beta, y, nullptr, nullptr, stats, B, HW, C, groups, eps, act, stream);
}

extern "C" int pvsg_groupnorm_nhwc_split(const float* x, const float* gamma, const float* beta, float* y, void* y_hi,
                                         void* y_lo, double* stats, int B, int64_t HW, int C, int groups, float eps,
                                         int act, void* stream) {
    PVSG_CHECK_ARG(y_hi && y_lo);
    return groupnorm_launch(x, gamma, beta, y, reinterpret_cast<uint16_t*>(y_hi), reinterpret_cast<uint16_t*>(y_lo),
                            stats, B, HW, C, groups, eps, act, stream);
}

extern "C" int pvsg_add_rowvec(const float* x, const float* v, float* y, int64_t rows, int C, void* stream) {
    PVSG_CHECK_ARG(x && v && y && rows > 0 && C > 0 && C % 4 == 0);
    const int64_t total = rows * (C / 4);
    const unsigned g = (unsigned)imin64((total + 255) / 256, 148 * 16);
    add_rowvec_kernel<<<g, 256, 0, as_stream(stream)>>>(x, v, y, total, C / 4);
    return pvsg_launch_status();
}
