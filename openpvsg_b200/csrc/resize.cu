// Bilinear resize (token-major), max pooling, NCHW<->NHWC, sine positional encodings.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) bilinear_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                            int IH, int IW, int OH, int OW, int cq,
                                                            float sh, float sw, int accumulate, int64_t total,
                                                            int64_t src_bs, uint2* __restrict__ d_hi = nullptr,
                                                            uint2* __restrict__ d_lo = nullptr) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(i % cq);
        int64_t t = i / cq;
        const int ox = (int)(t % OW);
        t /= OW;
        const int oy = (int)(t % OH);
        const int64_t b = t / OH;
        int y0, y1, x0, x1;
        float wy0, wy1, wx0, wx1;
        bilinear_coord(oy, sh, IH, y0, y1, wy0, wy1);
        bilinear_coord(ox, sw, IW, x0, x1, wx0, wx1);
        const float4* s = reinterpret_cast<const float4*>(src) + b * src_bs;   // batch stride in float4
        const float4 a = __ldg(s + ((int64_t)y0 * IW + x0) * cq + q);
        const float4 bb = __ldg(s + ((int64_t)y0 * IW + x1) * cq + q);
        const float4 c = __ldg(s + ((int64_t)y1 * IW + x0) * cq + q);
        const float4 d = __ldg(s + ((int64_t)y1 * IW + x1) * cq + q);
        float4 o;
        o.x = wy0 * (wx0 * a.x + wx1 * bb.x) + wy1 * (wx0 * c.x + wx1 * d.x);
        o.y = wy0 * (wx0 * a.y + wx1 * bb.y) + wy1 * (wx0 * c.y + wx1 * d.y);
        o.z = wy0 * (wx0 * a.z + wx1 * bb.z) + wy1 * (wx0 * c.z + wx1 * d.z);
        o.w = wy0 * (wx0 * a.w + wx1 * bb.w) + wy1 * (wx0 * c.w + wx1 * d.w);
        float4* dp = reinterpret_cast<float4*>(dst) + i;
        if (accumulate) {
            const float4 e = *dp;
            o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
        }
        *dp = o;
        if (d_hi) {   // split-bf16 operand planes of the result (for the conv that consumes it)
            const float a4[4] = {o.x, o.y, o.z, o.w};
            uint32_t hh[4], ll[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t u = __float_as_uint(a4[e]);
                const uint32_t hb = (u + 0x7fffu + ((u >> 16) & 1u)) & 0xffff0000u;    // RNE to bf16
                const uint32_t ur = __float_as_uint(a4[e] - __uint_as_float(hb));
                hh[e] = hb >> 16;
                ll[e] = ((ur + 0x7fffu + ((ur >> 16) & 1u)) >> 16) & 0xffffu;
            }
            d_hi[i] = make_uint2(hh[0] | (hh[1] << 16), hh[2] | (hh[3] << 16));
            d_lo[i] = make_uint2(ll[0] | (ll[1] << 16), ll[2] | (ll[3] << 16));
        }
    }
}

__global__ void __launch_bounds__(256) maxpool_kernel(const float* __restrict__ x, float* __restrict__ y, int H,
                                                      int W, int OH, int OW, int cq, int64_t total) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(i % cq);
        int64_t t = i / cq;
        const int ox = (int)(t % OW);
        t /= OW;
        const int oy = (int)(t % OH);
        const int64_t b = t / OH;
        const float4* s = reinterpret_cast<const float4*>(x) + b * H * W * (int64_t)cq;
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int iy = oy * 2 - 1 + dy;
            if (iy < 0 || iy >= H) continue;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int ix = ox * 2 - 1 + dx;
                if (ix < 0 || ix >= W) continue;
                const float4 v = __ldg(s + ((int64_t)iy * W + ix) * cq + q);
                m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
            }
        }
        reinterpret_cast<float4*>(y)[i] = m;
    }
}

// Tiled transpose of a [R, Ccols] matrix per batch: in [B, R, Ccols] -> out [B, Ccols, R].
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                        int64_t R, int64_t Cc) {
    __shared__ float tile[32][33];
    const int64_t b = blockIdx.z;
    const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
    const float* ib = in + b * R * Cc;
    float* ob = out + b * R * Cc;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        int64_t r = r0 + j, c = c0 + tx;
        tile[j][tx] = (r < R && c < Cc) ? __ldg(ib + r * Cc + c) : 0.f;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        int64_t c = c0 + j, r = r0 + tx;
        if (r < R && c < Cc) ob[c * R + r] = tile[tx][j];
    }
}

// out[(t*H + y)*W + x, :] = cat(pos_y, pos_x) (+ pos_z) (+ add_vec); channel k of each part:
// even k -> sin(e / dim_t[k]), odd k -> cos(e / dim_t[k]) with e the normalised cumsum.
__global__ void __launch_bounds__(256) sine_pe_kernel(float* __restrict__ out, const float* __restrict__ dim_t,
                                                      const float* __restrict__ dim_t_z,
                                                      const float* __restrict__ add_vec, int T, int H, int W,
                                                      int nf, float scale, float eps, int64_t total) {
    const int C = 2 * nf;
    const int Tn = T > 0 ? T : 1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        int64_t t = i / C;
        const int x = (int)(t % W);
        t /= W;
        const int y = (int)(t % H);
        const int tz = (int)(t / H);
        float e;
        int k;
        if (c < nf) {  // y part first
            e = (float)(y + 1) / ((float)H + eps) * scale;
            k = c;
        } else {
            e = (float)(x + 1) / ((float)W + eps) * scale;
            k = c - nf;
        }
        float arg = e / __ldg(dim_t + k);
        float v = (k & 1) ? cosf(arg) : sinf(arg);
        if (T > 0) {
            const float ez = (float)(tz + 1) / ((float)Tn + eps) * scale;
            const float az = ez / __ldg(dim_t_z + c);
            v += (c & 1) ? cosf(az) : sinf(az);
        }
        if (add_vec) v += __ldg(add_vec + c);
        out[i] = v;
    }
}

inline unsigned grid_for(int64_t total) { return (unsigned)imin64((total + 255) / 256, 148 * 16); }

}  // namespace

extern "C" int pvsg_bilinear_resize_nhwc(const float* src, float* dst, int B, int IH, int IW, int OH,
                                         int OW, int C, int accumulate, void* stream) {
    PVSG_CHECK_ARG(src && dst && B > 0 && IH > 0 && IW > 0 && OH > 0 && OW > 0 && C > 0 && C % 4 == 0);
    const int64_t total = (int64_t)B * OH * OW * (C / 4);
    bilinear_nhwc_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(
        src, dst, IH, IW, OH, OW, C / 4, (float)IH / (float)OH, (float)IW / (float)OW, accumulate, total,
        (int64_t)IH * IW * (C / 4));
    return pvsg_launch_status();
}

extern "C" int pvsg_bilinear_resize_nhwc_ex(const float* src, int64_t src_batch_stride, float* dst, void* dst_hi,
                                            void* dst_lo, int B, int IH, int IW, int OH, int OW, int C,
                                            int accumulate, void* stream) {
    PVSG_CHECK_ARG(src && dst && B > 0 && IH > 0 && IW > 0 && OH > 0 && OW > 0 && C > 0 && C % 4 == 0);
    PVSG_CHECK_ARG((dst_hi == nullptr) == (dst_lo == nullptr) && src_batch_stride % 4 == 0 &&
                   src_batch_stride >= (int64_t)IH * IW * C);
    const int64_t total = (int64_t)B * OH * OW * (C / 4);
    bilinear_nhwc_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(
        src, dst, IH, IW, OH, OW, C / 4, (float)IH / (float)OH, (float)IW / (float)OW, accumulate, total,
        src_batch_stride / 4, reinterpret_cast<uint2*>(dst_hi), reinterpret_cast<uint2*>(dst_lo));
    return pvsg_launch_status();
}

extern "C" int pvsg_bilinear_resize_scaled(const float* src, float* dst, int B, int IH, int IW, int OH, int OW, int C,
                                           float scale_h, float scale_w, void* stream) {
    PVSG_CHECK_ARG(src && dst && B > 0 && IH > 0 && IW > 0 && OH > 0 && OW > 0 && C > 0 && C % 4 == 0 && scale_h > 0.f && scale_w > 0.f);
    const int64_t total = (int64_t)B * OH * OW * (C / 4);
    bilinear_nhwc_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(src, dst, IH, IW, OH, OW, C / 4, scale_h, scale_w, 0, total,
                                                                         (int64_t)IH * IW * (C / 4));
    return pvsg_launch_status();
}

extern "C" int pvsg_maxpool3x3s2_nhwc(const float* x, float* y, int B, int H, int W, int C, void* stream) {
    PVSG_CHECK_ARG(x && y && B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0);
    const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
    const int64_t total = (int64_t)B * OH * OW * (C / 4);
    maxpool_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(x, y, H, W, OH, OW, C / 4, total);
    return pvsg_launch_status();
}

extern "C" int pvsg_nchw_to_nhwc(const float* x, float* y, int B, int C, int H, int W, void* stream) {
    PVSG_CHECK_ARG(x && y && B > 0 && C > 0 && H > 0 && W > 0);
    const int64_t HW = (int64_t)H * W;
    dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)B);
    PVSG_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535);
    transpose_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, y, C, HW);
    return pvsg_launch_status();
}

extern "C" int pvsg_nhwc_to_nchw(const float* x, float* y, int B, int C, int H, int W, void* stream) {
    PVSG_CHECK_ARG(x && y && B > 0 && C > 0 && H > 0 && W > 0);
    const int64_t HW = (int64_t)H * W;
    dim3 grid((unsigned)((C + 31) / 32), (unsigned)((HW + 31) / 32), (unsigned)B);
    PVSG_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535);
    transpose_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, y, HW, C);
    return pvsg_launch_status();
}

extern "C" int pvsg_sine_pe(float* out, const float* dim_t, const float* dim_t_z, const float* add_vec,
                            int T, int H, int W, int num_feats, float scale, float eps, void* stream) {
    PVSG_CHECK_ARG(out && dim_t && H > 0 && W > 0 && num_feats > 0 && T >= 0 && (T == 0 || dim_t_z));
    const int64_t total = (int64_t)(T > 0 ? T : 1) * H * W * 2 * num_feats;
    sine_pe_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(out, dim_t, dim_t_z, add_vec, T, H, W,
                                                                   num_feats, scale, eps, total);
    return pvsg_launch_status();
}
