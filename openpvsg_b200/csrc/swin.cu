// Swin backbone kernels (sm_100a): shifted-window attention and patch merging + LayerNorm.
//
// The reference names a Swin-B configuration (BASELINE configs[2]) but ships no Swin code; the
// semantics restated here are mmdet 2.25.0 `mmdet/models/backbones/swin.py` (WindowMSA :23-126,
// ShiftWindowMSA :129-285) and `mmdet/models/utils/transformer.py` (PatchMerging :235-352), the
// versions the reference pins (README.md:123-125).  Dense projections (qkv, proj, FFN with exact
// GELU, reduction) run on the tcgen05 GEMM engine; these two kernels are the glue around them.
#include "common.cuh"
#include <cuda_bf16.h>
#include <stdlib.h>

namespace {

constexpr int kHeadDim = 32;     // every Swin variant: C / heads = 32
constexpr int kMaxTokens = 144;  // window 12 x 12
constexpr int kMaxWs = 12;

// One CTA = one (window, head).  Pad + cyclic shift + window partition + reverse + crop are index
// arithmetic: token (iy, ix) of window (wy, wx) sits at (y, x) = (wy ws + iy, wx ws + ix) of the
// rolled, padded map, i.e. at ((y + shift) mod Hp, (x + shift) mod Wp) of the padded map; padded
// positions (>= H or >= W) carry a zero input, so their q/k/v are the qkv bias.  A thread owns TWO
// query rows (t and t + ceil(N/2)): q and the output accumulators stay in registers, K and V of the
// window live in shared memory and every broadcast read of a key / value row feeds both rows (the
// kernel is bound by the shared-memory pipe: 16 LDS.128 per key).  Softmax is online over chunks
// of four keys.  fp32 throughout; the result leaves as fp32 or directly as the split-bf16 operand
// planes of the proj GEMM.

__device__ __forceinline__ void store_row(float* __restrict__ out, __nv_bfloat16* __restrict__ out_hi,
                                          __nv_bfloat16* __restrict__ out_lo, int64_t off, const float* acc, float inv) {
    if (out) {
#pragma unroll
        for (int d = 0; d < kHeadDim; d += 4)
            *reinterpret_cast<float4*>(out + off + d) =
                make_float4(acc[d] * inv, acc[d + 1] * inv, acc[d + 2] * inv, acc[d + 3] * inv);
    }
    if (out_hi) {
#pragma unroll
        for (int d = 0; d < kHeadDim; d += 8) {
            uint32_t h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float v0 = acc[d + 2 * e] * inv, v1 = acc[d + 2 * e + 1] * inv;
                const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
                const __nv_bfloat16 l0 = __float2bfloat16_rn(v0 - __bfloat162float(h0));
                const __nv_bfloat16 l1 = __float2bfloat16_rn(v1 - __bfloat162float(h1));
                h[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                l[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
            }
            *reinterpret_cast<uint4*>(out_hi + off + d) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(out_lo + off + d) = make_uint4(l[0], l[1], l[2], l[3]);
        }
    }
}

template <int kRows>
__global__ void __launch_bounds__(kRows == 2 ? 96 : 160) window_attention_kernel(const float* __restrict__ qkv, const float* __restrict__ qkv_bias,
                                                              const float* __restrict__ bias_table, float* __restrict__ out,
                                                              __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
                                                              int H, int W, int C, int heads, int ws, int shift,
                                                              int nwy, int nwx, float scale) {
    __shared__ __align__(16) float Ks[kMaxTokens * kHeadDim];
    __shared__ __align__(16) float Vs[kMaxTokens * kHeadDim];
    __shared__ float tbl[(2 * kMaxWs - 1) * (2 * kMaxWs - 1)];
    __shared__ int64_t src[kMaxTokens];   // token row of the unpadded map, -1 = padded position
    __shared__ unsigned char region[kMaxTokens];

    const int N = ws * ws;
    const int half = kRows == 2 ? (N + 1) / 2 : N;
    const int head = blockIdx.y;
    int win = blockIdx.x;
    const int wx = win % nwx; win /= nwx;
    const int wy = win % nwy;
    const int b = win / nwy;
    const int Hp = nwy * ws, Wp = nwx * ws;
    const int tid = threadIdx.x;
    const int span = 2 * ws - 1;

    for (int i = tid; i < span * span; i += blockDim.x) tbl[i] = __ldg(bias_table + (int64_t)i * heads + head);
    for (int t = tid; t < N; t += blockDim.x) {
        const int iy = t / ws, ix = t - iy * ws;
        const int y = wy * ws + iy, x = wx * ws + ix;
        int ys = y + shift, xs = x + shift;
        if (ys >= Hp) ys -= Hp;
        if (xs >= Wp) xs -= Wp;
        src[t] = (ys < H && xs < W) ? (((int64_t)b * H + ys) * W + xs) : -1;
        int r = 0;
        if (shift > 0) {   // img_mask regions of ShiftWindowMSA: slices (0,-ws), (-ws,-shift), (-shift,None)
            const int ry = y < Hp - ws ? 0 : (y < Hp - shift ? 1 : 2);
            const int rx = x < Wp - ws ? 0 : (x < Wp - shift ? 1 : 2);
            r = ry * 3 + rx;
        }
        region[t] = (unsigned char)r;
    }
    __syncthreads();
    // K, V -> shared memory: 8 lanes x float4 per token
    const int C3 = 3 * C;
    for (int e = tid; e < N * 8; e += blockDim.x) {
        const int t = e >> 3, part = e & 7;
        const int64_t s = src[t];
        const int ch = head * kHeadDim + part * 4;
        float4 kk, vv;
        if (s >= 0) {
            kk = __ldg(reinterpret_cast<const float4*>(qkv + s * C3 + C + ch));
            vv = __ldg(reinterpret_cast<const float4*>(qkv + s * C3 + 2 * C + ch));
        } else {
            kk = __ldg(reinterpret_cast<const float4*>(qkv_bias + C + ch));
            vv = __ldg(reinterpret_cast<const float4*>(qkv_bias + 2 * C + ch));
        }
        reinterpret_cast<float4*>(Ks)[e] = kk;
        reinterpret_cast<float4*>(Vs)[e] = vv;
    }
    // rows of this thread; a row that does not exist or is a padded position computes on row 0 and is not stored
    float q[kRows][kHeadDim], acc[kRows][kHeadDim], m[kRows], l[kRows];
    int bias_row[kRows], my_region[kRows];
    int64_t my_src[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const int t = tid + r * half;
        const bool live = tid < half && t < N;
        const int tt = live ? t : 0;
        my_src[r] = live ? src[tt] : -1;
        my_region[r] = region[tt];
        const int iy = tt / ws, ix = tt - iy * ws;
        bias_row[r] = (iy + ws - 1) * span + (ix + ws - 1);   // index(i, j) = bias_row - (jy span + jx)
        const float* qp = my_src[r] >= 0 ? qkv + my_src[r] * C3 + head * kHeadDim : qkv_bias + head * kHeadDim;
#pragma unroll
        for (int d = 0; d < kHeadDim; d += 4) {
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(qp + d));
            q[r][d] = t4.x * scale; q[r][d + 1] = t4.y * scale; q[r][d + 2] = t4.z * scale; q[r][d + 3] = t4.w * scale;
        }
#pragma unroll
        for (int d = 0; d < kHeadDim; ++d) acc[r][d] = 0.f;
        m[r] = -INFINITY;
        l[r] = 0.f;
    }
    __syncthreads();
    if (my_src[0] < 0 && my_src[kRows - 1] < 0) return;   // nothing to store (padded rows are cropped away)

    for (int j0 = 0; j0 < N; j0 += 4) {
        float s[kRows][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u;
            if (j < N) {
                const float4* kr = reinterpret_cast<const float4*>(Ks + j * kHeadDim);
                float dot[kRows];
#pragma unroll
                for (int r = 0; r < kRows; ++r) dot[r] = 0.f;
#pragma unroll
                for (int d = 0; d < kHeadDim / 4; ++d) {
                    const float4 k4 = kr[d];
#pragma unroll
                    for (int r = 0; r < kRows; ++r) {
                        dot[r] = fmaf(q[r][4 * d], k4.x, dot[r]); dot[r] = fmaf(q[r][4 * d + 1], k4.y, dot[r]);
                        dot[r] = fmaf(q[r][4 * d + 2], k4.z, dot[r]); dot[r] = fmaf(q[r][4 * d + 3], k4.w, dot[r]);
                    }
                }
                const int jy = j / ws, jx = j - jy * ws;
                const int rj = region[j];
#pragma unroll
                for (int r = 0; r < kRows; ++r) {
                    float v = dot[r] + tbl[bias_row[r] - (jy * span + jx)];
                    if (rj != my_region[r]) v += -100.f;
                    s[r][u] = v;
                }
            } else {
#pragma unroll
                for (int r = 0; r < kRows; ++r) s[r][u] = -INFINITY;
            }
        }
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            const float cm = fmaxf(fmaxf(s[r][0], s[r][1]), fmaxf(s[r][2], s[r][3]));
            if (cm > m[r]) {
                const float corr = expf(m[r] - cm);   // m = -inf on the first chunk: corr = 0, acc = l = 0
                l[r] *= corr;
#pragma unroll
                for (int d = 0; d < kHeadDim; ++d) acc[r][d] *= corr;
                m[r] = cm;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u;
            if (j < N) {
                float p[kRows];
#pragma unroll
                for (int r = 0; r < kRows; ++r) {
                    p[r] = expf(s[r][u] - m[r]);
                    l[r] += p[r];
                }
                const float4* vr = reinterpret_cast<const float4*>(Vs + j * kHeadDim);
#pragma unroll
                for (int d = 0; d < kHeadDim / 4; ++d) {
                    const float4 v4 = vr[d];
#pragma unroll
                    for (int r = 0; r < kRows; ++r) {
                        acc[r][4 * d] = fmaf(p[r], v4.x, acc[r][4 * d]); acc[r][4 * d + 1] = fmaf(p[r], v4.y, acc[r][4 * d + 1]);
                        acc[r][4 * d + 2] = fmaf(p[r], v4.z, acc[r][4 * d + 2]); acc[r][4 * d + 3] = fmaf(p[r], v4.w, acc[r][4 * d + 3]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r)
        if (my_src[r] >= 0) store_row(out, out_hi, out_lo, my_src[r] * C + head * kHeadDim, acc[r], 1.f / l[r]);
}

// One warp = one merged token: gathers the 2x2 neighbourhood in nn.Unfold channel order
// (channel c of tap (kh, kw) -> c*4 + kh*2 + kw; zeros beyond an odd edge), LayerNorm over 4C.
__global__ void __launch_bounds__(128) patch_merge_ln_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float* __restrict__ y,
                                                             int B, int H, int W, int C, int OH, int OW, float eps) {
    extern __shared__ float rowbuf[];   // [4 warps][4C]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C4 = 4 * C;
    float* row = rowbuf + (size_t)warp * C4;
    const int64_t total = (int64_t)B * OH * OW;
    for (int64_t tok = (int64_t)blockIdx.x * 4 + warp; tok < total; tok += (int64_t)gridDim.x * 4) {
        const int ox = (int)(tok % OW);
        const int oy = (int)((tok / OW) % OH);
        const int b = (int)(tok / ((int64_t)OW * OH));
        float sum = 0.f;
#pragma unroll
        for (int tap = 0; tap < 4; ++tap) {
            const int yy = 2 * oy + (tap >> 1), xx = 2 * ox + (tap & 1);
            const bool in = yy < H && xx < W;
            const float* xp = x + (((int64_t)b * H + yy) * W + xx) * C;
            for (int c = lane; c < C; c += 32) {
                const float v = in ? __ldg(xp + c) : 0.f;
                row[c * 4 + tap] = v;
                sum += v;
            }
        }
        const float mean = warp_sum(sum) / (float)C4;
        __syncwarp();
        float var = 0.f;
        for (int i = lane; i < C4; i += 32) {
            const float d = row[i] - mean;
            var = fmaf(d, d, var);
        }
        const float rstd = rsqrtf(warp_sum(var) / (float)C4 + eps);
        float* yp = y + tok * C4;
        for (int i = lane * 4; i < C4; i += 128) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + i));
            const float4 bt = __ldg(reinterpret_cast<const float4*>(beta + i));
            const float4 v = *reinterpret_cast<const float4*>(row + i);
            *reinterpret_cast<float4*>(yp + i) = make_float4((v.x - mean) * rstd * g.x + bt.x, (v.y - mean) * rstd * g.y + bt.y,
                                                             (v.z - mean) * rstd * g.z + bt.z, (v.w - mean) * rstd * g.w + bt.w);
        }
        __syncwarp();
    }
}

}  // namespace

extern "C" int pvsg_window_attention(const float* qkv, const float* qkv_bias, const float* bias_table, float* out,
                                     void* out_hi, void* out_lo, int B, int H, int W, int C, int heads, int window,
                                     int shift, void* stream) {
    PVSG_CHECK_ARG(qkv && qkv_bias && bias_table && (out || out_hi) && (out_hi == nullptr) == (out_lo == nullptr));
    PVSG_CHECK_ARG(B > 0 && H > 0 && W > 0 && heads > 0 && window > 0 && shift >= 0 && shift < window);
    if (window > kMaxWs || C != heads * kHeadDim) return PVSG_ERR_UNSUPPORTED;
    const int nwy = (H + window - 1) / window, nwx = (W + window - 1) / window;
    const int64_t wins = (int64_t)B * nwy * nwx;
    PVSG_CHECK_ARG(wins <= 0x7fffffffLL && heads <= 65535);
    const float scale = 1.f / sqrtf((float)kHeadDim);
    dim3 grid((unsigned)wins, (unsigned)heads);
    // two query rows per thread halve the shared-memory traffic per FMA (measured faster, profiles/README.md);
    // PVSG_WINATT_ROWS=1 keeps the one-row variant selectable for A/B timing
    static const int rows = [] { const char* e = getenv("PVSG_WINATT_ROWS"); return (e && e[0] == '1') ? 1 : 2; }();
    __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(out_hi);
    __nv_bfloat16* ol = reinterpret_cast<__nv_bfloat16*>(out_lo);
    if (rows == 2)
        window_attention_kernel<2><<<grid, 96, 0, as_stream(stream)>>>(qkv, qkv_bias, bias_table, out, oh, ol, H, W, C, heads,
                                                                      window, shift, nwy, nwx, scale);
    else
        window_attention_kernel<1><<<grid, 160, 0, as_stream(stream)>>>(qkv, qkv_bias, bias_table, out, oh, ol, H, W, C, heads,
                                                                       window, shift, nwy, nwx, scale);
    return pvsg_launch_status();
}

extern "C" int pvsg_patch_merge_ln(const float* x, const float* gamma, const float* beta, float* y, int B, int H, int W,
                                   int C, float eps, void* stream) {
    PVSG_CHECK_ARG(x && gamma && beta && y && B > 0 && H > 0 && W > 0 && C > 0);
    if (C % 4 != 0 || C > 1024) return PVSG_ERR_UNSUPPORTED;
    const int OH = (H + 1) / 2, OW = (W + 1) / 2;
    const int64_t total = (int64_t)B * OH * OW;
    const size_t smem = (size_t)4 * 4 * C * sizeof(float);
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(patch_merge_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 4 * 1024 * 4) !=
            cudaSuccess)
            return PVSG_ERR_LAUNCH;
        configured = true;
    }
    const unsigned grid = (unsigned)imin64((total + 3) / 4, 148 * 16);
    patch_merge_ln_kernel<<<grid, 128, smem, as_stream(stream)>>>(x, gamma, beta, y, B, H, W, C, OH, OW, eps);
    return pvsg_launch_status();
}
