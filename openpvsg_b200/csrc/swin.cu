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

// ---- tensor-core variant ------------------------------------------------------------------
// Same (window, head) decomposition on mma.sync.m16n8k16 with split-bf16 operands (every product =
// lo.hi + hi.lo + hi.hi, fp32 accumulate: fp32-grade like the rest of the engine; scheme and
// fragment addressing as csrc/attention_mma.cu).  A warp owns 16 query rows; all keys of the
// window (<= 144) are staged once; the score row is processed in chunks of 48 keys with an online
// softmax in base 2 (ex2.approx).  K and V are converted to (hi, lo) planes while
// they are staged into shared memory ([key][32 + 8] bf16, conflict-free ldmatrix); the S fragment
// of Q K^T is the A fragment of P V.  ~20x fewer issue slots per row than the SIMT kernels above.
constexpr int kPitch = kHeadDim + 8;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// packed conversions (cvt.rn.bf16x2.f32, ALU pipe) as in gemm_tc.cu: hi = rn(x), lo = rn(x - hi)
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);          // .x (low half) = a
    hi = *reinterpret_cast<const uint32_t*>(&h2);
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - h0, b - h1);
    lo = *reinterpret_cast<const uint32_t*>(&l2);
}

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

// KS16 = key steps of 16 (9 for window 12, 4 for windows <= 8), processed in chunks of CH16 steps so that the score
// fragment stays at 8 CH16 registers and two CTAs fit an SM (the one-tile version needed 164 registers = one CTA of
// nine warps per SM and ran latency-bound); blockDim = 32 * ceil(N / 16)
template <int KS16, int CH16>
__global__ void __launch_bounds__(32 * KS16, CH16 == 1 ? 3 : 2) window_attention_mma_kernel(const float* __restrict__ qkv, const float* __restrict__ qkv_bias,
                                                                         const float* __restrict__ bias_table, float* __restrict__ out,
                                                                         __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo,
                                                                         int H, int W, int C, int heads, int ws, int shift,
                                                                         int nwy, int nwx, float scale) {
    constexpr int NK = 16 * KS16;          // padded key count
    extern __shared__ __align__(16) uint8_t wsm[];
    __nv_bfloat16 (*plane)[NK][kPitch] = reinterpret_cast<__nv_bfloat16 (*)[NK][kPitch]>(wsm);   // Khi Klo Vhi Vlo
    float* tbl = reinterpret_cast<float*>(wsm + sizeof(__nv_bfloat16) * 4 * NK * kPitch);       // [(2 ws - 1)^2]
    int64_t* src = reinterpret_cast<int64_t*>(tbl + (2 * kMaxWs - 1) * (2 * kMaxWs - 1) + 1);     // [NK] (8-byte aligned: 530 floats)
    int* kinfo = reinterpret_cast<int*>(src + NK);                                               // [NK]: key offset | region << 16

    const int N = ws * ws;
    const int head = blockIdx.y;
    int win = blockIdx.x;
    const int wx = win % nwx; win /= nwx;
    const int wy = win % nwy;
    const int b = win / nwy;
    const int Hp = nwy * ws, Wp = nwx * ws;
    const int tid = threadIdx.x;
    const int span = 2 * ws - 1;

    constexpr float kLog2e = 1.4426950408889634f;   // scores are kept in base 2: folded into q and the bias table
    for (int i = tid; i < span * span; i += blockDim.x) tbl[i] = kLog2e * __ldg(bias_table + (int64_t)i * heads + head);
    for (int tk = tid; tk < NK; tk += blockDim.x) {
        int64_t so = -1;
        int info = 0;
        if (tk < N) {
            const int iy = tk / ws, ix = tk - iy * ws;
            const int y = wy * ws + iy, x = wx * ws + ix;
            int ys = y + shift, xs = x + shift;
            if (ys >= Hp) ys -= Hp;
            if (xs >= Wp) xs -= Wp;
            so = (ys < H && xs < W) ? (((int64_t)b * H + ys) * W + xs) : -1;
            int r = 0;
            if (shift > 0) {
                const int ry = y < Hp - ws ? 0 : (y < Hp - shift ? 1 : 2);
                const int rx = x < Wp - ws ? 0 : (x < Wp - shift ? 1 : 2);
                r = ry * 3 + rx;
            }
            info = (iy * span + ix) | (r << 16);
        }
        src[tk] = so;
        kinfo[tk] = info;
    }
    __syncthreads();
    // K, V -> (hi, lo) planes in shared memory; rows beyond N are zero (0 * garbage must not become NaN)
    const int C3 = 3 * C;
    for (int e = tid; e < NK * 8; e += blockDim.x) {
        const int tk = e >> 3, part = e & 7;
        const int ch = head * kHeadDim + part * 4;
        float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
        if (tk < N) {
            const int64_t so = src[tk];
            const float* base = so >= 0 ? qkv + so * C3 : qkv_bias;
            kk = __ldg(reinterpret_cast<const float4*>(base + C + ch));
            vv = __ldg(reinterpret_cast<const float4*>(base + 2 * C + ch));
        }
        uint32_t h0, l0, h1, l1;
        split_pair(kk.x, kk.y, h0, l0); split_pair(kk.z, kk.w, h1, l1);
        *reinterpret_cast<uint2*>(&plane[0][tk][part * 4]) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(&plane[1][tk][part * 4]) = make_uint2(l0, l1);
        split_pair(vv.x, vv.y, h0, l0); split_pair(vv.z, vv.w, h1, l1);
        *reinterpret_cast<uint2*>(&plane[2][tk][part * 4]) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(&plane[3][tk][part * 4]) = make_uint2(l0, l1);
    }
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    // ---- Q fragments (rows 16 warp + g, + 8), scaled, split ----
    int row[2] = {16 * warp + g, 16 * warp + g + 8};
    int64_t my_src[2];
    int bias_row[2], my_region[2];
    uint32_t qh[2][4], ql[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const bool live = row[r] < N;
        const int tt = live ? row[r] : 0;
        my_src[r] = live ? src[tt] : -1;
        const int info = kinfo[tt];
        my_region[r] = info >> 16;
        bias_row[r] = (info & 0xffff) + (ws - 1) * span + (ws - 1);   // index(i, j) = bias_row - key offset(j)
    }
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int half = 0; half < 2; ++half)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const float* qp = my_src[r] >= 0 ? qkv + my_src[r] * C3 : qkv_bias;
                const float2 v = __ldg(reinterpret_cast<const float2*>(qp + head * kHeadDim + 16 * ks + 8 * half + 2 * t));
                split_pair(v.x * (scale * kLog2e), v.y * (scale * kLog2e), qh[ks][2 * half + r], ql[ks][2 * half + r]);
            }
    __syncthreads();
    if (16 * warp >= N) return;

    // ---- key chunks of 16 CH16 keys: S = Q K^T, bias + mask, online softmax (base 2), O += P V ----
    constexpr int CT = 2 * CH16;           // S n-tiles per chunk
    float o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[i][e] = 0.f;
    float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
#pragma unroll 1
    for (int c0 = 0; c0 < KS16; c0 += CH16) {
        const int key0 = 16 * c0;
        float s[CT][4];
#pragma unroll
        for (int j = 0; j < CT; ++j) {
#pragma unroll
            for (int e = 0; e < 4; ++e) s[j][e] = 0.f;
            uint32_t kh[4], kl[4];
            const int r = key0 + 8 * j + (lane & 7), c = (lane >> 3) * 8;
            ldsm4(kh, smem_addr(&plane[0][r][c]));
            ldsm4(kl, smem_addr(&plane[1][r][c]));
#pragma unroll
            for (int k2 = 0; k2 < 2; ++k2) {
                mma16816(s[j], ql[k2], kh[2 * k2], kh[2 * k2 + 1]);   // small terms first
                mma16816(s[j], qh[k2], kl[2 * k2], kl[2 * k2 + 1]);
                mma16816(s[j], qh[k2], kh[2 * k2], kh[2 * k2 + 1]);
            }
        }
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < CT; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int r = e >> 1;
                const int key = key0 + 8 * j + 2 * t + (e & 1);
                float v = -INFINITY;
                if (key < N) {
                    const int info = kinfo[key];
                    v = s[j][e] + tbl[bias_row[r] - (info & 0xffff)];
                    if ((info >> 16) != my_region[r]) v += -100.f * kLog2e;
                }
                s[j][e] = v;
                mx[r] = fmaxf(mx[r], v);
            }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float mn = fmaxf(m[r], mx[r]);          // finite from the first chunk on (key 0 always exists)
            const float corr = exp2f(m[r] - mn);           // exp2(-inf) = 0 on the first chunk
            m[r] = mn;
            l[r] *= corr;
#pragma unroll
            for (int i = 0; i < 4; ++i) { o[i][2 * r] *= corr; o[i][2 * r + 1] *= corr; }
        }
#pragma unroll
        for (int j = 0; j < CT; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float p = exp2f(s[j][e] - m[e >> 1]);
                s[j][e] = p;
                l[e >> 1] += p;
            }
#pragma unroll
        for (int kk = 0; kk < CH16; ++kk) {
            uint32_t ph[4], pl[4];
            split_pair(s[2 * kk][0], s[2 * kk][1], ph[0], pl[0]);
            split_pair(s[2 * kk][2], s[2 * kk][3], ph[1], pl[1]);
            split_pair(s[2 * kk + 1][0], s[2 * kk + 1][1], ph[2], pl[2]);
            split_pair(s[2 * kk + 1][2], s[2 * kk + 1][3], ph[3], pl[3]);
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                uint32_t vh[4], vl[4];
                const int mi = lane >> 3;
                const int r = key0 + 16 * kk + (mi & 1) * 8 + (lane & 7), c = 16 * np + (mi >> 1) * 8;
                ldsm4_t(vh, smem_addr(&plane[2][r][c]));
                ldsm4_t(vl, smem_addr(&plane[3][r][c]));
                mma16816(o[2 * np], pl, vh[0], vh[1]);
                mma16816(o[2 * np], ph, vl[0], vl[1]);
                mma16816(o[2 * np], ph, vh[0], vh[1]);
                mma16816(o[2 * np + 1], pl, vh[2], vh[3]);
                mma16816(o[2 * np + 1], ph, vl[2], vl[3]);
                mma16816(o[2 * np + 1], ph, vh[2], vh[3]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
        l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
        if (my_src[r] < 0) continue;      // padded position or no such row: cropped away
        const float inv = 1.f / l[r];
        const int64_t off = my_src[r] * C + head * kHeadDim + 2 * t;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float v0 = o[i][2 * r] * inv, v1 = o[i][2 * r + 1] * inv;
            if (out) *reinterpret_cast<float2*>(out + off + 8 * i) = make_float2(v0, v1);
            if (out_hi) {
                uint32_t hh, ll;
                split_pair(v0, v1, hh, ll);
                *reinterpret_cast<uint32_t*>(out_hi + off + 8 * i) = hh;
                *reinterpret_cast<uint32_t*>(out_lo + off + 8 * i) = ll;
            }
        }
    }
}

template <int KS16>
constexpr size_t window_smem() {
    return sizeof(__nv_bfloat16) * 4 * (16 * KS16) * kPitch + sizeof(float) * ((2 * kMaxWs - 1) * (2 * kMaxWs - 1) + 1) +
           sizeof(int64_t) * (16 * KS16) + sizeof(int) * (16 * KS16);
}

template <int KS16, int CH16>
static int configure_window_mma() {
    constexpr size_t smem = window_smem<KS16>();
    static bool configured[PVSG_MAX_DEVICES];
    if (pvsg_first_use_on_device(configured)) {
        // 50 KB per CTA: ask for the large shared-memory carve-out so that three CTAs fit an SM (ncu showed the default
        // carve-out capping the kernel at two, profiles/r01s_ncu_window_attention_mma.json); L1 is not reused here
        if (cudaFuncSetAttribute(window_attention_mma_kernel<KS16, CH16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
                cudaSuccess ||
            cudaFuncSetAttribute(window_attention_mma_kernel<KS16, CH16>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared) != cudaSuccess)
            return PVSG_ERR_LAUNCH;
    }
    return PVSG_OK;
}

template <int KS16, int CH16>
static int launch_window_mma(dim3 grid, int N, cudaStream_t st, const float* qkv, const float* qkv_bias, const float* bias_table,
                             float* out, __nv_bfloat16* oh, __nv_bfloat16* ol, int H, int W, int C, int heads, int window,
                             int shift, int nwy, int nwx, float scale) {
    constexpr size_t smem = window_smem<KS16>();
    if (const int rc = configure_window_mma<KS16, CH16>()) return rc;
    const int warps = (N + 15) / 16;
    window_attention_mma_kernel<KS16, CH16><<<grid, 32 * warps, smem, st>>>(qkv, qkv_bias, bias_table, out, oh, ol, H, W, C, heads,
                                                                      window, shift, nwy, nwx, scale);
    return pvsg_launch_status();
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

int configure_patch_merge() {
    static bool configured[PVSG_MAX_DEVICES];
    if (pvsg_first_use_on_device(configured) &&
        cudaFuncSetAttribute(patch_merge_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 4 * 1024 * 4) != cudaSuccess)
        return PVSG_ERR_LAUNCH;
    return PVSG_OK;
}

}  // namespace

int pvsg_internal::configure_swin() {
    int rc = configure_window_mma<4, 2>();
    if (!rc) rc = configure_window_mma<9, 1>();
    if (!rc) rc = configure_window_mma<9, 3>();
    return rc ? rc : configure_patch_merge();
}

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
    // PVSG_WINATT_IMPL: 0 / unset = tensor-core kernel; 1, 2 = the SIMT fp32 kernels with one / two query rows
    // per thread (kept for A/B timing and as the exact-fp32 cross-check of the split-bf16 path)
    static const int impl = [] { const char* e = getenv("PVSG_WINATT_IMPL"); return e ? atoi(e) : 0; }();
    __nv_bfloat16* oh = reinterpret_cast<__nv_bfloat16*>(out_hi);
    __nv_bfloat16* ol = reinterpret_cast<__nv_bfloat16*>(out_lo);
    cudaStream_t st = as_stream(stream);
    const int N = window * window;
    if (impl == 0) {
        // key chunk per online-softmax step: 16 keys / 72 registers / three CTAs per SM (default, measured 399 vs 453 us at
        // stage 0), or PVSG_WINATT_CHUNK=3: 48 keys / 96 registers / two CTAs per SM
        static const int chunk = [] { const char* e = getenv("PVSG_WINATT_CHUNK"); return e ? atoi(e) : 1; }();
        if (N <= 64)
            return launch_window_mma<4, 2>(grid, N, st, qkv, qkv_bias, bias_table, out, oh, ol, H, W, C, heads, window, shift,
                                           nwy, nwx, scale);
        return chunk == 1 ? launch_window_mma<9, 1>(grid, N, st, qkv, qkv_bias, bias_table, out, oh, ol, H, W, C, heads, window,
                                                    shift, nwy, nwx, scale)
                          : launch_window_mma<9, 3>(grid, N, st, qkv, qkv_bias, bias_table, out, oh, ol, H, W, C, heads, window,
                                                    shift, nwy, nwx, scale);
    }
    if (impl == 2)
        window_attention_kernel<2><<<grid, 96, 0, st>>>(qkv, qkv_bias, bias_table, out, oh, ol, H, W, C, heads, window, shift,
                                                        nwy, nwx, scale);
    else
        window_attention_kernel<1><<<grid, 160, 0, st>>>(qkv, qkv_bias, bias_table, out, oh, ol, H, W, C, heads, window, shift,
                                                         nwy, nwx, scale);
    return pvsg_launch_status();
}

extern "C" int pvsg_patch_merge_ln(const float* x, const float* gamma, const float* beta, float* y, int B, int H, int W,
                                   int C, float eps, void* stream) {
    PVSG_CHECK_ARG(x && gamma && beta && y && B > 0 && H > 0 && W > 0 && C > 0);
    if (C % 4 != 0 || C > 1024) return PVSG_ERR_UNSUPPORTED;
    const int OH = (H + 1) / 2, OW = (W + 1) / 2;
    const int64_t total = (int64_t)B * OH * OW;
    const size_t smem = (size_t)4 * 4 * C * sizeof(float);
    if (const int rc = configure_patch_merge()) return rc;
    const unsigned grid = (unsigned)imin64((total + 3) / 4, 148 * 16);
    patch_merge_ln_kernel<<<grid, 128, smem, as_stream(stream)>>>(x, gamma, beta, y, B, H, W, C, OH, OW, eps);
    return pvsg_launch_status();
}
