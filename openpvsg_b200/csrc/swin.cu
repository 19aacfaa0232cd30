// Swin backbone kernels (sm_100a): shifted-window attention and patch merging + LayerNorm.
//
// The reference names a Swin-B configuration (BASELINE configs[2]) but ships no Swin code; the
// semantics restated here are mmdet 2.25.0 `mmdet/models/backbones/swin.py` (WindowMSA :23-126,
// ShiftWindowMSA :129-285) and `mmdet/models/utils/transformer.py` (PatchMerging :235-352), the
// versions the reference pins (README.md:123-125).  Dense projections (qkv, proj, FFN with exact
// GELU, reduction) run on the tcgen05 GEMM engine; these two kernels are the glue around them.
#include "common.cuh"

namespace {

constexpr int kHeadDim = 32;     // every Swin variant: C / heads = 32
constexpr int kMaxTokens = 144;  // window 12 x 12
constexpr int kMaxWs = 12;

// One CTA = one (window, head).  Pad + cyclic shift + window partition + reverse + crop are index
// arithmetic: token (iy, ix) of window (wy, wx) sits at (y, x) = (wy ws + iy, wx ws + ix) of the
// rolled, padded map, i.e. at ((y + shift) mod Hp, (x + shift) mod Wp) of the padded map; padded
// positions (>= H or >= W) carry a zero input, so their q/k/v are the qkv bias.  Thread i owns
// query row i (q and the output accumulator in registers), K and V of the window live in shared
// memory and are read as broadcasts; softmax is online over chunks of four keys.  fp32 throughout.
__global__ void __launch_bounds__(160) window_attention_kernel(const float* __restrict__ qkv, const float* __restrict__ qkv_bias,
                                                               const float* __restrict__ bias_table, float* __restrict__ out,
                                                               int H, int W, int C, int heads, int ws, int shift,
                                                               int nwy, int nwx, float scale) {
    __shared__ __align__(16) float Ks[kMaxTokens * kHeadDim];
    __shared__ __align__(16) float Vs[kMaxTokens * kHeadDim];
    __shared__ float tbl[(2 * kMaxWs - 1) * (2 * kMaxWs - 1)];
    __shared__ int64_t src[kMaxTokens];   // token offset (elements) into qkv rows, -1 = padded position
    __shared__ unsigned char region[kMaxTokens];

    const int N = ws * ws;
    const int head = blockIdx.y;
    int win = blockIdx.x;
    const int wx = win % nwx; win /= nwx;
    const int wy = win % nwy;
    const int b = win / nwy;
    const int Hp = nwy * ws, Wp = nwx * ws;
    const int tid = threadIdx.x;
    const int span = 2 * ws - 1;

    for (int i = tid; i < span * span; i += blockDim.x) tbl[i] = __ldg(bias_table + (int64_t)i * heads + head);
    if (tid < N) {
        const int iy = tid / ws, ix = tid - iy * ws;
        const int y = wy * ws + iy, x = wx * ws + ix;
        int ys = y + shift, xs = x + shift;
        if (ys >= Hp) ys -= Hp;
        if (xs >= Wp) xs -= Wp;
        src[tid] = (ys < H && xs < W) ? (((int64_t)b * H + ys) * W + xs) : -1;
        int r = 0;
        if (shift > 0) {   // img_mask regions of ShiftWindowMSA: slices (0,-ws), (-ws,-shift), (-shift,None)
            const int ry = y < Hp - ws ? 0 : (y < Hp - shift ? 1 : 2);
            const int rx = x < Wp - ws ? 0 : (x < Wp - shift ? 1 : 2);
            r = ry * 3 + rx;
        }
        region[tid] = (unsigned char)r;
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
    float q[kHeadDim];
    int iy = 0, ix = 0, my_region = 0;
    int64_t my_src = -1;
    if (tid < N) {
        iy = tid / ws; ix = tid - iy * ws;
        my_src = src[tid];
        my_region = region[tid];
        const float* qp = my_src >= 0 ? qkv + my_src * C3 + head * kHeadDim : qkv_bias + head * kHeadDim;
#pragma unroll
        for (int d = 0; d < kHeadDim; d += 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(qp + d));
            q[d] = t.x * scale; q[d + 1] = t.y * scale; q[d + 2] = t.z * scale; q[d + 3] = t.w * scale;
        }
    }
    __syncthreads();
    if (tid >= N || my_src < 0) return;   // padded rows are cropped away

    float acc[kHeadDim];
#pragma unroll
    for (int d = 0; d < kHeadDim; ++d) acc[d] = 0.f;
    float m = -INFINITY, l = 0.f;
    const int bias_row = (iy + ws - 1) * span + (ix + ws - 1);   // index(i, j) = bias_row - (jy span + jx)
    for (int j0 = 0; j0 < N; j0 += 4) {
        float s[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u;
            if (j < N) {
                const float4* kr = reinterpret_cast<const float4*>(Ks + j * kHeadDim);
                float dot = 0.f;
#pragma unroll
                for (int d = 0; d < kHeadDim / 4; ++d) {
                    const float4 k4 = kr[d];
                    dot = fmaf(q[4 * d], k4.x, dot); dot = fmaf(q[4 * d + 1], k4.y, dot);
                    dot = fmaf(q[4 * d + 2], k4.z, dot); dot = fmaf(q[4 * d + 3], k4.w, dot);
                }
                const int jy = j / ws, jx = j - jy * ws;
                dot += tbl[bias_row - (jy * span + jx)];
                if (region[j] != my_region) dot += -100.f;
                s[u] = dot;
            } else {
                s[u] = -INFINITY;
            }
        }
        const float cm = fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3]));
        if (cm > m) {
            const float corr = expf(m - cm);   // m = -inf on the first chunk: corr = 0, acc = l = 0
            l *= corr;
#pragma unroll
            for (int d = 0; d < kHeadDim; ++d) acc[d] *= corr;
            m = cm;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u;
            if (j < N) {
                const float p = expf(s[u] - m);
                l += p;
                const float4* vr = reinterpret_cast<const float4*>(Vs + j * kHeadDim);
#pragma unroll
                for (int d = 0; d < kHeadDim / 4; ++d) {
                    const float4 v4 = vr[d];
                    acc[4 * d] = fmaf(p, v4.x, acc[4 * d]); acc[4 * d + 1] = fmaf(p, v4.y, acc[4 * d + 1]);
                    acc[4 * d + 2] = fmaf(p, v4.z, acc[4 * d + 2]); acc[4 * d + 3] = fmaf(p, v4.w, acc[4 * d + 3]);
                }
            }
        }
    }
    const float inv = 1.f / l;
    float* op = out + my_src * C + head * kHeadDim;
#pragma unroll
    for (int d = 0; d < kHeadDim; d += 4)
        *reinterpret_cast<float4*>(op + d) = make_float4(acc[d] * inv, acc[d + 1] * inv, acc[d + 2] * inv, acc[d + 3] * inv);
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

extern "C" int pvsg_window_attention(const float* qkv, const float* qkv_bias, const float* bias_table, float* out, int B,
                                     int H, int W, int C, int heads, int window, int shift, void* stream) {
    PVSG_CHECK_ARG(qkv && qkv_bias && bias_table && out);
    PVSG_CHECK_ARG(B > 0 && H > 0 && W > 0 && heads > 0 && window > 0 && shift >= 0 && shift < window);
    if (window > kMaxWs || C != heads * kHeadDim) return PVSG_ERR_UNSUPPORTED;
    const int nwy = (H + window - 1) / window, nwx = (W + window - 1) / window;
    const int64_t wins = (int64_t)B * nwy * nwx;
    PVSG_CHECK_ARG(wins <= 0x7fffffffLL && heads <= 65535);
    const float scale = 1.f / sqrtf((float)kHeadDim);
    dim3 grid((unsigned)wins, (unsigned)heads);
    window_attention_kernel<<<grid, 160, 0, as_stream(stream)>>>(qkv, qkv_bias, bias_table, out, H, W, C, heads, window,
                                                                shift, nwy, nwx, scale);
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
