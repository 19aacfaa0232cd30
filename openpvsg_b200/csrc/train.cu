// Training forward / backward of the head (first slice): the kernels behind
// Mask2FormerVideoHead.loss_single (models/mask2former_vps/mask2former_video_head.py:196-293) and the backward of the
// pixel decoder's MultiScaleDeformableAttention (mmcv MultiScaleDeformableAttnFunction.backward).
//
//   pvsg_point_sample[_backward]   mmcv.ops.point_sample: bilinear grid_sample (align_corners=False, zero padding) of
//                                  n single-channel maps at K normalised points each
//   pvsg_mask_point_losses         mmdet CrossEntropyLoss(use_sigmoid=True) + DiceLoss(naive_dice=True, eps) on the sampled
//                                  logits / targets: both loss sums and d loss / d logits in one pass
//   pvsg_weighted_ce               mmdet CrossEntropyLoss with class_weight: sum of -w[y] log softmax(x)[y] and its gradient
//   pvsg_msda_backward             gradients w.r.t. value, sampling locations and attention weights
#include "common.cuh"

namespace {

// ---- point_sample ------------------------------------------------------------------------------------------------
// pixel-space sample position of a normalised coordinate p in [0,1]: grid = 2p - 1, align_corners=False:
// x = ((grid + 1) * W - 1) / 2 = p * W - 0.5
__device__ __forceinline__ void ps_corners(float px, float py, int H, int W, int& x0, int& y0, float& lx, float& ly) {
    const float x = px * (float)W - 0.5f, y = py * (float)H - 0.5f;
    const float fx = floorf(x), fy = floorf(y);
    x0 = (int)fx; y0 = (int)fy;
    lx = x - fx; ly = y - fy;
}

__global__ void __launch_bounds__(256) point_sample_kernel(const float* __restrict__ maps, const float* __restrict__ pts,
                                                           float* __restrict__ out, int H, int W, int K, int pts_per_map,
                                                           int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t n = i / K;
    const int k = (int)(i % K);
    const float* p = pts + ((pts_per_map ? n : 0) * K + k) * 2;      // shared points: one [K, 2] set for all maps
    int x0, y0;
    float lx, ly;
    ps_corners(p[0], p[1], H, W, x0, y0, lx, ly);
    const float* m = maps + n * (int64_t)H * W;
    auto at = [&](int y, int x) { return (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(m + (int64_t)y * W + x) : 0.f; };
    out[i] = (1.f - ly) * ((1.f - lx) * at(y0, x0) + lx * at(y0, x0 + 1)) + ly * ((1.f - lx) * at(y0 + 1, x0) + lx * at(y0 + 1, x0 + 1));
}

__global__ void __launch_bounds__(256) point_sample_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ pts,
                                                               float* __restrict__ gmaps, int H, int W, int K, int pts_per_map,
                                                               int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t n = i / K;
    const int k = (int)(i % K);
    const float* p = pts + ((pts_per_map ? n : 0) * K + k) * 2;
    int x0, y0;
    float lx, ly;
    ps_corners(p[0], p[1], H, W, x0, y0, lx, ly);
    float* m = gmaps + n * (int64_t)H * W;
    const float g = gout[i];
    auto add = [&](int y, int x, float w) { if (y >= 0 && y < H && x >= 0 && x < W) atomicAdd(m + (int64_t)y * W + x, w * g); };
    add(y0, x0, (1.f - ly) * (1.f - lx)); add(y0, x0 + 1, (1.f - ly) * lx);
    add(y0 + 1, x0, ly * (1.f - lx)); add(y0 + 1, x0 + 1, ly * lx);
}

// ---- sampled-point mask losses -----------------------------------------------------------------------------------
// one CTA per mask row: BCE-with-logits sum and dice = 1 - (2 sum(s t) + eps) / (sum s + sum t + eps), s = sigmoid(x)
__global__ void __launch_bounds__(256) mask_losses_kernel(const float* __restrict__ x, const float* __restrict__ t, int K, float eps,
                                                          float bce_scale, float dice_scale, float* __restrict__ sums,
                                                          float* __restrict__ grad) {
    const int n = blockIdx.x;
    const float* xr = x + (int64_t)n * K;
    const float* tr = t + (int64_t)n * K;
    float bce = 0.f, a = 0.f, b = 0.f, c = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const float v = xr[k], y = tr[k];
        const float s = 1.f / (1.f + expf(-v));
        bce += fmaxf(v, 0.f) - v * y + log1pf(expf(-fabsf(v)));      // the numerically stable form ATen uses
        a += s * y; b += s; c += y;
    }
    __shared__ float sh[4][8];
    bce = warp_sum(bce); a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    if ((threadIdx.x & 31) == 0) { const int w = threadIdx.x >> 5; sh[0][w] = bce; sh[1][w] = a; sh[2][w] = b; sh[3][w] = c; }
    __syncthreads();
    float tot[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { tot[q] = 0.f; for (int w = 0; w < 8; ++w) tot[q] += sh[q][w]; }
    const float num = 2.f * tot[1] + eps, den = tot[2] + tot[3] + eps;
    if (threadIdx.x == 0) {
        atomicAdd(sums + 0, tot[0]);                 // sum of BCE terms
        atomicAdd(sums + 1, 1.f - num / den);        // sum of dice losses
    }
    if (grad) {
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            const float v = xr[k], y = tr[k];
            const float s = 1.f / (1.f + expf(-v));
            // d dice / d s_k = -(2 y den - num) / den^2 ; d s / d x = s (1 - s)
            const float gd = -(2.f * y * den - num) / (den * den) * s * (1.f - s);
            grad[(int64_t)n * K + k] = bce_scale * (s - y) + dice_scale * gd;
        }
    }
}

// ---- class-weighted cross entropy --------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) weighted_ce_kernel(const float* __restrict__ x, const int64_t* __restrict__ labels,
                                                          const float* __restrict__ cw, const float* __restrict__ lw, int C,
                                                          float scale, float* __restrict__ sums, float* __restrict__ grad) {
    const int r = blockIdx.x;
    const float* xr = x + (int64_t)r * C;
    const int y = (int)labels[r];
    float mx = -INFINITY;
    for (int c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, xr[c]);
    __shared__ float sh[4];
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(sh[0], sh[1]), fmaxf(sh[2], sh[3]));
    __syncthreads();
    float s = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) s += expf(xr[c] - mx);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    s = sh[0] + sh[1] + sh[2] + sh[3];
    const float w = cw[y] * (lw ? lw[r] : 1.f);
    const float lse = mx + logf(s);
    if (threadIdx.x == 0) {
        atomicAdd(sums + 0, w * (lse - xr[y]));      // sum of weighted losses
        atomicAdd(sums + 1, cw[y]);                  // avg_factor = class_weight[labels].sum()
    }
    if (grad)
        for (int c = threadIdx.x; c < C; c += blockDim.x)
            grad[(int64_t)r * C + c] = scale * w * (expf(xr[c] - lse) - (c == y ? 1.f : 0.f));
}

// ---- Hungarian matching cost (mmdet MaskHungarianAssigner: ClassificationCost + CrossEntropyLossCost + DiceCost) ------
// one CTA (32 warps: the loop is a chain of transcendentals, latency-bound with fewer) per query: the per-point terms of the prediction (softplus, sigmoid) are computed once and shared by all ground
// truths; threads stride over the K points, ground truths in chunks of MC_G register accumulators.  (First version: one warp
// per (query, gt) pair = 50 CTAs recomputing the transcendental terms per pair: 139 us per call, 22 ms per training step.)
constexpr int MC_G = 8;
__global__ void __launch_bounds__(1024) match_cost_kernel(const float* __restrict__ cls, const int64_t* __restrict__ labels,
                                                         const float* __restrict__ pred, const float* __restrict__ gt, int Q, int G,
                                                         int C, int K, float w_cls, float w_mask, float w_dice, float eps,
                                                         float* __restrict__ cost) {
    __shared__ float sh[32][3 * MC_G + 1];
    __shared__ float s_mx, s_se;
    const int nwarp = blockDim.x >> 5;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = blockIdx.x;
    // classification cost: -softmax(cls[q])[label_g]   (warp 0)
    const float* xr = cls + (int64_t)q * C;
    if (warp == 0) {
        float mx = -INFINITY;
        for (int c = lane; c < C; c += 32) mx = fmaxf(mx, xr[c]);
        mx = warp_max(mx);
        float se = 0.f;
        for (int c = lane; c < C; c += 32) se += expf(xr[c] - mx);
        se = warp_sum(se);
        if (lane == 0) { s_mx = mx; s_se = se; }
    }
    const float* pr = pred + (int64_t)q * K;
    for (int g0 = 0; g0 < G; g0 += MC_G) {
        const int ng = min(MC_G, G - g0);
        float bce[MC_G], a[MC_G], c2[MC_G], b = 0.f;
#pragma unroll
        for (int j = 0; j < MC_G; ++j) { bce[j] = 0.f; a[j] = 0.f; c2[j] = 0.f; }
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            const float v = pr[k];
            const float sp = log1pf(expf(-fabsf(v)));
            const float pos = fmaxf(-v, 0.f) + sp, neg = fmaxf(v, 0.f) + sp;     // BCE against all-ones / all-zeros targets
            const float s = 1.f / (1.f + expf(-v));
            b += s;
#pragma unroll
            for (int j = 0; j < MC_G; ++j) {
                if (j < ng) {
                    const float y = gt[(int64_t)(g0 + j) * K + k];
                    bce[j] += pos * y + neg * (1.f - y);
                    a[j] = fmaf(s, y, a[j]);
                    c2[j] += y;
                }
            }
        }
        b = warp_sum(b);
#pragma unroll
        for (int j = 0; j < MC_G; ++j) { bce[j] = warp_sum(bce[j]); a[j] = warp_sum(a[j]); c2[j] = warp_sum(c2[j]); }
        __syncthreads();                                   // previous chunk's results consumed; s_mx / s_se written
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < MC_G; ++j) { sh[warp][3 * j] = bce[j]; sh[warp][3 * j + 1] = a[j]; sh[warp][3 * j + 2] = c2[j]; }
            sh[warp][3 * MC_G] = b;
        }
        __syncthreads();
        if (threadIdx.x < ng) {
            const int j = threadIdx.x;
            float tb = 0.f, ta = 0.f, tc = 0.f, tbs = 0.f;
            for (int w = 0; w < nwarp; ++w) { tb += sh[w][3 * j]; ta += sh[w][3 * j + 1]; tc += sh[w][3 * j + 2]; tbs += sh[w][3 * MC_G]; }
            const float prob = expf(xr[labels[g0 + j]] - s_mx) / s_se;
            cost[(int64_t)q * G + g0 + j] = -w_cls * prob + w_mask * tb / (float)K + w_dice * (1.f - (2.f * ta + eps) / (tbs + tc + eps));
        }
    }
}

// ---- MSDeformAttn backward ---------------------------------------------------------------------------------------
constexpr int MAX_LEVELS = 8;
struct BwdLevels {
    int h[MAX_LEVELS], w[MAX_LEVELS];
    int64_t start[MAX_LEVELS];
};

// one warp per (b, query, head); lane = channel (D = 32)
__global__ void __launch_bounds__(256) msda_bwd_kernel(const float* __restrict__ value, BwdLevels lv, const float* __restrict__ loc,
                                                       const float* __restrict__ aw, const float* __restrict__ gout,
                                                       float* __restrict__ gvalue, float* __restrict__ gloc, float* __restrict__ gaw,
                                                       int64_t N, int64_t Nq, int H, int L, int P, int64_t total) {
    const int lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (item >= total) return;
    const int head = (int)(item % H);
    const int64_t bq = item / H;            // b * Nq + q
    const int64_t b = bq / Nq;
    const float g = gout[(bq * H + head) * 32 + lane];
    const float* vb = value + (b * N * H + head) * 32 + lane;
    float* gvb = gvalue + (b * N * H + head) * 32 + lane;
    const int64_t sbase = (bq * H + head) * (int64_t)L * P;
    for (int l = 0; l < L; ++l) {
        const int hgt = lv.h[l], wid = lv.w[l];
        for (int p = 0; p < P; ++p) {
            const int64_t si = sbase + l * P + p;
            const float a = aw[si];
            const float x = loc[si * 2] * (float)wid - 0.5f, y = loc[si * 2 + 1] * (float)hgt - 0.5f;
            float gx = 0.f, gy = 0.f, ga = 0.f;
            if (y > -1.f && x > -1.f && y < (float)hgt && x < (float)wid) {
                const float fy = floorf(y), fx = floorf(x);
                const int y0 = (int)fy, x0 = (int)fx;
                const float ly = y - fy, lx = x - fx, hy = 1.f - ly, hx = 1.f - lx;
                float v[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int yy = y0 + (c >> 1), xx = x0 + (c & 1);
                    const bool in = yy >= 0 && yy < hgt && xx >= 0 && xx < wid;
                    const int64_t off = (lv.start[l] + (int64_t)yy * wid + xx) * H * 32;
                    v[c] = in ? __ldg(vb + off) : 0.f;
                    const float w = ((c >> 1) ? ly : hy) * ((c & 1) ? lx : hx);
                    if (in) atomicAdd(gvb + off, w * a * g);
                }
                ga = g * (hy * (hx * v[0] + lx * v[1]) + ly * (hx * v[2] + lx * v[3]));
                gx = g * a * (hy * (v[1] - v[0]) + ly * (v[3] - v[2])) * (float)wid;
                gy = g * a * (hx * (v[2] - v[0]) + lx * (v[3] - v[1])) * (float)hgt;
            }
            ga = warp_sum(ga); gx = warp_sum(gx); gy = warp_sum(gy);
            if (lane == 0) { gaw[si] = ga; gloc[si * 2] = gx; gloc[si * 2 + 1] = gy; }
        }
    }
}

}  // namespace

extern "C" int pvsg_point_sample(const float* maps, const float* points, float* out, int n, int H, int W, int K,
                                 int points_per_map, void* stream) {
    PVSG_CHECK_ARG(maps && points && out && n > 0 && H > 0 && W > 0 && K > 0);
    const int64_t total = (int64_t)n * K;
    point_sample_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(maps, points, out, H, W, K, points_per_map, total);
    return pvsg_launch_status();
}

extern "C" int pvsg_point_sample_backward(const float* grad_out, const float* points, float* grad_maps, int n, int H, int W,
                                          int K, int points_per_map, void* stream) {
    PVSG_CHECK_ARG(grad_out && points && grad_maps && n > 0 && H > 0 && W > 0 && K > 0);
    if (cudaMemsetAsync(grad_maps, 0, sizeof(float) * (size_t)n * H * W, as_stream(stream)) != cudaSuccess) return PVSG_ERR_LAUNCH;
    const int64_t total = (int64_t)n * K;
    point_sample_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(grad_out, points, grad_maps, H, W, K,
                                                                                            points_per_map, total);
    return pvsg_launch_status();
}

extern "C" int pvsg_mask_point_losses(const float* logits, const float* targets, int n, int K, float dice_eps, float bce_grad_scale,
                                      float dice_grad_scale, float* sums, float* grad, void* stream) {
    PVSG_CHECK_ARG(logits && targets && sums && n > 0 && K > 0);
    if (cudaMemsetAsync(sums, 0, 2 * sizeof(float), as_stream(stream)) != cudaSuccess) return PVSG_ERR_LAUNCH;
    mask_losses_kernel<<<n, 256, 0, as_stream(stream)>>>(logits, targets, K, dice_eps, bce_grad_scale, dice_grad_scale, sums, grad);
    return pvsg_launch_status();
}

extern "C" int pvsg_weighted_ce(const float* logits, const int64_t* labels, const float* class_weight, const float* label_weight,
                                int rows, int C, float grad_scale, float* sums, float* grad, void* stream) {
    PVSG_CHECK_ARG(logits && labels && class_weight && sums && rows > 0 && C > 0);
    if (cudaMemsetAsync(sums, 0, 2 * sizeof(float), as_stream(stream)) != cudaSuccess) return PVSG_ERR_LAUNCH;
    weighted_ce_kernel<<<rows, 128, 0, as_stream(stream)>>>(logits, labels, class_weight, label_weight, C, grad_scale, sums, grad);
    return pvsg_launch_status();
}

extern "C" int pvsg_mask_match_cost(const float* cls_logits, const int64_t* gt_labels, const float* pred_points, const float* gt_points,
                                    int Q, int G, int C, int K, float w_cls, float w_mask, float w_dice, float dice_eps, float* cost,
                                    void* stream) {
    PVSG_CHECK_ARG(cls_logits && gt_labels && pred_points && gt_points && cost && Q > 0 && G > 0 && C > 0 && K > 0);
    match_cost_kernel<<<Q, 1024, 0, as_stream(stream)>>>(cls_logits, gt_labels, pred_points, gt_points, Q, G, C, K, w_cls,
                                                                      w_mask, w_dice, dice_eps, cost);
    return pvsg_launch_status();
}

extern "C" int pvsg_msda_backward(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                                  const float* sampling_locations, const float* attention_weights, const float* grad_out,
                                  float* grad_value, float* grad_loc, float* grad_attn, int B, int64_t N, int64_t Nq, int H, int D,
                                  int L, int P, void* stream) {
    PVSG_CHECK_ARG(value && spatial_shapes && level_start_index && sampling_locations && attention_weights && grad_out &&
                   grad_value && grad_loc && grad_attn && B > 0 && N > 0 && Nq > 0 && H > 0 && P > 0);
    if (D != 32 || L <= 0 || L > MAX_LEVELS) return PVSG_ERR_UNSUPPORTED;
    BwdLevels lv{};
    int64_t tot = 0;
    for (int l = 0; l < L; ++l) {
        lv.h[l] = (int)spatial_shapes[2 * l];
        lv.w[l] = (int)spatial_shapes[2 * l + 1];
        lv.start[l] = level_start_index[l];
        if (lv.h[l] <= 0 || lv.w[l] <= 0 || lv.start[l] != tot) return PVSG_ERR_INVALID_ARG;
        tot += (int64_t)lv.h[l] * lv.w[l];
    }
    if (tot != N) return PVSG_ERR_INVALID_ARG;
    if (cudaMemsetAsync(grad_value, 0, sizeof(float) * (size_t)B * N * H * 32, as_stream(stream)) != cudaSuccess) return PVSG_ERR_LAUNCH;
    const int64_t total = (int64_t)B * Nq * H;
    msda_bwd_kernel<<<(unsigned)((total + 7) / 8), 256, 0, as_stream(stream)>>>(value, lv, sampling_locations, attention_weights, grad_out,
                                                                                grad_value, grad_loc, grad_attn, N, Nq, H, L, P, total);
    return pvsg_launch_status();
}
