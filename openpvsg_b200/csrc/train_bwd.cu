// Training of the transformer decoder head (second slice): the backward kernels that, together with
// the forward engine (pvsg_linear for every dense product, forward AND backward), make
// Mask2FormerVideoHead.forward_train (models/mask2former_vps/mask2former_video_head.py:464-522) run on the device:
//
//   pvsg_layernorm_backward      nn.LayerNorm backward: dx, and dgamma / dbeta accumulated over the rows
//   pvsg_relu_backward           dx = dy where the forward output was positive (FFN / mask-embed MLP)
//   pvsg_colsum                  out[n] = sum_m x[m, n]: bias gradients, level_embed, the broadcast of the query embeddings
//   pvsg_attention_train_forward masked multi-head attention that also returns the row log-sum-exp (exact fp32, SIMT)
//   pvsg_attention_train_backward dq, dk, dv from (q, k, v, o, do, lse) with the same mask -- nn.MultiheadAttention's
//                                softmax(q k^T / sqrt(d) + mask) v as mmcv's MultiheadAttention wrapper calls it
//                                (mask2former_head.py:457-468); recomputes the probabilities, never stores them
//
// First cut, sized for correctness at training shapes (batch 1-4, 100 queries, <= 15 k keys, head dim 32): exact fp32
// SIMT, deterministic (no atomics except the column sums).
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------ LayerNorm backward
// One warp per row (rows strided over the grid); lane owns columns lane, lane + 32, ...  CPL = C / 32 columns per lane.
template <int CPL>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ dy, float* __restrict__ dx,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t M, float eps) {
    constexpr int C = 32 * CPL;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    float g[CPL], dg[CPL], db[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) { g[i] = gamma[lane + 32 * i]; dg[i] = 0.f; db[i] = 0.f; }
    for (int64_t r = (int64_t)blockIdx.x * nwarp + warp; r < M; r += (int64_t)gridDim.x * nwarp) {
        float xv[CPL], dv[CPL];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < CPL; ++i) { xv[i] = x[r * C + lane + 32 * i]; dv[i] = dy[r * C + lane + 32 * i]; s += xv[i]; }
        const float mean = warp_sum(s) * (1.f / C);
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < CPL; ++i) { xv[i] -= mean; v = fmaf(xv[i], xv[i], v); }
        const float rstd = rsqrtf(warp_sum(v) * (1.f / C) + eps);
        float sg = 0.f, sgx = 0.f;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            xv[i] *= rstd;                         // xhat
            const float gi = dv[i] * g[i];
            sg += gi;
            sgx = fmaf(gi, xv[i], sgx);
            dg[i] = fmaf(dv[i], xv[i], dg[i]);
            db[i] += dv[i];
        }
        sg = warp_sum(sg) * (1.f / C);
        sgx = warp_sum(sgx) * (1.f / C);
#pragma unroll
        for (int i = 0; i < CPL; ++i) dx[r * C + lane + 32 * i] = rstd * (dv[i] * g[i] - sg - xv[i] * sgx);
    }
    // CTA reduction of the parameter gradients, then one atomic per column and CTA
    __shared__ float sh[2][8][C];
#pragma unroll
    for (int i = 0; i < CPL; ++i) { sh[0][warp][lane + 32 * i] = dg[i]; sh[1][warp][lane + 32 * i] = db[i]; }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < nwarp; ++w) { a += sh[0][w][c]; b += sh[1][w][c]; }
        atomicAdd(dgamma + c, a);
        atomicAdd(dbeta + c, b);
    }
}

// any C: one warp per row, parameter gradients by direct atomics (slow path, small tensors only)
__global__ void __launch_bounds__(256) ln_bwd_generic_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ dy, float* __restrict__ dx,
                                                             float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t M,
                                                             int C, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= M) return;
    const float* xr = x + r * C;
    const float* dr = dy + r * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += xr[c];
    const float mean = warp_sum(s) / (float)C;
    float v = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = xr[c] - mean; v = fmaf(d, d, v); }
    const float rstd = rsqrtf(warp_sum(v) / (float)C + eps);
    float sg = 0.f, sgx = 0.f;
    for (int c = lane; c < C; c += 32) {
        const float xh = (xr[c] - mean) * rstd, gi = dr[c] * gamma[c];
        sg += gi;
        sgx = fmaf(gi, xh, sgx);
    }
    sg = warp_sum(sg) / (float)C;
    sgx = warp_sum(sgx) / (float)C;
    for (int c = lane; c < C; c += 32) {
        const float xh = (xr[c] - mean) * rstd;
        dx[r * C + c] = rstd * (dr[c] * gamma[c] - sg - xh * sgx);
        atomicAdd(dgamma + c, dr[c] * xh);
        atomicAdd(dbeta + c, dr[c]);
    }
}

__global__ void __launch_bounds__(256) relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                       float* __restrict__ dx, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dx[i] = y[i] > 0.f ? dy[i] : 0.f;
}

// out[n] (+)= sum over a slab of rows; block (32, 8), grid (ceil(N/32), row slabs)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t M, int N,
                                                     int64_t ld, int64_t rows_per_slab) {
    __shared__ float sh[8][33];
    const int col = blockIdx.x * 32 + threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_slab;
    const int64_t r1 = r0 + rows_per_slab < M ? r0 + rows_per_slab : M;
    float s = 0.f;
    if (col < N)
        for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) s += x[r * ld + col];
    sh[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && col < N) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += sh[w][threadIdx.x];
        atomicAdd(out + col, t);
    }
}

// ------------------------------------------------------------------------------------------ attention (training)
constexpr int AD = 32;      // head dim

struct AttnT {
    const float *q, *k, *v;
    const uint8_t* mask;        // [B, Lq, Lk], non-zero = blocked (shared by the heads), or null
    const int32_t* row_open;    // [B, Lq]: number of open keys of the row (0 -> the row ignores the mask), or null
    int B, H, Lq, Lk;
    int64_t qb, qr, kb, kr, vb, vr;   // batch / row strides (floats)
    float scale;
};

__device__ __forceinline__ void load_row(const float* p, float (&r)[AD]) {
#pragma unroll
    for (int i = 0; i < AD / 4; ++i) {
        const float4 t = *reinterpret_cast<const float4*>(p + 4 * i);
        r[4 * i] = t.x; r[4 * i + 1] = t.y; r[4 * i + 2] = t.z; r[4 * i + 3] = t.w;
    }
}
__device__ __forceinline__ float dot_row(const float* p, const float (&r)[AD]) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < AD / 4; ++i) {
        const float4 t = *reinterpret_cast<const float4*>(p + 4 * i);
        s = fmaf(t.x, r[4 * i], s); s = fmaf(t.y, r[4 * i + 1], s); s = fmaf(t.z, r[4 * i + 2], s); s = fmaf(t.w, r[4 * i + 3], s);
    }
    return s;
}
// after the call lane d holds the warp-wide sum of acc[d]
__device__ __forceinline__ float reduce_rows(float (&acc)[AD], int lane) {
    float mine = 0.f;
#pragma unroll
    for (int d = 0; d < AD; ++d) {
        const float t = warp_sum(acc[d]);
        if (lane == d) mine = t;
    }
    return mine;
}

// one warp per (b, h, query): lanes stride over the keys
__global__ void __launch_bounds__(128) attn_train_fwd_kernel(AttnT a, float* __restrict__ out, int64_t ob, int64_t orow,
                                                             float* __restrict__ lse) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= (int64_t)a.B * a.H * a.Lq) return;
    const int qi = (int)(w % a.Lq), h = (int)((w / a.Lq) % a.H), b = (int)(w / ((int64_t)a.Lq * a.H));
    float q[AD];
    load_row(a.q + b * a.qb + qi * a.qr + h * AD, q);
    const uint8_t* m = a.mask ? a.mask + ((int64_t)b * a.Lq + qi) * a.Lk : nullptr;
    if (m && a.row_open && a.row_open[b * a.Lq + qi] == 0) m = nullptr;
    const float* kp = a.k + b * a.kb + h * AD;
    const float* vp = a.v + b * a.vb + h * AD;
    float mx = -INFINITY;
    for (int j = lane; j < a.Lk; j += 32)
        if (!m || !m[j]) mx = fmaxf(mx, a.scale * dot_row(kp + j * a.kr, q));
    mx = warp_max(mx);
    float l = 0.f, acc[AD];
#pragma unroll
    for (int d = 0; d < AD; ++d) acc[d] = 0.f;
    for (int j = lane; j < a.Lk; j += 32) {
        if (m && m[j]) continue;
        const float p = expf(a.scale * dot_row(kp + j * a.kr, q) - mx);
        l += p;
        const float* vr = vp + j * a.vr;
#pragma unroll
        for (int i = 0; i < AD / 4; ++i) {
            const float4 t = *reinterpret_cast<const float4*>(vr + 4 * i);
            acc[4 * i] = fmaf(p, t.x, acc[4 * i]); acc[4 * i + 1] = fmaf(p, t.y, acc[4 * i + 1]);
            acc[4 * i + 2] = fmaf(p, t.z, acc[4 * i + 2]); acc[4 * i + 3] = fmaf(p, t.w, acc[4 * i + 3]);
        }
    }
    l = warp_sum(l);
    const float o = reduce_rows(acc, lane);
    out[b * ob + qi * orow + h * AD + lane] = o / l;
    if (lane == 0) lse[w] = mx + logf(l);
}

// dq: one warp per (b, h, query); also writes delta = <do, o> for the dk / dv pass
__global__ void __launch_bounds__(128) attn_train_dq_kernel(AttnT a, const float* __restrict__ o, int64_t ob, int64_t orow,
                                                            const float* __restrict__ dout, int64_t dob, int64_t dor,
                                                            const float* __restrict__ lse, float* __restrict__ delta,
                                                            float* __restrict__ dq, int64_t dqb, int64_t dqr) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= (int64_t)a.B * a.H * a.Lq) return;
    const int qi = (int)(w % a.Lq), h = (int)((w / a.Lq) % a.H), b = (int)(w / ((int64_t)a.Lq * a.H));
    float q[AD], g[AD], acc[AD];
    load_row(a.q + b * a.qb + qi * a.qr + h * AD, q);
    load_row(dout + b * dob + qi * dor + h * AD, g);
    const float dl = dot_row(o + b * ob + qi * orow + h * AD, g);
    if (lane == 0) delta[w] = dl;
    const float ls = lse[w];
    const uint8_t* m = a.mask ? a.mask + ((int64_t)b * a.Lq + qi) * a.Lk : nullptr;
    if (m && a.row_open && a.row_open[b * a.Lq + qi] == 0) m = nullptr;
    const float* kp = a.k + b * a.kb + h * AD;
    const float* vp = a.v + b * a.vb + h * AD;
#pragma unroll
    for (int d = 0; d < AD; ++d) acc[d] = 0.f;
    for (int j = lane; j < a.Lk; j += 32) {
        if (m && m[j]) continue;
        const float* kr = kp + j * a.kr;
        const float p = expf(a.scale * dot_row(kr, q) - ls);
        const float ds = p * (dot_row(vp + j * a.vr, g) - dl) * a.scale;
#pragma unroll
        for (int i = 0; i < AD / 4; ++i) {
            const float4 t = *reinterpret_cast<const float4*>(kr + 4 * i);
            acc[4 * i] = fmaf(ds, t.x, acc[4 * i]); acc[4 * i + 1] = fmaf(ds, t.y, acc[4 * i + 1]);
            acc[4 * i + 2] = fmaf(ds, t.z, acc[4 * i + 2]); acc[4 * i + 3] = fmaf(ds, t.w, acc[4 * i + 3]);
        }
    }
    dq[b * dqb + qi * dqr + h * AD + lane] = reduce_rows(acc, lane);
}

// dk, dv: one warp per (b, h, 32 keys), lane = key: the lane keeps its key / value rows and both gradient rows in registers
// and loops over the queries, whose q / dO rows are the same address for every lane (broadcast loads, L1-resident: Lq rows
// per head).  4 D FMAs per (query, key) and no shuffles.  (First version: lane = channel with two warp reductions per
// (query, key): 21 ms per training step.)
__global__ void __launch_bounds__(128) attn_train_dkv_kernel(AttnT a, const float* __restrict__ dout, int64_t dob, int64_t dor,
                                                             const float* __restrict__ lse, const float* __restrict__ delta,
                                                             float* __restrict__ dk, int64_t dkb, int64_t dkr,
                                                             float* __restrict__ dv, int64_t dvb, int64_t dvr) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int ktiles = (a.Lk + 31) / 32;
    if (w >= (int64_t)a.B * a.H * ktiles) return;
    const int kt = (int)(w % ktiles), h = (int)((w / ktiles) % a.H), b = (int)(w / ((int64_t)ktiles * a.H));
    const int j = kt * 32 + lane;
    const bool valid = j < a.Lk;
    const int jj = valid ? j : a.Lk - 1;
    float kr[AD], vr[AD], gk[AD], gv[AD];
    load_row(a.k + b * a.kb + (int64_t)jj * a.kr + h * AD, kr);
    load_row(a.v + b * a.vb + (int64_t)jj * a.vr + h * AD, vr);
#pragma unroll
    for (int d = 0; d < AD; ++d) { gk[d] = 0.f; gv[d] = 0.f; }
    const float* qp = a.q + b * a.qb + h * AD;
    const float* gp = dout + b * dob + h * AD;
    const float* lp = lse + ((int64_t)b * a.H + h) * a.Lq;
    const float* dp = delta + ((int64_t)b * a.H + h) * a.Lq;
    for (int qi = 0; qi < a.Lq; ++qi) {
        const bool open = !a.mask || !a.mask[((int64_t)b * a.Lq + qi) * a.Lk + jj] || (a.row_open && a.row_open[b * a.Lq + qi] == 0);
        const float* qrow = qp + (int64_t)qi * a.qr;
        const float* grow = gp + (int64_t)qi * dor;
        float s = 0.f, dpv = 0.f;
        float qv[AD], gvv[AD];
        load_row(qrow, qv);
        load_row(grow, gvv);
#pragma unroll
        for (int d = 0; d < AD; ++d) { s = fmaf(qv[d], kr[d], s); dpv = fmaf(gvv[d], vr[d], dpv); }
        const float p = open ? expf(a.scale * s - lp[qi]) : 0.f;
        const float ds = p * (dpv - dp[qi]) * a.scale;
#pragma unroll
        for (int d = 0; d < AD; ++d) { gv[d] = fmaf(p, gvv[d], gv[d]); gk[d] = fmaf(ds, qv[d], gk[d]); }
    }
    if (valid) {
        float* ko = dk + b * dkb + (int64_t)j * dkr + h * AD;
        float* vo = dv + b * dvb + (int64_t)j * dvr + h * AD;
#pragma unroll
        for (int i = 0; i < AD / 4; ++i) {
            *reinterpret_cast<float4*>(ko + 4 * i) = make_float4(gk[4 * i], gk[4 * i + 1], gk[4 * i + 2], gk[4 * i + 3]);
            *reinterpret_cast<float4*>(vo + 4 * i) = make_float4(gv[4 * i], gv[4 * i + 1], gv[4 * i + 2], gv[4 * i + 3]);
        }
    }
}

// ------------------------------------------------------------------------------------------ GroupNorm backward (NHWC)
// x, dy [B, HW, C], G groups of cpg = C / G channels.  Kernel A: one CTA per (b, g) recomputes mean / rstd, then the two
// group sums of the backward formula and this group's share of dgamma / dbeta (atomics over b).  Kernel B: dx.
// relu != 0: the forward applied ReLU after the affine map; dy is masked where y = xhat * gamma + beta <= 0.
__device__ __forceinline__ float block_sum(float v, float* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < nw; ++w) t += sh[w];
    return t;
}

__global__ void __launch_bounds__(256) gn_bwd_stats_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, const float* __restrict__ dy,
                                                           float* __restrict__ stats, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, int64_t HW, int C, int G, float eps, int relu) {
    __shared__ float sh[8];
    __shared__ float chan[2][32];       // cpg <= 32
    const int b = blockIdx.x / G, g = blockIdx.x % G, cpg = C / G;
    const int64_t n = HW * cpg;
    const float* xb = x + (int64_t)b * HW * C + g * cpg;
    const float* db = dy + (int64_t)b * HW * C + g * cpg;
    float s = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += xb[(i / cpg) * C + (i % cpg)];
    const float mean = block_sum(s, sh) / (float)n;
    float v = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { const float d = xb[(i / cpg) * C + (i % cpg)] - mean; v = fmaf(d, d, v); }
    const float rstd = rsqrtf(block_sum(v, sh) / (float)n + eps);
    if (threadIdx.x < 64) chan[threadIdx.x >> 5][threadIdx.x & 31] = 0.f;
    __syncthreads();
    float sg = 0.f, sgx = 0.f, pg = 0.f, pb = 0.f;
    // when cpg divides the block size a thread only ever visits channel threadIdx.x % cpg: register partials, one
    // shared-memory atomic per thread at the end
    const bool fixed = blockDim.x % cpg == 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const int c = (int)(i % cpg);
        const int64_t off = (i / cpg) * C + c;
        const float xh = (xb[off] - mean) * rstd;
        const float ga = gamma[g * cpg + c];
        float d = db[off];
        if (relu && fmaf(xh, ga, beta[g * cpg + c]) <= 0.f) d = 0.f;
        const float gi = d * ga;
        sg += gi;
        sgx = fmaf(gi, xh, sgx);
        if (fixed) {
            pg = fmaf(d, xh, pg);
            pb += d;
        } else {
            atomicAdd(&chan[0][c], d * xh);
            atomicAdd(&chan[1][c], d);
        }
    }
    if (fixed) {
        atomicAdd(&chan[0][threadIdx.x % cpg], pg);
        atomicAdd(&chan[1][threadIdx.x % cpg], pb);
    }
    sg = block_sum(sg, sh);
    sgx = block_sum(sgx, sh);
    __syncthreads();
    if (threadIdx.x == 0) {
        float* st = stats + (int64_t)blockIdx.x * 4;
        st[0] = mean; st[1] = rstd; st[2] = sg / (float)n; st[3] = sgx / (float)n;
    }
    if (threadIdx.x < cpg) {
        atomicAdd(dgamma + g * cpg + threadIdx.x, chan[0][threadIdx.x]);
        atomicAdd(dbeta + g * cpg + threadIdx.x, chan[1][threadIdx.x]);
    }
}

__global__ void __launch_bounds__(256) gn_bwd_dx_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, const float* __restrict__ dy,
                                                        const float* __restrict__ stats, float* __restrict__ dx, int64_t HW,
                                                        int C, int G, int relu, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C);
    const int64_t b = i / (HW * C);
    const int cpg = C / G;
    const float* st = stats + (b * G + c / cpg) * 4;
    const float xh = (x[i] - st[0]) * st[1];
    float d = dy[i];
    if (relu && fmaf(xh, gamma[c], beta[c]) <= 0.f) d = 0.f;
    dx[i] = st[1] * (d * gamma[c] - st[2] - xh * st[3]);
}

// ------------------------------------------------------------------------------------------ bilinear resize backward
// adjoint of bilinear_nhwc (align_corners=False): every output pixel scatters its gradient to its 4 source taps
__global__ void __launch_bounds__(256) resize_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dsrc, int B, int IH,
                                                         int IW, int OH, int OW, int C, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C);
    const int ox = (int)((i / C) % OW), oy = (int)((i / ((int64_t)C * OW)) % OH);
    const int64_t b = i / ((int64_t)C * OW * OH);
    int y0, y1, x0, x1;
    float wy0, wy1, wx0, wx1;
    bilinear_coord(oy, (float)IH / (float)OH, IH, y0, y1, wy0, wy1);
    bilinear_coord(ox, (float)IW / (float)OW, IW, x0, x1, wx0, wx1);
    const float g = dout[i];
    float* base = dsrc + b * (int64_t)IH * IW * C + c;
    atomicAdd(base + ((int64_t)y0 * IW + x0) * C, g * wy0 * wx0);
    atomicAdd(base + ((int64_t)y0 * IW + x1) * C, g * wy0 * wx1);
    atomicAdd(base + ((int64_t)y1 * IW + x0) * C, g * wy1 * wx0);
    atomicAdd(base + ((int64_t)y1 * IW + x1) * C, g * wy1 * wx1);
}

// ------------------------------------------------------------------------------------------ transpose + split
// Operand preparation of the weight-gradient GEMMs (dW = dZ^T X: the reduction runs over the tokens, so both operands
// are needed token-minor): source rows r = (b, oh, ow) of a (possibly strided / shifted) token-major view [.., C] ->
// bf16 (hi, lo) planes [C, ldt] with column r; one pass instead of transpose + split (+ the optional add of a second
// source with the same layout).
__global__ void __launch_bounds__(256) transpose_split_kernel(const float* __restrict__ x, const float* __restrict__ add,
                                                              __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                              int64_t T, int C, int OH, int OW, int64_t sb, int64_t sh, int64_t sw,
                                                              int64_t ldt) {
    // tile: 64 source rows x 32 channels.  Loads: a warp reads the 32 channels of one row (128 B).  Stores: a warp writes
    // 64 consecutive columns of one channel row as bf16x2 (128 B); rows beyond T are written as zeros (ldt >= ceil64(T)).
    __shared__ float tile[64][33];
    __shared__ int64_t row_off[64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * 64;
    const int c0 = blockIdx.y * 32;
    if (threadIdx.x < 64) {            // one (b, oh, ow) decomposition per source row, 32-bit when it fits
        const int64_t r = r0 + threadIdx.x;
        int64_t off = -1;
        if (r < T) {
            if (T < (1LL << 31)) {
                const unsigned ru = (unsigned)r, hw = (unsigned)(OH * OW), b = ru / hw, rem = ru - b * hw, oh = rem / (unsigned)OW;
                off = (int64_t)b * sb + (int64_t)oh * sh + (int64_t)(rem - oh * (unsigned)OW) * sw;
            } else {
                const int64_t hw = (int64_t)OH * OW, b = r / hw, rem = r - b * hw;
                off = b * sb + (rem / OW) * sh + (rem % OW) * sw;
            }
        }
        row_off[threadIdx.x] = off;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = warp + 8 * i;
        const int64_t off = row_off[row];
        float v = 0.f;
        if (off >= 0 && c0 + lane < C) {
            v = x[off + c0 + lane];
            if (add) v += add[off + c0 + lane];
        }
        tile[row][lane] = v;
    }
    __syncthreads();
    if (r0 + 2 * lane >= ldt) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = warp + 8 * i;
        if (c0 + c >= C) continue;
        const float v0 = tile[2 * lane][c], v1 = tile[2 * lane + 1][c];
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
        const int64_t o = (int64_t)(c0 + c) * ldt + r0 + 2 * lane;
        *reinterpret_cast<__nv_bfloat162*>(hi + o) = __nv_bfloat162(h0, h1);
        *reinterpret_cast<__nv_bfloat162*>(lo + o) = __nv_bfloat162(__float2bfloat16_rn(v0 - __bfloat162float(h0)),
                                                                    __float2bfloat16_rn(v1 - __bfloat162float(h1)));
    }
}

// ------------------------------------------------------------------------------------------ max pooling backward
// 3x3 / stride 2 / pad 1: every output routes its gradient to the FIRST maximum of its window in (row, column) scan order
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                          float* __restrict__ dx, int H, int W, int OH, int OW, int C, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C);
    const int ox = (int)((i / C) % OW), oy = (int)((i / ((int64_t)C * OW)) % OH);
    const int64_t b = i / ((int64_t)C * OW * OH);
    const float* xb = x + b * (int64_t)H * W * C + c;
    float best = -INFINITY;
    int64_t arg = -1;
    for (int r = 0; r < 3; ++r) {
        const int y = 2 * oy - 1 + r;
        if (y < 0 || y >= H) continue;
        for (int s = 0; s < 3; ++s) {
            const int xx = 2 * ox - 1 + s;
            if (xx < 0 || xx >= W) continue;
            const float v = xb[((int64_t)y * W + xx) * C];
            if (v > best || arg < 0) { best = v; arg = (int64_t)y * W + xx; }
        }
    }
    atomicAdd(dx + b * (int64_t)H * W * C + arg * C + c, dy[i]);
}

// ------------------------------------------------------------------------------------------ MSDeformAttn projections
// proj [B*Nq, H*L*P*3] = offsets (H, L, P, 2) | logits (H, L, P)  <->  sampling locations [.., H, L, P, 2] and
// attention weights [.., H, L, P] (softmax over L*P), as MultiScaleDeformableAttention.forward computes them.
struct LevelWH { float w[8], h[8]; };

__global__ void __launch_bounds__(256) msda_expand_kernel(const float* __restrict__ proj, const float* __restrict__ ref,
                                                          float* __restrict__ loc, float* __restrict__ aw, int64_t rows, int64_t Nq,
                                                          int H, int L, int P, LevelWH lv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * H) return;
    const int h = (int)(i % H);
    const int64_t r = i / H;
    const int LP = L * P;
    const float* off = proj + r * (int64_t)H * LP * 3 + (int64_t)h * LP * 2;
    const float* lg = proj + r * (int64_t)H * LP * 3 + (int64_t)H * LP * 2 + (int64_t)h * LP;
    const float rx = ref[(r % Nq) * 2], ry = ref[(r % Nq) * 2 + 1];
    float mx = -INFINITY;
    for (int k = 0; k < LP; ++k) mx = fmaxf(mx, lg[k]);
    float sum = 0.f;
    for (int k = 0; k < LP; ++k) sum += expf(lg[k] - mx);
    for (int k = 0; k < LP; ++k) {
        const int l = k / P;
        loc[(i * LP + k) * 2] = rx + off[2 * k] / lv.w[l];
        loc[(i * LP + k) * 2 + 1] = ry + off[2 * k + 1] / lv.h[l];
        aw[i * LP + k] = expf(lg[k] - mx) / sum;
    }
}

__global__ void __launch_bounds__(256) msda_proj_bwd_kernel(const float* __restrict__ aw, const float* __restrict__ dloc,
                                                            const float* __restrict__ daw, float* __restrict__ dproj,
                                                            int64_t rows, int H, int L, int P, LevelWH lv) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * H) return;
    const int h = (int)(i % H);
    const int64_t r = i / H;
    const int LP = L * P;
    float* doff = dproj + r * (int64_t)H * LP * 3 + (int64_t)h * LP * 2;
    float* dlg = dproj + r * (int64_t)H * LP * 3 + (int64_t)H * LP * 2 + (int64_t)h * LP;
    float dot = 0.f;
    for (int k = 0; k < LP; ++k) dot = fmaf(aw[i * LP + k], daw[i * LP + k], dot);
    for (int k = 0; k < LP; ++k) {
        const int l = k / P;
        doff[2 * k] = dloc[(i * LP + k) * 2] / lv.w[l];
        doff[2 * k + 1] = dloc[(i * LP + k) * 2 + 1] / lv.h[l];
        dlg[k] = aw[i * LP + k] * (daw[i * LP + k] - dot);
    }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

inline bool attn_args_ok(const AttnT& a) {
    return a.q && a.k && a.v && a.B > 0 && a.H > 0 && a.Lq > 0 && a.Lk > 0 && al16(a.q) && al16(a.k) && al16(a.v) &&
           a.qb % 4 == 0 && a.qr % 4 == 0 && a.kb % 4 == 0 && a.kr % 4 == 0 && a.vb % 4 == 0 && a.vr % 4 == 0;
}

}  // namespace

extern "C" int pvsg_layernorm_backward(const float* x, const float* gamma, const float* dy, float* dx, float* dgamma,
                                       float* dbeta, int64_t M, int C, float eps, void* stream) {
    PVSG_CHECK_ARG(x && gamma && dy && dx && dgamma && dbeta && M > 0 && C > 0);
    cudaStream_t st = as_stream(stream);
    if (cudaMemsetAsync(dgamma, 0, sizeof(float) * C, st) != cudaSuccess || cudaMemsetAsync(dbeta, 0, sizeof(float) * C, st) != cudaSuccess)
        return PVSG_ERR_LAUNCH;
    if (C == 256) {
        const unsigned grid = (unsigned)imin64((M + 7) / 8, 148 * 4);
        ln_bwd_kernel<8><<<grid, 256, 0, st>>>(x, gamma, dy, dx, dgamma, dbeta, M, eps);
    } else {
        ln_bwd_generic_kernel<<<(unsigned)((M + 7) / 8), 256, 0, st>>>(x, gamma, dy, dx, dgamma, dbeta, M, C, eps);
    }
    return pvsg_launch_status();
}

extern "C" int pvsg_relu_backward(const float* dy, const float* y, float* dx, int64_t n, void* stream) {
    PVSG_CHECK_ARG(dy && y && dx && n > 0);
    relu_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(dy, y, dx, n);
    return pvsg_launch_status();
}

extern "C" int pvsg_colsum(const float* x, float* out, int64_t M, int N, int64_t ld, void* stream) {
    PVSG_CHECK_ARG(x && out && M > 0 && N > 0 && ld >= N);
    cudaStream_t st = as_stream(stream);
    if (cudaMemsetAsync(out, 0, sizeof(float) * N, st) != cudaSuccess) return PVSG_ERR_LAUNCH;
    const int64_t col_blocks = (N + 31) / 32;
    int64_t slabs = imin64((M + 63) / 64, (148 * 8 + col_blocks - 1) / col_blocks);
    if (slabs < 1) slabs = 1;
    if (slabs > 65535) slabs = 65535;
    const int64_t rows_per_slab = (M + slabs - 1) / slabs;
    colsum_kernel<<<dim3((unsigned)col_blocks, (unsigned)slabs), dim3(32, 8), 0, st>>>(x, out, M, N, ld, rows_per_slab);
    return pvsg_launch_status();
}

extern "C" int pvsg_attention_train_forward(const float* q, const float* k, const float* v, const uint8_t* mask,
                                            const int32_t* row_open, float* out, float* lse, int B, int H, int Lq, int Lk,
                                            int D, int64_t q_bs, int64_t q_rs, int64_t k_bs, int64_t k_rs, int64_t v_bs,
                                            int64_t v_rs, int64_t o_bs, int64_t o_rs, float scale, void* stream) {
    AttnT a{q, k, v, mask, row_open, B, H, Lq, Lk, q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, scale};
    PVSG_CHECK_ARG(out && lse && attn_args_ok(a));
    if (D != AD) return PVSG_ERR_UNSUPPORTED;
    const int64_t warps = (int64_t)B * H * Lq;
    attn_train_fwd_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, as_stream(stream)>>>(a, out, o_bs, o_rs, lse);
    return pvsg_launch_status();
}

extern "C" int pvsg_attention_train_backward(const float* q, const float* k, const float* v, const uint8_t* mask,
                                             const int32_t* row_open, const float* out, const float* dout, const float* lse,
                                             float* delta, float* dq, float* dk, float* dv, int B, int H, int Lq, int Lk, int D,
                                             int64_t q_bs, int64_t q_rs, int64_t k_bs, int64_t k_rs, int64_t v_bs, int64_t v_rs,
                                             int64_t o_bs, int64_t o_rs, float scale, void* stream) {
    AttnT a{q, k, v, mask, row_open, B, H, Lq, Lk, q_bs, q_rs, k_bs, k_rs, v_bs, v_rs, scale};
    PVSG_CHECK_ARG(out && dout && lse && delta && dq && dk && dv && attn_args_ok(a) && al16(out) && al16(dout) &&
                   o_bs % 4 == 0 && o_rs % 4 == 0);
    if (D != AD) return PVSG_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    const int E = H * AD;
    // gradients are written contiguous: dq [B, Lq, E], dk / dv [B, Lk, E]; dout shares the layout of out
    const int64_t wq = (int64_t)B * H * Lq, wk = (int64_t)B * H * ((Lk + 31) / 32);
    attn_train_dq_kernel<<<(unsigned)((wq + 3) / 4), 128, 0, st>>>(a, out, o_bs, o_rs, dout, o_bs, o_rs, lse, delta, dq,
                                                                   (int64_t)Lq * E, E);
    attn_train_dkv_kernel<<<(unsigned)((wk + 3) / 4), 128, 0, st>>>(a, dout, o_bs, o_rs, lse, delta, dk, (int64_t)Lk * E, E, dv,
                                                                    (int64_t)Lk * E, E);
    return pvsg_launch_status();
}

extern "C" int pvsg_groupnorm_nhwc_backward(const float* x, const float* gamma, const float* beta, const float* dy, float* dx,
                                            float* dgamma, float* dbeta, float* stats, int B, int64_t HW, int C, int G, float eps,
                                            int relu, void* stream) {
    PVSG_CHECK_ARG(x && gamma && beta && dy && dx && dgamma && dbeta && stats && B > 0 && HW > 0 && C > 0 && G > 0 && C % G == 0);
    if (C / G > 32) return PVSG_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    if (cudaMemsetAsync(dgamma, 0, sizeof(float) * C, st) != cudaSuccess || cudaMemsetAsync(dbeta, 0, sizeof(float) * C, st) != cudaSuccess)
        return PVSG_ERR_LAUNCH;
    gn_bwd_stats_kernel<<<(unsigned)(B * G), 256, 0, st>>>(x, gamma, beta, dy, stats, dgamma, dbeta, HW, C, G, eps, relu);
    const int64_t total = (int64_t)B * HW * C;
    gn_bwd_dx_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, gamma, beta, dy, stats, dx, HW, C, G, relu, total);
    return pvsg_launch_status();
}

extern "C" int pvsg_bilinear_resize_nhwc_backward(const float* dout, float* dsrc, int B, int IH, int IW, int OH, int OW, int C,
                                                  void* stream) {
    PVSG_CHECK_ARG(dout && dsrc && B > 0 && IH > 0 && IW > 0 && OH > 0 && OW > 0 && C > 0);
    cudaStream_t st = as_stream(stream);
    if (cudaMemsetAsync(dsrc, 0, sizeof(float) * (size_t)B * IH * IW * C, st) != cudaSuccess) return PVSG_ERR_LAUNCH;
    const int64_t total = (int64_t)B * OH * OW * C;
    resize_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dout, dsrc, B, IH, IW, OH, OW, C, total);
    return pvsg_launch_status();
}

static bool level_wh(const int64_t* spatial_shapes, int L, LevelWH& lv) {
    if (!spatial_shapes || L < 1 || L > 8) return false;
    for (int l = 0; l < L; ++l) {
        lv.h[l] = (float)spatial_shapes[2 * l];
        lv.w[l] = (float)spatial_shapes[2 * l + 1];
        if (lv.h[l] <= 0.f || lv.w[l] <= 0.f) return false;
    }
    return true;
}

extern "C" int pvsg_msda_proj_expand(const float* proj, const float* ref, const int64_t* spatial_shapes, float* loc, float* aw,
                                     int B, int64_t Nq, int H, int L, int P, void* stream) {
    LevelWH lv;
    PVSG_CHECK_ARG(proj && ref && loc && aw && B > 0 && Nq > 0 && H > 0 && P > 0 && level_wh(spatial_shapes, L, lv));
    const int64_t n = (int64_t)B * Nq * H;
    msda_expand_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(proj, ref, loc, aw, (int64_t)B * Nq, Nq, H, L, P, lv);
    return pvsg_launch_status();
}

extern "C" int pvsg_msda_proj_backward(const float* aw, const float* dloc, const float* daw, const int64_t* spatial_shapes,
                                       float* dproj, int B, int64_t Nq, int H, int L, int P, void* stream) {
    LevelWH lv;
    PVSG_CHECK_ARG(aw && dloc && daw && dproj && B > 0 && Nq > 0 && H > 0 && P > 0 && level_wh(spatial_shapes, L, lv));
    const int64_t n = (int64_t)B * Nq * H;
    msda_proj_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(aw, dloc, daw, dproj, (int64_t)B * Nq, H, L, P, lv);
    return pvsg_launch_status();
}

extern "C" int pvsg_maxpool3x3s2_nhwc_backward(const float* x, const float* dy, float* dx, int B, int H, int W, int C, void* stream) {
    PVSG_CHECK_ARG(x && dy && dx && B > 0 && H > 0 && W > 0 && C > 0);
    cudaStream_t st = as_stream(stream);
    if (cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)B * H * W * C, st) != cudaSuccess) return PVSG_ERR_LAUNCH;
    const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
    const int64_t total = (int64_t)B * OH * OW * C;
    maxpool_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, dy, dx, H, W, OH, OW, C, total);
    return pvsg_launch_status();
}

extern "C" int pvsg_transpose_split(const float* x, const float* add, void* hi, void* lo, int64_t T, int C, int OH, int OW,
                                    int64_t sb, int64_t sh, int64_t sw, int64_t ldt, void* stream) {
    PVSG_CHECK_ARG(x && hi && lo && T > 0 && C > 0 && OH > 0 && OW > 0 && ldt >= T && ldt % 2 == 0);
    PVSG_CHECK_ARG(((reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 3) == 0);
    dim3 grid((unsigned)((T + 63) / 64), (unsigned)((C + 31) / 32));
    if (grid.y > 65535) return PVSG_ERR_UNSUPPORTED;
    transpose_split_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, add, reinterpret_cast<__nv_bfloat16*>(hi),
                                                                        reinterpret_cast<__nv_bfloat16*>(lo), T, C, OH, OW, sb, sh, sw, ldt);
    return pvsg_launch_status();
}
