// Relation-head glue kernels (reference models/relation_head/base.py:49-62,
// test_utils.py:4-22, train_utils.py:67-81, transformer.py:77-81).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) max_over_time_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                            int T, int C, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // (n, c)
    if (i >= total) return;
    const int64_t n = i / C;
    const int c = (int)(i % C);
    const float* p = x + n * T * C + c;
    float m = -INFINITY;
    for (int t = 0; t < T; ++t) m = fmaxf(m, __ldg(p + (int64_t)t * C));
    y[i] = m;
}

// One warp per (i, j): pair = b2 + sum_h w2[h] * relu(U[i,h] + V[j,h]); diagonal = 0.
__global__ void __launch_bounds__(256) pair_kernel(const float* __restrict__ U, const float* __restrict__ V,
                                                   const float* __restrict__ w2, const float* __restrict__ b2,
                                                   float* __restrict__ pair, int N, int Hd) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= (int64_t)N * N) return;
    const int i = (int)(wid / N), j = (int)(wid % N);
    if (i == j) {
        if (lane == 0) pair[wid] = 0.f;
        return;
    }
    const float4* u = reinterpret_cast<const float4*>(U + (int64_t)i * Hd);
    const float4* v = reinterpret_cast<const float4*>(V + (int64_t)j * Hd);
    const float4* w = reinterpret_cast<const float4*>(w2);
    float s = 0.f;
    for (int h = lane; h < Hd / 4; h += 32) {
        const float4 a = __ldg(u + h), b = __ldg(v + h), c = __ldg(w + h);
        s = fmaf(c.x, fmaxf(a.x + b.x, 0.f), s);
        s = fmaf(c.y, fmaxf(a.y + b.y, 0.f), s);
        s = fmaf(c.z, fmaxf(a.z + b.z, 0.f), s);
        s = fmaf(c.w, fmaxf(a.w + b.w, 0.f), s);
    }
    s = warp_sum(s);
    if (lane == 0) pair[wid] = s + __ldg(b2);
}

// Single CTA: iterative selection of the k largest entries (diagonal = -inf) in descending
// order, ties -> lower flat index.
constexpr int TP_THREADS = 1024;
constexpr int TP_PER = 256;  // up to 1024*256 = 262144 entries
__global__ void __launch_bounds__(TP_THREADS) top_pairs_kernel(const float* __restrict__ pair, int N, int k,
                                                               int32_t* __restrict__ pairs, int32_t* __restrict__ n_out) {
    __shared__ float sv[32];
    __shared__ int si[32];
    __shared__ int chosen;
    const int total = N * N;
    const int kk = min(k, total);
    int count = 0;
    // Bounded selection: at step `it` pick the largest (value, -index) strictly below the
    // previously chosen one in the (value desc, index asc) order.  No scratch memory needed.
    float pv = INFINITY;
    int pi = -1;
    for (int it = 0; it < kk; ++it) {
        float bv = -INFINITY;
        int bi = INT_MAX;
        bool have = false;
        for (int e = threadIdx.x; e < total; e += TP_THREADS) {
            const int r = e / N, c = e % N;
            const float v = (r == c) ? -INFINITY : __ldg(pair + e);
            // strictly after (pv, pi) in the order: v < pv, or v == pv and e > pi
            const bool after = (v < pv) || (v == pv && e > pi);
            if (!after) continue;
            if (!have || v > bv || (v == bv && e < bi)) { bv = v; bi = e; have = true; }
        }
        // warp reduce (value desc, index asc); lanes without a candidate carry bi = INT_MAX
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi != INT_MAX && (bi == INT_MAX || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
        }
        if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
        __syncthreads();
        if (threadIdx.x < 32) {
            bv = sv[threadIdx.x];
            bi = si[threadIdx.x];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (oi != INT_MAX && (bi == INT_MAX || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
            }
            if (threadIdx.x == 0) {
                chosen = bi;
                sv[0] = bv;
            }
        }
        __syncthreads();
        const int ci = chosen;
        const float cv = sv[0];
        __syncthreads();
        if (ci == INT_MAX) break;
        pv = cv;
        pi = ci;
        const int r = ci / N, c = ci % N;
        if (r != c) {  // diagonal entries (value -inf) are dropped like the reference's filter
            if (threadIdx.x == 0) {
                pairs[2 * count] = r;
                pairs[2 * count + 1] = c;
            }
            ++count;
        }
    }
    if (threadIdx.x == 0) *n_out = count;
}

__global__ void __launch_bounds__(256) gather_pairs_kernel(const float* __restrict__ sub, const float* __restrict__ obj,
                                                           const int32_t* __restrict__ pairs,
                                                           const float* __restrict__ pe, float* __restrict__ out,
                                                           int T, int fq, int64_t total) {
    // out [P, T, 2F] in float4 units: fq = F/4
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % (2 * fq));
        int64_t t = i / (2 * fq);
        const int tt = (int)(t % T);
        const int p = (int)(t / T);
        const int s = pairs[2 * p], o = pairs[2 * p + 1];
        float4 v = c < fq ? __ldg(reinterpret_cast<const float4*>(sub) + ((int64_t)s * T + tt) * fq + c)
                          : __ldg(reinterpret_cast<const float4*>(obj) + ((int64_t)o * T + tt) * fq + (c - fq));
        if (pe) {
            const float4 e = __ldg(reinterpret_cast<const float4*>(pe) + (int64_t)tt * 2 * fq + c);
            v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
        }
        reinterpret_cast<float4*>(out)[i] = v;
    }
}

}  // namespace

extern "C" int pvsg_max_over_time(const float* x, float* y, int N, int T, int C, void* stream) {
    PVSG_CHECK_ARG(x && y && N > 0 && T > 0 && C > 0);
    const int64_t total = (int64_t)N * C;
    max_over_time_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(x, y, T, C, total);
    return pvsg_launch_status();
}

extern "C" int pvsg_pair_proposal(const float* U, const float* V, const float* w2, const float* b2,
                                  float* pair, int N, int Hd, void* stream) {
    PVSG_CHECK_ARG(U && V && w2 && b2 && pair && N > 0 && Hd > 0 && Hd % 4 == 0);
    const int64_t warps = (int64_t)N * N;
    pair_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, as_stream(stream)>>>(U, V, w2, b2, pair, N, Hd);
    return pvsg_launch_status();
}

extern "C" int pvsg_top_pairs(const float* pair, int N, int k, int32_t* pairs, int32_t* n_out, void* stream) {
    PVSG_CHECK_ARG(pair && pairs && n_out && N > 0 && k > 0);
    if ((int64_t)N * N > (int64_t)TP_THREADS * TP_PER) return PVSG_ERR_UNSUPPORTED;
    top_pairs_kernel<<<1, TP_THREADS, 0, as_stream(stream)>>>(pair, N, k, pairs, n_out);
    return pvsg_launch_status();
}

extern "C" int pvsg_gather_pairs(const float* sub, const float* obj, const int32_t* pairs, const float* pe,
                                 float* out, int P, int T, int F, void* stream) {
    PVSG_CHECK_ARG(sub && obj && pairs && out && P > 0 && T > 0 && F > 0 && F % 4 == 0);
    const int64_t total = (int64_t)P * T * (F / 2);
    gather_pairs_kernel<<<(unsigned)imin64((total + 255) / 256, 148 * 16), 256, 0, as_stream(stream)>>>(
        sub, obj, pairs, pe, out, T, F / 4, total);
    return pvsg_launch_status();
}
