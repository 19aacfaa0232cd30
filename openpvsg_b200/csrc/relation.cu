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

// Temporal taps of a [P, T, C] tube-pair tensor (zero padded, cross-correlation as F.conv1d):
//   FIR    : y[p,t,c]       = sum_k w[k] * x[p, t + k - K/2, c]          (HandcraftedFilter, convolution.py:26-30)
//   UNFOLD : y[p,t,k*C + c] =              x[p, t + k - K/2, c]          (operand of Learnable1DConv's Conv1d as one GEMM)
template <bool FIR>
__global__ void __launch_bounds__(256) temporal_taps_kernel(const float4* __restrict__ x, const float* __restrict__ w,
                                                            float4* __restrict__ y, int T, int C4, int K, int64_t total) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4);
        const int64_t pt = i / C4;
        const int t = (int)(pt % T);
        const int64_t p = pt / T;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < K; ++k) {
            const int ts = t + k - K / 2;
            const float4 v = (ts >= 0 && ts < T) ? __ldg(x + (p * T + ts) * C4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (FIR) {
                const float wk = __ldg(w + k);
                acc.x = fmaf(wk, v.x, acc.x); acc.y = fmaf(wk, v.y, acc.y);
                acc.z = fmaf(wk, v.z, acc.z); acc.w = fmaf(wk, v.w, acc.w);
            } else {
                y[(pt * K + k) * C4 + c] = v;
            }
        }
        if (FIR) y[i] = acc;
    }
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

// Single CTA: the k largest entries (diagonal = -inf) in descending order, ties -> lower flat index.
// Radix select on order-preserving integer keys (4 passes of 8 bits) finds the k-th largest value,
// the <= k survivors are compacted in index order into shared memory and ranked by counting.
// (The first version re-scanned the whole matrix once per selected pair: 937 us for 200 x 200, k = 100.)
constexpr int TP_THREADS = 1024;
constexpr int TP_MAXK = 1024;
constexpr int TP_CACHE_BYTES = 200 * 1024;     // key cache in dynamic shared memory (N <= 226)
__device__ __forceinline__ unsigned tp_key(const float* __restrict__ pair, int e, int N) {
    const int r = e / N, c = e - r * N;
    const unsigned u = __float_as_uint((r == c) ? -INFINITY : __ldg(pair + e));
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);     // larger float <=> larger key
}
// cache_elems >= N * N: the order-preserving keys are computed ONCE into dynamic shared memory (160 KB for 200 x 200) and
// the four radix passes + the survivor pass read them from there; before, every pass re-read the matrix from L2 with one
// exposed load latency per 1024 elements (121 us for 200 x 200; now the first pass is the only one that touches memory).
__global__ void __launch_bounds__(TP_THREADS) top_pairs_kernel(const float* __restrict__ pair, int N, int k,
                                                               int32_t* __restrict__ pairs, int32_t* __restrict__ n_out,
                                                               int cache_elems) {
    extern __shared__ unsigned tp_cache[];
    __shared__ unsigned hist[256];
    __shared__ unsigned sel_prefix, sel_remaining, out_gt, out_eq;
    __shared__ unsigned wc_gt[32], wc_eq[32];
    __shared__ unsigned ckey[TP_MAXK];
    __shared__ int cidx[TP_MAXK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int total = N * N;
    const int kk = min(k, total);
    const bool cached = cache_elems >= total;
    if (threadIdx.x == 0) { sel_prefix = 0; sel_remaining = (unsigned)kk; out_gt = 0; out_eq = 0; }
    if (cached) {
        int e = threadIdx.x;
        for (; e + 3 * TP_THREADS < total; e += 4 * TP_THREADS) {      // four independent loads in flight per thread
            const unsigned a = tp_key(pair, e, N), b = tp_key(pair, e + TP_THREADS, N), c = tp_key(pair, e + 2 * TP_THREADS, N),
                           d = tp_key(pair, e + 3 * TP_THREADS, N);
            tp_cache[e] = a; tp_cache[e + TP_THREADS] = b; tp_cache[e + 2 * TP_THREADS] = c; tp_cache[e + 3 * TP_THREADS] = d;
        }
        for (; e < total; e += TP_THREADS) tp_cache[e] = tp_key(pair, e, N);
    }
    auto key = [&](int e) { return cached ? tp_cache[e] : tp_key(pair, e, N); };
    __syncthreads();
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        const unsigned prefix = sel_prefix;
        const unsigned pmask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        for (int base = 0; base < total; base += blockDim.x) {
            const int e = base + threadIdx.x;
            const unsigned u = e < total ? key(e) : 0u;
            const bool in = e < total && (u & pmask) == prefix;
            const unsigned bin = in ? ((u >> shift) & 255u) : 256u;
            const unsigned peers = __match_any_sync(0xffffffffu, bin);
            if (in && lane == (__ffs(peers) - 1)) atomicAdd(&hist[bin], (unsigned)__popc(peers));
        }
        __syncthreads();
        if (warp == 0) {
            const unsigned rem0 = sel_remaining;
            unsigned mine[8], tot = 0;
#pragma unroll
            for (int e = 0; e < 8; ++e) { mine[e] = hist[255 - 8 * lane - e]; tot += mine[e]; }
            unsigned incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const unsigned excl = incl - tot;
            if (excl < rem0 && rem0 <= incl) {
                unsigned rem = rem0 - excl;
                int b = 255 - 8 * lane;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    if (mine[e] >= rem) { b = 255 - 8 * lane - e; break; }
                    rem -= mine[e];
                }
                sel_prefix = prefix | ((unsigned)b << shift);
                sel_remaining = rem;
            }
        }
        __syncthreads();
    }
    const unsigned thr = sel_prefix;
    const unsigned n_eq = sel_remaining;
    // survivors -> shared memory, in flat-index order
    for (int base = 0; base < total; base += blockDim.x) {
        const int e = base + threadIdx.x;
        const unsigned u = e < total ? key(e) : 0u;
        const bool gt = e < total && u > thr, eq = e < total && u == thr;
        const unsigned bg = __ballot_sync(0xffffffffu, gt), be = __ballot_sync(0xffffffffu, eq);
        if (lane == 0) { wc_gt[warp] = __popc(bg); wc_eq[warp] = __popc(be); }
        __syncthreads();
        unsigned og = out_gt, oe = out_eq;
        for (int w = 0; w < warp; ++w) { og += wc_gt[w]; oe += wc_eq[w]; }
        const unsigned lower = (1u << lane) - 1u;
        if (gt) {
            const unsigned slot = og + __popc(bg & lower);
            ckey[slot] = u; cidx[slot] = e;
        } else if (eq) {
            const unsigned r = oe + __popc(be & lower);
            if (r < n_eq) {
                const unsigned slot = (unsigned)kk - n_eq + r;
                ckey[slot] = u; cidx[slot] = e;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned tg = 0, te = 0;
            for (int w = 0; w < nwarp; ++w) { tg += wc_gt[w]; te += wc_eq[w]; }
            out_gt += tg; out_eq += te;
        }
        __syncthreads();
    }
    // rank by counting: (value desc, index asc); diagonal entries (-inf) rank last and are dropped
    int local_valid = 0;
    for (int i = threadIdx.x; i < kk; i += blockDim.x) {
        const unsigned ki = ckey[i];
        const int ei = cidx[i];
        int rank = 0;
        for (int j = 0; j < kk; ++j) rank += (ckey[j] > ki || (ckey[j] == ki && cidx[j] < ei)) ? 1 : 0;
        const int r = ei / N, c = ei - r * N;
        if (r != c) {
            pairs[2 * rank] = r;
            pairs[2 * rank + 1] = c;
            ++local_valid;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) { out_gt = 0; }
    __syncthreads();
    if (local_valid) atomicAdd(&out_gt, (unsigned)local_valid);
    __syncthreads();
    if (threadIdx.x == 0) n_out[0] = (int)out_gt;
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

int pvsg_internal::configure_relation() {
    static bool configured[PVSG_MAX_DEVICES];
    if (pvsg_first_use_on_device(configured) &&
        cudaFuncSetAttribute(top_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_CACHE_BYTES) != cudaSuccess)
        return PVSG_ERR_LAUNCH;
    return PVSG_OK;
}

extern "C" int pvsg_max_over_time(const float* x, float* y, int N, int T, int C, void* stream) {
    PVSG_CHECK_ARG(x && y && N > 0 && T > 0 && C > 0);
    const int64_t total = (int64_t)N * C;
    max_over_time_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(x, y, T, C, total);
    return pvsg_launch_status();
}

extern "C" int pvsg_temporal_fir(const float* x, const float* w, float* y, int P, int T, int C, int K, void* stream) {
    PVSG_CHECK_ARG(x && w && y && P > 0 && T > 0 && C > 0 && C % 4 == 0 && K > 0 && (K & 1));
    const int64_t total = (int64_t)P * T * (C / 4);
    temporal_taps_kernel<true><<<(unsigned)imin64((total + 255) / 256, 148 * 16), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(x), w, reinterpret_cast<float4*>(y), T, C / 4, K, total);
    return pvsg_launch_status();
}

extern "C" int pvsg_temporal_unfold(const float* x, float* y, int P, int T, int C, int K, void* stream) {
    PVSG_CHECK_ARG(x && y && P > 0 && T > 0 && C > 0 && C % 4 == 0 && K > 0 && (K & 1));
    const int64_t total = (int64_t)P * T * (C / 4);
    temporal_taps_kernel<false><<<(unsigned)imin64((total + 255) / 256, 148 * 16), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(x), nullptr, reinterpret_cast<float4*>(y), T, C / 4, K, total);
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
    if ((int64_t)N * N > (1LL << 30) || k > TP_MAXK) return PVSG_ERR_UNSUPPORTED;
    const int64_t total = (int64_t)N * N;
    const bool cache = total * 4 <= TP_CACHE_BYTES;
    if (cache && pvsg_internal::configure_relation() != PVSG_OK) return PVSG_ERR_LAUNCH;
    top_pairs_kernel<<<1, TP_THREADS, cache ? (size_t)total * 4 : 0, as_stream(stream)>>>(pair, N, k, pairs, n_out,
                                                                                          cache ? (int)total : 0);
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
