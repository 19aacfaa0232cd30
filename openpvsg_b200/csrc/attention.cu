// Multi-head attention core  out = softmax(scale * Q K^T + mask) V  (fp32, flash-style).
//
// A CTA streams a range of keys (one "split") through shared memory for one (batch, head);
// each query is owned by TPQ threads holding D/TPQ = 32 dims of q and of the output
// accumulator in registers, with an online softmax.  Splits are merged by a second tiny
// kernel (log-sum-exp combine).  Blocked keys (mask != 0) are skipped; rows whose row_open
// count is 0 ignore the mask (the reference resets fully-masked rows to all-False,
// mask2former_head.py:453-454).
#include "common.cuh"

namespace {

constexpr int KC = 32;        // keys staged per iteration
constexpr int THREADS = 128;

template <int D, int TPQ>
__global__ void __launch_bounds__(THREADS) attn_kernel(
    const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V,
    const uint8_t* __restrict__ mask, const int32_t* __restrict__ row_open, float* __restrict__ out,
    float* __restrict__ part_o, float* __restrict__ part_ml, int H, int Lq, int Lk, int64_t q_bs, int64_t q_ts,
    int64_t k_bs, int64_t k_ts, int64_t v_bs, int64_t v_ts, int64_t o_bs, int64_t o_ts, float scale,
    int nsplit, int keys_per_split) {
    constexpr int DPT = D / TPQ;  // 32
    constexpr int QPB = THREADS / TPQ;
    static_assert(DPT == 16, "each thread owns 16 dims");
    __shared__ __align__(16) float Ks[KC][D];
    __shared__ __align__(16) float Vs[KC][D];

    const int bh = blockIdx.z;
    const int b = bh / H, h = bh % H;
    const int split = blockIdx.y;
    const int qi = blockIdx.x * QPB + threadIdx.x / TPQ;
    const int sub = threadIdx.x % TPQ;  // which 32-dim slice
    const bool qvalid = qi < Lq;

    float q[DPT], o[DPT];
#pragma unroll
    for (int d = 0; d < DPT; ++d) o[d] = 0.f;
    if (qvalid) {
        const float4* qp = reinterpret_cast<const float4*>(Q + b * q_bs + (int64_t)qi * q_ts + h * D + sub * DPT);
#pragma unroll
        for (int d = 0; d < DPT / 4; ++d) {
            float4 v = __ldg(qp + d);
            q[4 * d] = v.x * scale; q[4 * d + 1] = v.y * scale; q[4 * d + 2] = v.z * scale; q[4 * d + 3] = v.w * scale;
        }
    } else {
#pragma unroll
        for (int d = 0; d < DPT; ++d) q[d] = 0.f;
    }
    bool use_mask = mask != nullptr;
    if (use_mask && row_open && qvalid) use_mask = __ldg(row_open + (int64_t)b * Lq + qi) > 0;
    const uint8_t* mrow = (mask && qvalid) ? mask + ((int64_t)b * Lq + qi) * Lk : nullptr;

    float m = -INFINITY, l = 0.f;
    const int k_begin = split * keys_per_split;
    const int k_end = min(Lk, k_begin + keys_per_split);
    for (int k0 = k_begin; k0 < k_end; k0 += KC) {
        const int nk = min(KC, k_end - k0);
        __syncthreads();
        // cooperative, coalesced staging of K and V rows [k0, k0+nk) x head slice
        constexpr int F4 = D / 4;
        for (int i = threadIdx.x; i < KC * F4; i += THREADS) {
            const int r = i / F4, c = i % F4;
            float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
            if (r < nk) {
                kv = __ldg(reinterpret_cast<const float4*>(K + b * k_bs + (int64_t)(k0 + r) * k_ts + h * D) + c);
                vv = __ldg(reinterpret_cast<const float4*>(V + b * v_bs + (int64_t)(k0 + r) * v_ts + h * D) + c);
            }
            reinterpret_cast<float4*>(&Ks[r][0])[c] = kv;
            reinterpret_cast<float4*>(&Vs[r][0])[c] = vv;
        }
        __syncthreads();
        // keys in groups of 4: four independent dot-product chains, one softmax update per group
        // (rows >= nk of the staged chunk are zero and are treated as blocked)
        for (int j = 0; j < nk; j += 4) {
            float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int d = 0; d < DPT / 4; ++d) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 kk = reinterpret_cast<const float4*>(&Ks[j + e][sub * DPT])[d];
                    s[e] = fmaf(q[4 * d], kk.x, s[e]);
                    s[e] = fmaf(q[4 * d + 1], kk.y, s[e]);
                    s[e] = fmaf(q[4 * d + 2], kk.z, s[e]);
                    s[e] = fmaf(q[4 * d + 3], kk.w, s[e]);
                }
            }
            if (TPQ > 1) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
#pragma unroll
                    for (int off = 1; off < TPQ; off <<= 1) s[e] += __shfl_xor_sync(0xffffffffu, s[e], off);
            }
            float mx = -INFINITY;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const bool blocked = !qvalid || (j + e >= nk) || (use_mask && mrow && mrow[k0 + j + e] != 0);
                if (blocked) s[e] = -INFINITY;
                mx = fmaxf(mx, s[e]);
            }
            if (mx > -INFINITY) {  // else the whole group is blocked (uniform across the TPQ lanes of a query)
                if (mx > m) {
                    const float c = __expf(m - mx);  // exp(-inf) = 0 on the first open key
                    l *= c;
#pragma unroll
                    for (int d = 0; d < DPT; ++d) o[d] *= c;
                    m = mx;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float p = __expf(s[e] - m);  // 0 for blocked keys
                    l += p;
                    const float4* vr = reinterpret_cast<const float4*>(&Vs[j + e][sub * DPT]);
#pragma unroll
                    for (int d = 0; d < DPT / 4; ++d) {
                        const float4 vv = vr[d];
                        o[4 * d] = fmaf(p, vv.x, o[4 * d]);
                        o[4 * d + 1] = fmaf(p, vv.y, o[4 * d + 1]);
                        o[4 * d + 2] = fmaf(p, vv.z, o[4 * d + 2]);
                        o[4 * d + 3] = fmaf(p, vv.w, o[4 * d + 3]);
                    }
                }
            }
        }
    }
    if (!qvalid) return;
    if (nsplit == 1) {
        const float inv = l > 0.f ? 1.f / l : 0.f;
        float4* op = reinterpret_cast<float4*>(out + b * o_bs + (int64_t)qi * o_ts + h * D + sub * DPT);
#pragma unroll
        for (int d = 0; d < DPT / 4; ++d)
            op[d] = make_float4(o[4 * d] * inv, o[4 * d + 1] * inv, o[4 * d + 2] * inv, o[4 * d + 3] * inv);
    } else {
        const int64_t row = ((int64_t)bh * Lq + qi) * nsplit + split;
        float4* op = reinterpret_cast<float4*>(part_o + row * D + sub * DPT);
#pragma unroll
        for (int d = 0; d < DPT / 4; ++d) op[d] = make_float4(o[4 * d], o[4 * d + 1], o[4 * d + 2], o[4 * d + 3]);
        if (sub == 0) {
            part_ml[row * 2] = m;
            part_ml[row * 2 + 1] = l;
        }
    }
}

// one thread per (bh, q, d): out = sum_s o_s * exp(m_s - M) / sum_s l_s * exp(m_s - M)
__global__ void __launch_bounds__(256) attn_combine_kernel(const float* __restrict__ part_o,
                                                           const float* __restrict__ part_ml,
                                                           float* __restrict__ out, int H, int Lq, int D, int nsplit,
                                                           int64_t o_bs, int64_t o_ts, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int d = (int)(i % D);
    const int64_t row = i / D;  // (bh * Lq + q)
    const int qi = (int)(row % Lq);
    const int bh = (int)(row / Lq);
    const int b = bh / H, h = bh % H;
    float M = -INFINITY;
    for (int s = 0; s < nsplit; ++s) M = fmaxf(M, part_ml[(row * nsplit + s) * 2]);
    float num = 0.f, den = 0.f;
    for (int s = 0; s < nsplit; ++s) {
        const float ms = part_ml[(row * nsplit + s) * 2];
        if (ms == -INFINITY) continue;
        const float c = __expf(ms - M);
        num = fmaf(part_o[(row * nsplit + s) * D + d], c, num);
        den = fmaf(part_ml[(row * nsplit + s) * 2 + 1], c, den);
    }
    out[b * o_bs + (int64_t)qi * o_ts + h * D + d] = den > 0.f ? num / den : 0.f;
}

int pick_splits(int B, int H, int Lq, int Lk, int qpb) {
    const int qtiles = (Lq + qpb - 1) / qpb;
    const int64_t base = (int64_t)B * H * qtiles;
    int ns = 1;
    // aim for ~4 CTAs per SM while keeping >= 2 key chunks per split
    while (base * ns < 592 && Lk / (ns * 2) >= 2 * KC && ns < 128) ns *= 2;
    return ns;
}

}  // namespace

extern "C" int64_t pvsg_attention_workspace_bytes(int B, int H, int Lq, int Lk, int D) {
    const int qpb = D == 32 ? THREADS / 2 : THREADS / 8;
    const int ns = pick_splits(B, H, Lq, Lk, qpb);
    if (ns == 1) return 16;
    return (int64_t)B * H * Lq * ns * (D + 2) * sizeof(float) + 16;
}

extern "C" int pvsg_attention(const float* Q, const float* K, const float* V, const uint8_t* mask,
                              const int32_t* row_open, float* out, void* ws, int B, int H, int Lq, int Lk,
                              int D, int64_t q_bs, int64_t q_ts, int64_t k_bs, int64_t k_ts, int64_t v_bs,
                              int64_t v_ts, int64_t o_bs, int64_t o_ts, float scale, void* stream) {
    PVSG_CHECK_ARG(Q && K && V && out && B > 0 && H > 0 && Lq > 0 && Lk > 0);
    PVSG_CHECK_ARG((q_bs | q_ts | k_bs | k_ts | v_bs | v_ts | o_bs | o_ts) % 4 == 0);
    if (D != 32 && D != 128) return PVSG_ERR_UNSUPPORTED;
    const int qpb = D == 32 ? THREADS / 2 : THREADS / 8;
    const int ns = pick_splits(B, H, Lq, Lk, qpb);
    PVSG_CHECK_ARG(ns == 1 || ws);
    int kps = (Lk + ns - 1) / ns;
    kps = (kps + KC - 1) / KC * KC;
    float* part_o = reinterpret_cast<float*>(ws);
    float* part_ml = part_o ? part_o + (int64_t)B * H * Lq * ns * D : nullptr;
    dim3 grid((unsigned)((Lq + qpb - 1) / qpb), (unsigned)ns, (unsigned)(B * H));
    PVSG_CHECK_ARG(grid.z <= 65535);
    cudaStream_t st = as_stream(stream);
    if (D == 32)
        attn_kernel<32, 2><<<grid, THREADS, 0, st>>>(Q, K, V, mask, row_open, out, part_o, part_ml, H, Lq, Lk, q_bs,
                                                      q_ts, k_bs, k_ts, v_bs, v_ts, o_bs, o_ts, scale, ns, kps);
    else
        attn_kernel<128, 8><<<grid, THREADS, 0, st>>>(Q, K, V, mask, row_open, out, part_o, part_ml, H, Lq, Lk, q_bs,
                                                       q_ts, k_bs, k_ts, v_bs, v_ts, o_bs, o_ts, scale, ns, kps);
    if (ns > 1) {
        const int64_t total = (int64_t)B * H * Lq * D;
        attn_combine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(part_o, part_ml, out, H, Lq, D, ns,
                                                                            o_bs, o_ts, total);
    }
    return pvsg_launch_status();
}
