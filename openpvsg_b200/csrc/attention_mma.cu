// Masked attention on the tensor cores, head dim 32 or 128, fp32-grade ("split-bf16").
//
//   out = softmax(scale * Q K^T + mask) V        per (batch, head)
//
// K and V arrive as the (hi, lo) bf16 operand planes their projection GEMM emitted (same bytes as
// fp32, no conversion pass); Q (<= a few hundred rows) is scaled and split in registers.  Every
// product is three warp-level MMAs (lo.hi + hi.lo + hi.hi, m16n8k16, fp32 accumulate), so S and
// O carry fp32-level error like the rest of the engine.  The thread-per-query fp32 kernel in
// attention.cu spent its time re-reading K / V rows from shared memory (one LDS.128 per four FMAs,
// 81 % LSU-pipe busy, 11-15 TFLOP/s); here a warp owns 16 queries, a CTA (<= 8 warps) shares
// 64-key tiles staged with cp.async (double buffered), fragments come from ldmatrix, and the
// online softmax works on the accumulator fragments (FlashAttention-2 register layout: the S
// fragment of QK^T is the A fragment of PV).
//
// mask: uint8 [B, Lq, Lk], non-zero = blocked; rows whose row_open count is 0 ignore the mask
// (mask2former_head.py:453-454).  Keys are split over CTAs; partial (O, m, l) are merged by
// attn_combine_kernel's log-sum-exp rule (same workspace layout as pvsg_attention).
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int MAXWARP = 8;    // 128 queries per CTA
// head dim D in {32, 128}; KT keys per tile (64 / 32); shared-memory rows of D + 8 bf16 (pitch = 16 B mod
// 128 B: conflict-free ldmatrix)
template <int D> struct Tile { static constexpr int KT = D == 32 ? 64 : 32; static constexpr int PITCH = D + 8; };

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat16 ha = __float2bfloat16_rn(a), hb = __float2bfloat16_rn(b);
    const __nv_bfloat16 la = __float2bfloat16_rn(a - __bfloat162float(ha));
    const __nv_bfloat16 lb = __float2bfloat16_rn(b - __bfloat162float(hb));
    hi = (uint32_t)__bfloat16_as_ushort(ha) | ((uint32_t)__bfloat16_as_ushort(hb) << 16);
    lo = (uint32_t)__bfloat16_as_ushort(la) | ((uint32_t)__bfloat16_as_ushort(lb) << 16);
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
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int n = valid ? 16 : 0;   // src-size 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}

struct AttnArgs {
    const float* Q;
    const __nv_bfloat16 *K_hi, *K_lo, *V_hi, *V_lo;
    const uint8_t* mask;
    const int32_t* row_open;
    float* out;
    float* part_o;
    float* part_ml;
    int H, Lq, Lk;
    int64_t q_bs, q_ts, k_bs, k_ts, v_bs, v_ts, o_bs, o_ts;
    float scale;
    int nsplit, keys_per_split;
};

template <int D>
__global__ void __launch_bounds__(32 * MAXWARP) attn_mma_kernel(AttnArgs a) {
    constexpr int KT = Tile<D>::KT, PITCH = Tile<D>::PITCH;
    constexpr int NT = KT / 8;      // S n-tiles (8 keys each)
    constexpr int KS = D / 16;      // k-steps of QK^T
    constexpr int OT = D / 8;       // O n-tiles (8 dims each)
    extern __shared__ __align__(16) uint8_t attn_smem[];
    // [stage][plane (Khi Klo Vhi Vlo)][key][dim]
    __nv_bfloat16 (*tile)[4][KT][PITCH] = reinterpret_cast<__nv_bfloat16 (*)[4][KT][PITCH]>(attn_smem);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int bh = blockIdx.z, b = bh / a.H, h = bh % a.H;
    const int split = blockIdx.y;
    const int q0 = blockIdx.x * (blockDim.x >> 5) * 16 + warp * 16;
    const int k_begin = split * a.keys_per_split;
    const int k_end = min(a.Lk, k_begin + a.keys_per_split);
    const int ntiles = (k_end - k_begin + KT - 1) / KT;

    // ---- Q fragments (rows q0 + g, q0 + g + 8), scaled, split ----
    const int row[2] = {q0 + g, q0 + g + 8};
    const bool rvalid[2] = {row[0] < a.Lq, row[1] < a.Lq};
    uint32_t qh[KS][4], ql[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int half = 0; half < 2; ++half)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                float2 v = make_float2(0.f, 0.f);
                if (rvalid[r])
                    v = __ldg(reinterpret_cast<const float2*>(a.Q + b * a.q_bs + (int64_t)row[r] * a.q_ts + h * D +
                                                              16 * ks + 8 * half + 2 * t));
                split_pair(v.x * a.scale, v.y * a.scale, qh[ks][2 * half + r], ql[ks][2 * half + r]);
            }
    bool use_mask[2];
    const uint8_t* mrow[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        use_mask[r] = a.mask != nullptr && rvalid[r];
        if (use_mask[r] && a.row_open) use_mask[r] = __ldg(a.row_open + (int64_t)b * a.Lq + row[r]) > 0;
        mrow[r] = a.mask ? a.mask + ((int64_t)b * a.Lq + min(row[r], a.Lq - 1)) * a.Lk : nullptr;
    }
    const bool pair_ok = (a.Lk & 1) == 0;   // 16-bit mask loads need an even row pitch

    const __nv_bfloat16* plane[4] = {a.K_hi + b * a.k_bs + h * D, a.K_lo + b * a.k_bs + h * D,
                                     a.V_hi + b * a.v_bs + h * D, a.V_lo + b * a.v_bs + h * D};
    auto issue_tile = [&](int ti, int stage) {
        const int k0 = k_begin + ti * KT;
        for (int c = threadIdx.x; c < KT * (D / 8); c += blockDim.x) {
            const int r = c / (D / 8), ch = c % (D / 8);
            const bool ok = k0 + r < k_end;
            const int64_t key = ok ? k0 + r : k_begin;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int64_t ts = p < 2 ? a.k_ts : a.v_ts;
                cp_async16(smem_addr(&tile[stage][p][r][ch * 8]), plane[p] + key * ts + ch * 8, ok);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    float o[OT][4];
#pragma unroll
    for (int i = 0; i < OT; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) o[i][e] = 0.f;
    float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};

    if (ntiles > 0) issue_tile(0, 0);
    for (int ti = 0; ti < ntiles; ++ti) {
        const int stage = ti & 1;
        const int k0 = k_begin + ti * KT;
        if (ti + 1 < ntiles) {
            issue_tile(ti + 1, stage ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        // mask bytes of this tile (independent of the shared-memory tile: issued before the barrier)
        uint32_t mk[NT][2];
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int key = k0 + 8 * j + 2 * t;
                uint32_t v = 0;
                if (use_mask[r] && key < k_end) {
                    if (pair_ok) v = __ldg(reinterpret_cast<const unsigned short*>(mrow[r] + key));
                    else v = (uint32_t)__ldg(mrow[r] + key) | (key + 1 < k_end ? (uint32_t)__ldg(mrow[r] + key + 1) << 8 : 0u);
                }
                mk[j][r] = v;
            }
        __syncthreads();

        // ---- S = Q K^T (16 x KT per warp) ----
        float s[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
#pragma unroll
            for (int e = 0; e < 4; ++e) s[j][e] = 0.f;
#pragma unroll
            for (int dq = 0; dq < D / 32; ++dq) {      // 32 dims (two k-steps) per ldmatrix.x4
                uint32_t kh[4], kl[4];
                const int r = 8 * j + (lane & 7), c = 32 * dq + (lane >> 3) * 8;
                ldsm4(kh, smem_addr(&tile[stage][0][r][c]));
                ldsm4(kl, smem_addr(&tile[stage][1][r][c]));
#pragma unroll
                for (int k2 = 0; k2 < 2; ++k2) {
                    const int ks = 2 * dq + k2;
                    mma16816(s[j], ql[ks], kh[2 * k2], kh[2 * k2 + 1]);   // small terms first
                    mma16816(s[j], qh[ks], kl[2 * k2], kl[2 * k2 + 1]);
                    mma16816(s[j], qh[ks], kh[2 * k2], kh[2 * k2 + 1]);
                }
            }
        }
        // ---- mask, online softmax ----
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int r = e >> 1;
                const int key = k0 + 8 * j + 2 * t + (e & 1);
                const bool blocked = !rvalid[r] || key >= k_end || ((mk[j][r] >> (8 * (e & 1))) & 0xffu) != 0;
                if (blocked) s[j][e] = -INFINITY;
                mx[r] = fmaxf(mx[r], s[j][e]);
            }
        float corr[2], mu[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float mn = fmaxf(m[r], mx[r]);
            corr[r] = mn == -INFINITY ? 1.f : __expf(m[r] - mn);   // exp(-inf) = 0 on the first open key
            mu[r] = mn == -INFINITY ? 0.f : mn;                      // all blocked so far: p = exp(-inf) = 0
            m[r] = mn;
            l[r] *= corr[r];
        }
#pragma unroll
        for (int i = 0; i < OT; ++i) {
            o[i][0] *= corr[0]; o[i][1] *= corr[0];
            o[i][2] *= corr[1]; o[i][3] *= corr[1];
        }
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float p = __expf(s[j][e] - mu[e >> 1]);
                s[j][e] = p;
                l[e >> 1] += p;
            }
        // ---- O += P V ----
#pragma unroll
        for (int kk = 0; kk < KT / 16; ++kk) {
            uint32_t ph[4], pl[4];
            split_pair(s[2 * kk][0], s[2 * kk][1], ph[0], pl[0]);
            split_pair(s[2 * kk][2], s[2 * kk][3], ph[1], pl[1]);
            split_pair(s[2 * kk + 1][0], s[2 * kk + 1][1], ph[2], pl[2]);
            split_pair(s[2 * kk + 1][2], s[2 * kk + 1][3], ph[3], pl[3]);
#pragma unroll
            for (int np = 0; np < D / 16; ++np) {
                uint32_t vh[4], vl[4];
                const int mi = lane >> 3;
                const int r = 16 * kk + (mi & 1) * 8 + (lane & 7), c = 16 * np + (mi >> 1) * 8;
                ldsm4_t(vh, smem_addr(&tile[stage][2][r][c]));
                ldsm4_t(vl, smem_addr(&tile[stage][3][r][c]));
                mma16816(o[2 * np], pl, vh[0], vh[1]);
                mma16816(o[2 * np], ph, vl[0], vl[1]);
                mma16816(o[2 * np], ph, vh[0], vh[1]);
                mma16816(o[2 * np + 1], pl, vh[2], vh[3]);
                mma16816(o[2 * np + 1], ph, vl[2], vl[3]);
                mma16816(o[2 * np + 1], ph, vh[2], vh[3]);
            }
        }
        __syncthreads();   // the tile is free for the copy issued in the next iteration
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
        l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        if (!rvalid[r]) continue;
        if (a.nsplit == 1) {
            const float inv = l[r] > 0.f ? 1.f / l[r] : 0.f;
            float* op = a.out + b * a.o_bs + (int64_t)row[r] * a.o_ts + h * D + 2 * t;
#pragma unroll
            for (int i = 0; i < OT; ++i)
                *reinterpret_cast<float2*>(op + 8 * i) = make_float2(o[i][2 * r] * inv, o[i][2 * r + 1] * inv);
        } else {
            const int64_t prow = ((int64_t)bh * a.Lq + row[r]) * a.nsplit + split;
            float* op = a.part_o + prow * D + 2 * t;
#pragma unroll
            for (int i = 0; i < OT; ++i) *reinterpret_cast<float2*>(op + 8 * i) = make_float2(o[i][2 * r], o[i][2 * r + 1]);
            if (t == 0) {
                a.part_ml[prow * 2] = m[r];
                a.part_ml[prow * 2 + 1] = l[r];
            }
        }
    }
}

// one thread per (bh, q, d): out = sum_s o_s * exp(m_s - M) / sum_s l_s * exp(m_s - M)
__global__ void __launch_bounds__(256) attn_mma_combine_kernel(const float* __restrict__ part_o,
                                                               const float* __restrict__ part_ml,
                                                               float* __restrict__ out, int H, int Lq, int nsplit,
                                                               int64_t o_bs, int64_t o_ts, int64_t total, int D) {
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

int pick_splits_tc(int B, int H, int qtiles, int Lk) {
    // Key splits so that the CTA count fills whole waves of the GPU (2 CTAs per SM): with 160
    // (batch, head) pairs a 2-way split gave 320 CTAs = 1.08 waves, i.e. the last 24 CTAs ran alone.
    const int64_t base = (int64_t)B * H * qtiles;
    const double slots = 2.0 * 148.0;
    int best = 1;
    double best_score = -1.0;
    for (int ns = 1; ns <= 32; ns *= 2) {
        if (ns > 1 && Lk / ns < 2 * 64) break;               // at least two 64-key tiles per split
        const double waves = (double)(base * ns) / slots;
        const double eff = waves / (double)(int64_t)(waves + 0.999999);
        double score = eff - 0.015 * (ns > 1 ? __builtin_ctz((unsigned)ns) : 0);  // partials + combine cost
        if (score > best_score + 1e-9) { best_score = score; best = ns; }
    }
    return best;
}

template <int D>
constexpr size_t attn_smem() { return 2ull * 4 * Tile<D>::KT * Tile<D>::PITCH * sizeof(__nv_bfloat16); }

template <int D>
int configure_attn() {
    static bool configured[PVSG_MAX_DEVICES];
    if (pvsg_first_use_on_device(configured) &&
        cudaFuncSetAttribute(attn_mma_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_smem<D>()) != cudaSuccess)
        return PVSG_ERR_LAUNCH;
    return PVSG_OK;
}

template <int D>
int launch_attn(const AttnArgs& a, dim3 grid, int nwarp, cudaStream_t st) {
    if (const int rc = configure_attn<D>()) return rc;
    attn_mma_kernel<D><<<grid, 32 * nwarp, attn_smem<D>(), st>>>(a);
    return PVSG_OK;
}

}  // namespace

int pvsg_internal::configure_attention_mma() {
    const int rc = configure_attn<32>();
    return rc ? rc : configure_attn<128>();
}

extern "C" int64_t pvsg_attention_tc_workspace_bytes(int B, int H, int Lq, int Lk, int D) {
    const int qtiles = (Lq + 16 * MAXWARP - 1) / (16 * MAXWARP);
    const int ns = pick_splits_tc(B, H, qtiles, Lk);
    if (ns == 1) return 16;
    return (int64_t)B * H * Lq * ns * (D + 2) * sizeof(float) + 16;
}

extern "C" int pvsg_attention_tc(const float* Q, const void* K_hi, const void* K_lo, const void* V_hi,
                                 const void* V_lo, const uint8_t* mask, const int32_t* row_open, float* out,
                                 void* ws, int B, int H, int Lq, int Lk, int Dh, int64_t q_bs, int64_t q_ts,
                                 int64_t k_bs, int64_t k_ts, int64_t v_bs, int64_t v_ts, int64_t o_bs,
                                 int64_t o_ts, float scale, void* stream) {
    PVSG_CHECK_ARG(Q && K_hi && K_lo && V_hi && V_lo && out && B > 0 && H > 0 && Lq > 0 && Lk > 0);
    if (Dh != 32 && Dh != 128) return PVSG_ERR_UNSUPPORTED;
    const int D = Dh;
    // 16-byte cp.async chunks of the planes, 8-byte loads / stores of Q and out
    PVSG_CHECK_ARG((k_bs | k_ts | v_bs | v_ts) % 8 == 0 && (q_bs | q_ts | o_bs | o_ts) % 2 == 0);
    PVSG_CHECK_ARG(((reinterpret_cast<uintptr_t>(K_hi) | reinterpret_cast<uintptr_t>(K_lo) |
                     reinterpret_cast<uintptr_t>(V_hi) | reinterpret_cast<uintptr_t>(V_lo)) & 15) == 0);
    PVSG_CHECK_ARG(((reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(out)) & 7) == 0);
    const int nwarp = Lq >= 16 * MAXWARP ? MAXWARP : (Lq + 15) / 16;
    const int qtiles = (Lq + 16 * nwarp - 1) / (16 * nwarp);
    const int ns = pick_splits_tc(B, H, (Lq + 16 * MAXWARP - 1) / (16 * MAXWARP), Lk);
    PVSG_CHECK_ARG(ns == 1 || ws);
    int kps = (Lk + ns - 1) / ns;
    kps = (kps + 63) / 64 * 64;      // multiple of both tile sizes
    AttnArgs a;
    a.Q = Q;
    a.K_hi = reinterpret_cast<const __nv_bfloat16*>(K_hi); a.K_lo = reinterpret_cast<const __nv_bfloat16*>(K_lo);
    a.V_hi = reinterpret_cast<const __nv_bfloat16*>(V_hi); a.V_lo = reinterpret_cast<const __nv_bfloat16*>(V_lo);
    a.mask = mask; a.row_open = row_open; a.out = out;
    a.part_o = reinterpret_cast<float*>(ws);
    a.part_ml = a.part_o ? a.part_o + (int64_t)B * H * Lq * ns * D : nullptr;
    a.H = H; a.Lq = Lq; a.Lk = Lk;
    a.q_bs = q_bs; a.q_ts = q_ts; a.k_bs = k_bs; a.k_ts = k_ts; a.v_bs = v_bs; a.v_ts = v_ts; a.o_bs = o_bs; a.o_ts = o_ts;
    a.scale = scale; a.nsplit = ns; a.keys_per_split = kps;
    dim3 grid((unsigned)qtiles, (unsigned)ns, (unsigned)(B * H));
    PVSG_CHECK_ARG(grid.z <= 65535);
    cudaStream_t st = as_stream(stream);
    const int rc = D == 32 ? launch_attn<32>(a, grid, nwarp, st) : launch_attn<128>(a, grid, nwarp, st);
    if (rc != PVSG_OK) return rc;
    if (ns > 1) {
        const int64_t total = (int64_t)B * H * Lq * D;
        attn_mma_combine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a.part_o, a.part_ml, out, H, Lq, ns,
                                                                                o_bs, o_ts, total, D);
    }
    return pvsg_launch_status();
}
