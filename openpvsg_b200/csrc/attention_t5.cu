// Masked attention on the 5th-generation tensor cores (tcgen05 + TMEM), head dim 32 or 128, fp32-grade ("split-bf16").
//
//   out = softmax(scale * Q K^T + mask) V        per (batch, head)
//
// Relation head (ObjectEncoder: 200 x 200 x 32 per (frame, head), models/relation_head/base.py:26-40; TemporalTransformer:
// 128 x 128 x 128, transformer.py:35-56) and the decoder's masked cross-attention (100 x 14 720 x 32,
// models/mask2former/mask2former_head.py:457-468).  One CTA owns a 128-query tile of one (batch, head) and walks its share
// of the keys in 128-key tiles:
//
//   control warp (one lane) : TMA (cp.async.bulk.tensor, 64B / 128B swizzle) of the K and V operand planes (hi, lo);
//                             S = Q K^T as tcgen05.mma kind::f16 into TMEM (three MMAs per k-step: lo.hi + hi.lo + hi.hi),
//                             then O_j = P V with V as an MN-MAJOR B operand (no transpose anywhere);
//   8 softmax warps         : two threads per query row (= TMEM lane; warps w and w + 4 share a lane quarter and take
//                             half of the S columns / O channels each): tcgen05.ld of the S row, scale / mask, online
//                             softmax in fp32 (base 2, row maximum exchanged through shared memory), P written as
//                             split-bf16 operand planes in the canonical 128B-swizzled K-major layout (over the K tile,
//                             which is dead by then), then tcgen05.ld of O_j and the rescaled accumulation in registers.
//
// The steps of a tile are sequential inside a CTA (mbarrier hand-offs, one arrival per warp); two CTAs per SM (head dim 32:
// 98 KB of shared memory, 256 TMEM columns each) overlap each other's phases, keys are split over CTAs for long rows and
// merged by the log-sum-exp rule.  Measured per 128-key tile (clock64 in the control lane, B200): TMA 1.4-2.5 k cycles,
// QK^T 0.8 k, softmax + P 3.7 k, PV 1.9 k -- the softmax warps' instruction issue is the pacing phase.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128;
constexpr int TPR = 2;                    // softmax threads per query row: warps w and w + 4 share a TMEM lane quarter and
                                          // take half of the S columns / O channels each (a lone warp per scheduler runs at
                                          // IPC ~0.2: the first version, one thread per row, spent 12 us per key tile here)
constexpr int NSW = 4 * TPR;              // softmax warps
constexpr int NTHR = 32 * NSW + 32;       // + control warp
constexpr float kLog2e = 1.4426950408889634f;

template <int D>
struct Cfg {
    static constexpr int ROWB = D == 32 ? 64 : 128;              // bytes of one operand row inside a swizzle block
    static constexpr int KBLK = D == 32 ? 1 : 2;                 // 64-element k-blocks of Q / K (and MN blocks of V)
    static constexpr int Q_PLANE = BM * D * 2;                   // bytes of one Q plane
    static constexpr int KV_PLANE = BN * D * 2;
    static constexpr int P_PLANE = BM * BN * 2;                  // 32 KB: two 64-key k-blocks of [128 x 128 B]
    static constexpr int OFF_Q = 0;
    static constexpr int OFF_P = 2 * Q_PLANE;                    // P (hi | lo); the K tile (hi | lo) aliases its head
    static constexpr int OFF_V = OFF_P + (2 * P_PLANE > 2 * KV_PLANE ? 2 * P_PLANE : 2 * KV_PLANE);
    static constexpr int OFF_BAR = OFF_V + 2 * KV_PLANE;
    static constexpr int OFF_RED = OFF_BAR + 128;                // row-maximum exchange between the TPR threads of a row
    static constexpr int SMEM = OFF_RED + TPR * BM * 4 + 1024;   // + alignment slack
    static constexpr int TMEM_COLS = 256;                        // S: 128 columns, O: D columns
    static constexpr uint32_t SWZ = D == 32 ? 4u : 2u;           // UMMA layout type: SWIZZLE_64B / SWIZZLE_128B
    static constexpr uint32_t SBO = D == 32 ? 512u : 1024u;      // bytes between 8-row groups
};

struct T5Args {
    const float* Q;
    const uint8_t* mask;
    const int32_t* row_open;
    float* out;
    float* part_o;
    float* part_ml;
    float* lse;           // optional [B, H, Lq]: natural-log sum-exp of the scaled, masked scores (training backward)
    int H, Lq, Lk;
    int64_t q_bs, q_ts, o_bs, o_ts;
    float scale;
    int nsplit, tiles_per_split;
    int tok_dim;          // which TMA coordinate is the token axis (1 or 2), the other one is the batch
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, uint32_t dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// shared-memory operand descriptor: start address, leading / stride byte offsets, Blackwell version bit, swizzle mode
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float ex2(float x) {      // MUFU.EX2 (ex2(-inf) = 0); inputs are <= 0 here
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// (x0, x1) -> packed bf16 pairs hi = bf16(x), lo = bf16(x - hi)
__device__ __forceinline__ void split_pair_packed(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(x0, x1);
    hi = *reinterpret_cast<const uint32_t*>(&h2);
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(x0 - h0, x1 - h1);
    lo = *reinterpret_cast<const uint32_t*>(&l2);
}

// Byte offset of the 16-byte chunk (row r, chunk c of 8 bf16) inside a K-major operand block of ROWB-byte rows:
// 64B swizzle (Swizzle<2,4,3>): chunk ^= (r >> 1) & 3;  128B swizzle (Swizzle<3,4,3>): chunk ^= r & 7.
template <int ROWB>
__device__ __forceinline__ uint32_t swz_off(int r, int c) {
    return ROWB == 64 ? (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)) : (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4));
}

template <int D>
__global__ void __launch_bounds__(NTHR, D == 32 ? 2 : 1)
attn_t5_kernel(const __grid_constant__ CUtensorMap tmK_hi, const __grid_constant__ CUtensorMap tmK_lo,
               const __grid_constant__ CUtensorMap tmV_hi, const __grid_constant__ CUtensorMap tmV_lo, T5Args a) {
    using C = Cfg<D>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t q_hi = sbase + C::OFF_Q, q_lo = q_hi + C::Q_PLANE;
    const uint32_t p_hi = sbase + C::OFF_P, p_lo = p_hi + C::P_PLANE;
    const uint32_t k_hi = sbase + C::OFF_P, k_lo = k_hi + C::KV_PLANE;          // the K tile lives where P will be written
    const uint32_t v_hi = sbase + C::OFF_V, v_lo = v_hi + C::KV_PLANE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t *bar_kv = bars, *bar_s = bars + 1, *bar_p = bars + 2, *bar_o = bars + 3, *bar_free = bars + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.z, b = bh / a.H, h = bh - b * a.H;
    const int q0 = blockIdx.x * BM;
    const int split = blockIdx.y;
    const int ntiles = (a.Lk + BN - 1) / BN;
    const int t0 = split * a.tiles_per_split, t1 = min(ntiles, t0 + a.tiles_per_split);
    float* red = reinterpret_cast<float*>(smem + C::OFF_RED);          // [TPR][BM]

    if (threadIdx.x == 0) {
        mbar_init(bar_kv, 1); mbar_init(bar_s, 1); mbar_init(bar_p, NSW); mbar_init(bar_o, 1); mbar_init(bar_free, NSW);   // one arrival per softmax warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == NSW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // softmax thread -> (query row r = TMEM lane, part = which share of the columns / channels)
    const int part = warp / 4;                       // 0 .. TPR-1 (control warp: unused)
    const int r = (warp % 4) * 32 + lane;
    // ---- Q tile: fp32 rows -> split-bf16 planes in the K-major swizzled layout
    if (warp < NSW) {
        const int row = q0 + r;
        const float* qp = a.Q + (int64_t)b * a.q_bs + (int64_t)row * a.q_ts + h * D;
#pragma unroll
        for (int cc = 0; cc < D / 8 / TPR; ++cc) {
            const int c = part * (D / 8 / TPR) + cc;
            uint32_t hh[4], ll[4];
            if (row < a.Lq) {
                const float4 x = *reinterpret_cast<const float4*>(qp + 8 * c), y = *reinterpret_cast<const float4*>(qp + 8 * c + 4);
                split_pair_packed(x.x, x.y, hh[0], ll[0]); split_pair_packed(x.z, x.w, hh[1], ll[1]);
                split_pair_packed(y.x, y.y, hh[2], ll[2]); split_pair_packed(y.z, y.w, hh[3], ll[3]);
            } else {
                hh[0] = hh[1] = hh[2] = hh[3] = ll[0] = ll[1] = ll[2] = ll[3] = 0u;
            }
            const uint32_t off = (uint32_t)(c / 8) * (BM * 128) + swz_off<C::ROWB>(r, c % 8);
            sts_v4(q_hi + off, hh[0], hh[1], hh[2], hh[3]);
            sts_v4(q_lo + off, ll[0], ll[1], ll[2], ll[3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the tensor core
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + BN;

    if (warp == NSW) {
        // ------------------------------ control: TMA + MMA issue ------------------------------
        if (lane == 0) {
            // instruction descriptors: D = F32, A = B = BF16, M = 128; QK: N = 128, both K-major; PV: N = D, B MN-major
            const uint32_t idesc_qk = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
            const uint32_t idesc_pv = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(D >> 3) << 17) |
                                      ((uint32_t)(BM >> 4) << 24);
            uint32_t ph = 0;
            for (int t = t0; t < t1; ++t, ph ^= 1) {
                // K (aliasing P) and V of the previous tile are free once its PV MMAs retired (bar_o, waited below);
                // S is free since the softmax threads arrived on bar_p
                mbar_expect_tx(bar_kv, 4u * C::KV_PLANE);
                const int ck = a.tok_dim == 1 ? t * BN : b, cb = a.tok_dim == 1 ? b : t * BN;
#pragma unroll
                for (int kb = 0; kb < C::KBLK; ++kb) {
                    const int ch = h * D + kb * 64;
                    const uint32_t o = (uint32_t)kb * (BN * 128);
                    tma_load_3d(&tmK_hi, bar_kv, k_hi + o, ch, ck, cb);
                    tma_load_3d(&tmK_lo, bar_kv, k_lo + o, ch, ck, cb);
                    tma_load_3d(&tmV_hi, bar_kv, v_hi + o, ch, ck, cb);
                    tma_load_3d(&tmV_lo, bar_kv, v_lo + o, ch, ck, cb);
                }
                mbar_wait(bar_kv, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // S = Q K^T
#pragma unroll
                for (int ks = 0; ks < D / 16; ++ks) {
                    const uint32_t off = (uint32_t)(ks / 4) * (BM * 128) + (uint32_t)(ks % 4) * 32;
                    const uint64_t dqh = umma_desc(q_hi + off, 16, C::SBO, C::SWZ), dql = umma_desc(q_lo + off, 16, C::SBO, C::SWZ);
                    const uint64_t dkh = umma_desc(k_hi + off, 16, C::SBO, C::SWZ), dkl = umma_desc(k_lo + off, 16, C::SBO, C::SWZ);
                    umma_bf16(tmem_s, dql, dkh, idesc_qk, ks != 0);
                    umma_bf16(tmem_s, dqh, dkl, idesc_qk, 1);
                    umma_bf16(tmem_s, dqh, dkh, idesc_qk, 1);
                }
                umma_commit(bar_s);
                // O_j = P V   (P: K-major 128B-swizzled, two 64-key blocks; V: MN-major, 16 keys per step)
                mbar_wait(bar_p, ph);                              // P written (and S read)
                if (t > t0) mbar_wait(bar_free, ph ^ 1);           // the previous O_j has been read out of TMEM
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int ks = 0; ks < BN / 16; ++ks) {
                    const uint32_t poff = (uint32_t)(ks / 4) * (BM * 128) + (uint32_t)(ks % 4) * 32;
                    const uint32_t voff = (uint32_t)ks * 16 * C::ROWB;
                    const uint64_t dph = umma_desc(p_hi + poff, 16, 1024, 2), dpl = umma_desc(p_lo + poff, 16, 1024, 2);
                    const uint64_t dvh = umma_desc(v_hi + voff, BN * 128, C::SBO, C::SWZ), dvl = umma_desc(v_lo + voff, BN * 128, C::SBO, C::SWZ);
                    umma_bf16(tmem_o, dpl, dvh, idesc_pv, ks != 0);
                    umma_bf16(tmem_o, dph, dvl, idesc_pv, 1);
                    umma_bf16(tmem_o, dph, dvh, idesc_pv, 1);
                }
                umma_commit(bar_o);
                mbar_wait(bar_o, ph);                              // P / K / V buffers reusable
            }
        }
    } else {
        // ------------------------------ softmax / accumulation: TPR threads per query row ------------------------------
        constexpr int NC = BN / 32 / TPR;          // 32-column chunks of S per thread
        constexpr int OD = D / TPR;                // O channels per thread
        const int row = q0 + r;
        const bool rvalid = row < a.Lq;
        const uint32_t lane_addr = (uint32_t)((warp % 4) * 32) << 16;
        const bool use_mask = a.mask != nullptr && rvalid && (a.row_open == nullptr || a.row_open[(int64_t)b * a.Lq + row] > 0);
        const uint8_t* mrow = use_mask ? a.mask + ((int64_t)b * a.Lq + row) * a.Lk : nullptr;
        const bool mask4 = (a.Lk & 3) == 0;
        const float sl2 = a.scale * kLog2e;
        float m = -INFINITY, l = 0.f;
        float o[OD];
#pragma unroll
        for (int i = 0; i < OD; ++i) o[i] = 0.f;
        uint32_t ph = 0;
        for (int t = t0; t < t1; ++t, ph ^= 1) {
            const int key0 = t * BN + part * (BN / TPR);
            // mask words of this thread's columns (issued before the wait: their latency hides behind the MMA)
            uint32_t mw[NC * 8];
#pragma unroll
            for (int i = 0; i < NC * 8; ++i) {
                const int k = key0 + 4 * i;
                mw[i] = 0u;
                if (use_mask) {
                    if (mask4) mw[i] = k < a.Lk ? *reinterpret_cast<const uint32_t*>(mrow + k) : 0u;
                    else
                        for (int e = 0; e < 4; ++e) mw[i] |= (k + e < a.Lk ? (uint32_t)mrow[k + e] : 0u) << (8 * e);
                }
            }
            mbar_wait(bar_s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float sc[NC][32];
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                uint32_t v[32];
                tmem_ld32(tmem_s + lane_addr + part * (BN / TPR) + c * 32, v);
                const int rem = a.Lk - (key0 + c * 32);               // keys beyond Lk are padding
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const bool blocked = ((mw[c * 8 + e / 4] >> (8 * (e & 3))) & 0xffu) != 0u || e >= rem;
                    sc[c][e] = blocked ? -INFINITY : __uint_as_float(v[e]) * sl2;
                    mx = fmaxf(mx, sc[c][e]);
                }
            }
            // row maximum over the TPR threads of the row
            red[part * BM + r] = mx;
            asm volatile("bar.sync %0, %1;" ::"r"(1 + (warp % 4)), "r"(32 * TPR) : "memory");
#pragma unroll
            for (int pp = 0; pp < TPR; ++pp) mx = fmaxf(mx, red[pp * BM + r]);
            const float m_new = fmaxf(m, mx);
            const float alpha = m_new == -INFINITY ? 1.f : ex2(m - m_new);        // m = -inf: exp2(-inf) = 0
            const float msub = m_new == -INFINITY ? 0.f : m_new;                 // all blocked so far: every p is exp2(-inf) = 0
            float lsum = 0.f;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {
                    float p[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        p[e] = ex2(sc[c][g8 * 8 + e] - msub);
                        lsum += p[e];
                    }
                    uint32_t hh[4], ll[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) split_pair_packed(p[2 * e], p[2 * e + 1], hh[e], ll[e]);
                    const int chunk = (part * NC + c) * 4 + g8;        // 16-byte chunk of the 128-key row: 16 chunks, 8 per block
                    const uint32_t off = (uint32_t)(chunk / 8) * (BM * 128) + swz_off<128>(r, chunk % 8);
                    sts_v4(p_hi + off, hh[0], hh[1], hh[2], hh[3]);
                    sts_v4(p_lo + off, ll[0], ll[1], ll[2], ll[3]);
                }
            }
            l = l * alpha + lsum;
            m = m_new;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_p);      // 8 arrivals instead of 256 serialised shared-memory atomics
            // O += O_j  (rescaled); this thread's share of the channels
            mbar_wait(bar_o, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int c = 0; c < (OD + 31) / 32; ++c) {
                if constexpr (OD >= 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_o + lane_addr + part * OD + c * 32, v);
#pragma unroll
                    for (int e = 0; e < 32; ++e) o[c * 32 + e] = fmaf(o[c * 32 + e], alpha, __uint_as_float(v[e]));
                } else {
                    uint32_t v[16];
                    tmem_ld16(tmem_o + lane_addr + part * OD, v);
#pragma unroll
                    for (int e = 0; e < 16; ++e) o[e] = fmaf(o[e], alpha, __uint_as_float(v[e]));
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_free);
        }
        // total row sum over the TPR threads (their maxima are identical)
        asm volatile("bar.sync %0, %1;" ::"r"(1 + (warp % 4)), "r"(32 * TPR) : "memory");
        red[part * BM + r] = l;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + (warp % 4)), "r"(32 * TPR) : "memory");
        l = 0.f;
#pragma unroll
        for (int pp = 0; pp < TPR; ++pp) l += red[pp * BM + r];
        if (rvalid) {
            if (a.nsplit == 1) {
                const float inv = l > 0.f ? 1.f / l : 0.f;
                float* op = a.out + (int64_t)b * a.o_bs + (int64_t)row * a.o_ts + h * D + part * OD;
#pragma unroll
                for (int i = 0; i < OD; i += 4)
                    *reinterpret_cast<float4*>(op + i) = make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv);
                if (part == 0 && a.lse) a.lse[(int64_t)bh * a.Lq + row] = l > 0.f ? m / kLog2e + logf(l) : -INFINITY;
            } else {
                const int64_t prow = ((int64_t)bh * a.Lq + row) * a.nsplit + split;
                float* op = a.part_o + prow * D + part * OD;
#pragma unroll
                for (int i = 0; i < OD; i += 4) *reinterpret_cast<float4*>(op + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
                if (part == 0) {
                    a.part_ml[prow * 2] = m == -INFINITY ? -INFINITY : m / kLog2e;      // natural-log units, as the combine kernel expects
                    a.part_ml[prow * 2 + 1] = l;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == NSW) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
    }
}

// one thread per (bh, q, d): out = sum_s o_s exp(m_s - M) / sum_s l_s exp(m_s - M)
__global__ void __launch_bounds__(256) attn_t5_combine_kernel(const float* __restrict__ part_o, const float* __restrict__ part_ml,
                                                              float* __restrict__ out, float* __restrict__ lse, int H, int Lq,
                                                              int nsplit, int64_t o_bs, int64_t o_ts, int64_t total, int D) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int d = (int)(i % D);
    const int64_t row = i / D;
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
    if (lse && d == 0) lse[row] = den > 0.f ? M + logf(den) : -INFINITY;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 3-D map over a bf16 plane [B, Lk, H*D] with element strides (bs, ts, 1); the two outer dimensions are ordered by stride
// (tensor maps want ascending strides).  box = {32 | 64 channels, 128 tokens}; tok_dim tells the kernel which coordinate
// is the token axis.
bool make_plane_map(CUtensorMap* m, const void* ptr, int B, int Lk, int E, int64_t bs, int64_t ts, int D, int* tok_dim) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    const bool tok_first = B == 1 || ts <= bs;
    *tok_dim = tok_first ? 1 : 2;
    cuuint64_t dims[3] = {(cuuint64_t)E, (cuuint64_t)(tok_first ? Lk : B), (cuuint64_t)(tok_first ? B : Lk)};
    cuuint64_t strides[2] = {(cuuint64_t)(tok_first ? ts : bs) * 2, (cuuint64_t)(tok_first ? (B == 1 ? ts * Lk : bs) : ts) * 2};
    cuuint32_t box[3] = {(cuuint32_t)(D == 32 ? 32 : 64), (cuuint32_t)(tok_first ? BN : 1), (cuuint32_t)(tok_first ? 1 : BN)};
    cuuint32_t es[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               D == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int pick_splits(int B, int H, int qtiles, int ntiles, int D) {
    const int64_t base = (int64_t)B * H * qtiles;
    const double slots = (D == 32 ? 2.0 : 1.0) * 148.0;
    int best = 1;
    double best_score = -1.0;
    for (int ns = 1; ns <= 32; ns *= 2) {
        if (ns > 1 && ntiles / ns < 2) break;                 // at least two key tiles per split
        const double waves = (double)(base * ns) / slots;
        const double eff = waves / (double)(int64_t)(waves + 0.999999);
        const double score = eff - 0.015 * (ns > 1 ? __builtin_ctz((unsigned)ns) : 0);
        if (score > best_score + 1e-9) { best_score = score; best = ns; }
    }
    return best;
}

template <int D>
int configure() {
    static bool configured[PVSG_MAX_DEVICES];
    if (pvsg_first_use_on_device(configured) &&
        (cudaFuncSetAttribute(attn_t5_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<D>::SMEM) != cudaSuccess ||
         cudaFuncSetAttribute(attn_t5_kernel<D>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared) != cudaSuccess))
        return PVSG_ERR_LAUNCH;
    return PVSG_OK;
}

template <int D>
int launch(const CUtensorMap* maps, const T5Args& a, dim3 grid, cudaStream_t st) {
    if (const int rc = configure<D>()) return rc;
    attn_t5_kernel<D><<<grid, NTHR, Cfg<D>::SMEM, st>>>(maps[0], maps[1], maps[2], maps[3], a);
    return PVSG_OK;
}

}  // namespace

int pvsg_internal::configure_attention_t5() {
    const int rc = configure<32>();
    return rc ? rc : configure<128>();
}

extern "C" int64_t pvsg_attention_t5_workspace_bytes(int B, int H, int Lq, int Lk, int D) {
    const int ns = pick_splits(B, H, (Lq + BM - 1) / BM, (Lk + BN - 1) / BN, D);
    if (ns == 1) return 16;
    return (int64_t)B * H * Lq * ns * (D + 2) * sizeof(float) + 16;
}

static int attention_t5_impl(const float* Q, const void* K_hi, const void* K_lo, const void* V_hi, const void* V_lo,
                             const uint8_t* mask, const int32_t* row_open, float* out, float* lse, void* ws, int B, int H, int Lq, int Lk,
                             int Dh, int64_t q_bs, int64_t q_ts, int64_t k_bs, int64_t k_ts, int64_t v_bs, int64_t v_ts,
                             int64_t o_bs, int64_t o_ts, float scale, void* stream) {
    PVSG_CHECK_ARG(Q && K_hi && K_lo && V_hi && V_lo && out && B > 0 && H > 0 && Lq > 0 && Lk > 0);
    if (Dh != 32 && Dh != 128) return PVSG_ERR_UNSUPPORTED;
    PVSG_CHECK_ARG((k_bs | k_ts | v_bs | v_ts) % 8 == 0 && (q_bs | q_ts | o_bs | o_ts) % 4 == 0);
    PVSG_CHECK_ARG(((reinterpret_cast<uintptr_t>(K_hi) | reinterpret_cast<uintptr_t>(K_lo) | reinterpret_cast<uintptr_t>(V_hi) |
                     reinterpret_cast<uintptr_t>(V_lo) | reinterpret_cast<uintptr_t>(Q) | reinterpret_cast<uintptr_t>(out)) & 15) == 0);
    if (k_bs != v_bs || k_ts != v_ts) return PVSG_ERR_UNSUPPORTED;        // K and V planes share their layout (slices of one projection)
    const int qtiles = (Lq + BM - 1) / BM, ntiles = (Lk + BN - 1) / BN;
    const int ns = pick_splits(B, H, qtiles, ntiles, Dh);
    PVSG_CHECK_ARG(ns == 1 || ws);
    CUtensorMap maps[4];
    int tok_dim = 1, td = 1;
    const void* planes[4] = {K_hi, K_lo, V_hi, V_lo};
    for (int i = 0; i < 4; ++i) {
        if (!make_plane_map(&maps[i], planes[i], B, Lk, H * Dh, k_bs, k_ts, Dh, &td)) return PVSG_ERR_UNSUPPORTED;
        tok_dim = td;
    }
    T5Args a;
    a.Q = Q; a.mask = mask; a.row_open = row_open; a.out = out; a.lse = lse;
    a.part_o = reinterpret_cast<float*>(ws);
    a.part_ml = a.part_o ? a.part_o + (int64_t)B * H * Lq * ns * Dh : nullptr;
    a.H = H; a.Lq = Lq; a.Lk = Lk;
    a.q_bs = q_bs; a.q_ts = q_ts; a.o_bs = o_bs; a.o_ts = o_ts;
    a.scale = scale; a.nsplit = ns; a.tiles_per_split = (ntiles + ns - 1) / ns; a.tok_dim = tok_dim;
    dim3 grid((unsigned)qtiles, (unsigned)ns, (unsigned)(B * H));
    PVSG_CHECK_ARG(grid.z <= 65535);
    cudaStream_t st = as_stream(stream);
    const int rc = Dh == 32 ? launch<32>(maps, a, grid, st) : launch<128>(maps, a, grid, st);
    if (rc != PVSG_OK) return rc;
    if (ns > 1) {
        const int64_t total = (int64_t)B * H * Lq * Dh;
        attn_t5_combine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a.part_o, a.part_ml, out, lse, H, Lq, ns, o_bs, o_ts, total, Dh);
    }
    return pvsg_launch_status();
}

extern "C" int pvsg_attention_t5(const float* Q, const void* K_hi, const void* K_lo, const void* V_hi, const void* V_lo,
                                 const uint8_t* mask, const int32_t* row_open, float* out, void* ws, int B, int H, int Lq, int Lk,
                                 int Dh, int64_t q_bs, int64_t q_ts, int64_t k_bs, int64_t k_ts, int64_t v_bs, int64_t v_ts,
                                 int64_t o_bs, int64_t o_ts, float scale, void* stream) {
    return attention_t5_impl(Q, K_hi, K_lo, V_hi, V_lo, mask, row_open, out, nullptr, ws, B, H, Lq, Lk, Dh, q_bs, q_ts, k_bs, k_ts, v_bs,
                             v_ts, o_bs, o_ts, scale, stream);
}

extern "C" int pvsg_attention_t5_lse(const float* Q, const void* K_hi, const void* K_lo, const void* V_hi, const void* V_lo,
                                     const uint8_t* mask, const int32_t* row_open, float* out, float* lse, void* ws, int B, int H,
                                     int Lq, int Lk, int Dh, int64_t q_bs, int64_t q_ts, int64_t k_bs, int64_t k_ts, int64_t v_bs,
                                     int64_t v_ts, int64_t o_bs, int64_t o_ts, float scale, void* stream) {
    PVSG_CHECK_ARG(lse);
    return attention_t5_impl(Q, K_hi, K_lo, V_hi, V_lo, mask, row_open, out, lse, ws, B, H, Lq, Lk, Dh, q_bs, q_ts, k_bs, k_ts, v_bs,
                             v_ts, o_bs, o_ts, scale, stream);
}
