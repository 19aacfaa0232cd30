// tcgen05 GEMM / implicit-GEMM convolution engine with fp32-grade accuracy ("split-bf16").
//
//   C[M,N] = act(A[M,K] . W[N,K]^T + bias + R)
//
// Every fp32 operand x is carried as two bf16 planes (hi = bf16(x), lo = bf16(x - hi)), the
// same 4 bytes per element as fp32.  One logical product is three tensor-core MMAs
// accumulated in the same TMEM tile:  hi.hi + hi.lo + lo.hi  (the dropped lo.lo term and
// the residual of the split are ~2^-17 relative, i.e. fp32 re-association level).
//
// Persistent kernel, one CTA per SM, 320 threads, tiles of 128 x 128 (or 128 x 256) handed out
// round-robin:
//   warp 0   : TMA producer  -- cp.async.bulk.tensor (128B swizzle) of A_hi, A_lo, W_hi, W_lo
//              k-blocks of 64 into a 3-stage shared-memory ring (mbarrier complete_tx); runs
//              ahead across tile boundaries
//   warp 1   : TMEM allocator + MMA issuer -- one elected lane issues tcgen05.mma
//              (kind::f16, bf16 x bf16 -> fp32, M = 128, N = 128, K = 16) x 3 x 4 per stage into
//              one of TWO TMEM accumulators; tcgen05.commit releases the stage / publishes
//              the accumulator
//   warps 2-9: epilogue -- tcgen05.ld (32 lanes x 32 columns per warp), bias / residual / ReLU,
//              fp32 / split-bf16 boxes staged in shared memory and written by TMA, or sign-mask
//              stores; overlaps the next tile's main loop through the second accumulator
// Convolutions use a 4-D tensor map over the NHWC planes: the M tile is an 8 x 16 patch of
// output pixels and every filter tap is the same TMA box shifted by (r - pad, s - pad); TMA's
// out-of-bounds zero fill is the convolution's zero padding.  No im2col buffer exists.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;             // bf16 elements = 128 bytes = one swizzle row
constexpr int EPI_WARPS = 8;
constexpr int NTHREADS = 64 + 32 * EPI_WARPS;
constexpr int EPI_STAGE_BYTES = 4096;      // per epilogue warp: 32 x 32 fp32, or 32 x 32 bf16 hi | lo
constexpr int PATCH_H = 8, PATCH_W = 16;   // conv M tile = 8 x 16 output pixels
constexpr int A_BYTES = BM * BK * 2;        // 16 KB
// Tile width BN = 64 (4-stage ring; N <= 64: the 64-channel convolutions of layer1 / the stem, where a 128-wide tile spends
// half of every MMA on padding), 128 (3-stage ring) or 256 (2-stage ring).  The wide tile reads each A k-block
// once per 256 output columns instead of once per 128: the engine is L2->SM bandwidth bound
// (four bf16 planes per k-block), so operand bytes per flop, not MMA issue, set its speed.
template <int BN>
struct Cfg {
    static constexpr int STAGES = BN == 64 ? 4 : (BN == 128 ? 3 : 2);
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int DATA_BYTES = STAGES * STAGE_BYTES + EPI_WARPS * EPI_STAGE_BYTES;
    static constexpr int SMEM_TOTAL = DATA_BYTES + 256 /*barriers*/ + 1024 /*align slack*/;
    static constexpr int TMEM_COLS = 2 * BN;   // two accumulators
};

struct TcParams {
    const float* bias;
    const float* R;
    const __nv_bfloat16* R_hi;   // residual carried as split planes (r = hi + lo)
    const __nv_bfloat16* R_lo;
    const void* ident;           // bf16 [256, 256] identity matrix (B operand of the residual k-blocks)
    float* C;
    __nv_bfloat16* C_hi;
    __nv_bfloat16* C_lo;
    uint8_t* mask;
    int32_t* row_open;
    int64_t M, N, ldc, ldr;
    int num_kb;
    int act;
    int tiles_m, tiles_n;
    int batch, tiles_mb;   // batched linear mode: `batch` independent problems, tiles_mb M-tiles each
    int tma_out;   // bit 0: C leaves through TMA stores, bit 1: C_hi / C_lo do, bit 2: fp32 R arrives by TMA
    int res_mma;   // residual planes are added by the tensor core (extra k-blocks against the identity)
    // conv mode
    int conv, OH, OW, cin_kb, S, pad, stride, tiles_h, tiles_w;
};

// ------------------------------------------------------------------ PTX wrappers -------
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
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)
                 : "memory");
    return v;
}
__device__ __forceinline__ uint4 v4u(float a, float b, float c, float d) {
    return make_uint4(__float_as_uint(a), __float_as_uint(b), __float_as_uint(c), __float_as_uint(d));
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);       // start address
    d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
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

// tcgen05.ld issued early (next chunk) and completed later: the wait names the destination
// registers as in/out operands so that no use of them can be scheduled before it.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32_wait(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                   "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]),
                   "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]),
                   "+r"(v[30]), "+r"(v[31])
                 :: "memory");
}

// (x0, x1) -> packed bf16 pairs hi = bf16(x), lo = bf16(x - hi).  The packed conversion
// (cvt.rn.bf16x2.f32 = F2FP, ALU pipe) replaces four single F2F conversions, which issue on the
// quarter-rate XU pipe and dominated the epilogue's instruction time.
__device__ __forceinline__ void split_pair_packed(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(x0, x1);          // .x (low half) = x0
    hi = *reinterpret_cast<const uint32_t*>(&h2);
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(x0 - h0, x1 - h1);
    lo = *reinterpret_cast<const uint32_t*>(&l2);
}

__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

__device__ __forceinline__ int64_t imin64_dev(int64_t a, int64_t b) { return a < b ? a : b; }

struct Tile {
    int64_t m0;       // linear mode: first row
    int n0;
    int tb, oh0, ow0; // conv mode
};

// i-th tile of this CTA: round-robin over all tiles
__device__ __forceinline__ bool next_tile(const TcParams& p, int i, int& mt, int& nt) {
    const int64_t t = (int64_t)blockIdx.x + (int64_t)i * gridDim.x;
    if (t >= (int64_t)p.tiles_m * p.tiles_n) return false;
    nt = (int)(t % p.tiles_n);
    mt = (int)(t / p.tiles_n);
    return true;
}

__device__ __forceinline__ Tile decode_tile(const TcParams& p, int mt, int nt, int bn) {
    Tile tl;
    tl.n0 = nt * bn;
    tl.m0 = (int64_t)mt * BM;
    tl.tb = tl.oh0 = tl.ow0 = 0;
    if (p.conv) {
        const int tw = mt % p.tiles_w;
        mt /= p.tiles_w;
        const int th = mt % p.tiles_h;
        tl.tb = mt / p.tiles_h;
        tl.oh0 = th * PATCH_H;
        tl.ow0 = tw * PATCH_W;
    } else if (p.batch > 1) {
        tl.tb = mt / p.tiles_mb;
        tl.m0 = (int64_t)(mt - tl.tb * p.tiles_mb) * BM;
    }
    return tl;
}

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC_hi,
               const __grid_constant__ CUtensorMap tmC_lo, const __grid_constant__ CUtensorMap tmR,
               const __grid_constant__ CUtensorMap tmRa_hi, const __grid_constant__ CUtensorMap tmRa_lo,
               const __grid_constant__ CUtensorMap tmE, TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int STAGES = Cfg<BN>::STAGES, B_BYTES = Cfg<BN>::B_BYTES, STAGE_BYTES = Cfg<BN>::STAGE_BYTES;
    constexpr int TMEM_COLS = Cfg<BN>::TMEM_COLS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg<BN>::DATA_BYTES);
    uint64_t* full = bars;                    // [STAGES]  TMA -> MMA
    uint64_t* empty = bars + STAGES;          // [STAGES]  MMA -> TMA
    uint64_t* acc_full = bars + 2 * STAGES;   // [2]       MMA -> epilogue
    uint64_t* acc_empty = acc_full + 2;       // [2]       epilogue -> MMA
    uint64_t* r_full = acc_empty + 2;         // [EPI_WARPS] fp32 residual box landed (per epilogue warp)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(r_full + EPI_WARPS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmA_hi); prefetch_tmap(&tmA_lo); prefetch_tmap(&tmB_hi); prefetch_tmap(&tmB_lo);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], EPI_WARPS); }
        for (int w = 0; w < EPI_WARPS; ++w) mbar_init(&r_full[w], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM: two BN-column fp32 accumulators x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // Residual carried as split planes (linear mode): it goes through the TENSOR CORE.  After the
    // K blocks of A.W^T the tile gets BN/64 more k-blocks whose A operand is the residual's own
    // (hi, lo) planes, columns n0 + 64j .., and whose B operand is a slice of the identity matrix:
    // acc += R_hi.I + R_lo.I, exact in the fp32 accumulator.  The residual bytes then ride the deep
    // TMA ring of the main loop instead of a 4 KB-at-a-time epilogue fetch (which capped these
    // HBM-bound layers at ~1.6 TB/s of residual traffic), and the epilogue has no residual at all.
    const int res_kb = p.res_mma ? BN / BK : 0;

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------------
        if (lane == 0) {
            uint32_t it = 0;
            int mt, nt;
            for (int i = 0; next_tile(p, i, mt, nt); ++i) {
                const Tile tl = decode_tile(p, mt, nt, BN);
                for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* st = smem + s * STAGE_BYTES;
                    mbar_expect_tx(&full[s], STAGE_BYTES);
                    if (p.conv) {
                        const int rs = kb / p.cin_kb, cb = kb - rs * p.cin_kb;
                        const int r = rs / p.S, sx = rs - r * p.S;
                        const int iy = tl.oh0 * p.stride + r - p.pad, ix = tl.ow0 * p.stride + sx - p.pad;
                        tma_load_4d(&tmA_hi, &full[s], st, cb * BK, ix, iy, tl.tb);
                        tma_load_4d(&tmA_lo, &full[s], st + A_BYTES, cb * BK, ix, iy, tl.tb);
                    } else if (p.batch > 1) {
                        tma_load_3d(&tmA_hi, &full[s], st, kb * BK, (int)tl.m0, tl.tb);
                        tma_load_3d(&tmA_lo, &full[s], st + A_BYTES, kb * BK, (int)tl.m0, tl.tb);
                    } else {
                        tma_load_2d(&tmA_hi, &full[s], st, kb * BK, (int)tl.m0);
                        tma_load_2d(&tmA_lo, &full[s], st + A_BYTES, kb * BK, (int)tl.m0);
                    }
                    if (p.batch > 1) {   // every problem of the batch has its own W
                        tma_load_3d(&tmB_hi, &full[s], st + 2 * A_BYTES, kb * BK, tl.n0, tl.tb);
                        tma_load_3d(&tmB_lo, &full[s], st + 2 * A_BYTES + B_BYTES, kb * BK, tl.n0, tl.tb);
                    } else {
                        tma_load_2d(&tmB_hi, &full[s], st + 2 * A_BYTES, kb * BK, tl.n0);
                        tma_load_2d(&tmB_lo, &full[s], st + 2 * A_BYTES + B_BYTES, kb * BK, tl.n0);
                    }
                }
                for (int j = 0; j < res_kb; ++j) {
                    if ((int64_t)tl.n0 + j * BK >= p.N) break;      // same rule in the MMA warp
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    ++it;
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* st = smem + s * STAGE_BYTES;
                    mbar_expect_tx(&full[s], 2 * A_BYTES + B_BYTES);
                    tma_load_2d(&tmRa_hi, &full[s], st, tl.n0 + j * BK, (int)tl.m0);
                    tma_load_2d(&tmRa_lo, &full[s], st + A_BYTES, tl.n0 + j * BK, (int)tl.m0);
                    tma_load_2d(&tmE, &full[s], st + 2 * A_BYTES, j * BK, 0);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer --------------------------------------
        if (lane == 0) {
            // instruction descriptor: D = F32, A = B = BF16, both K-major, N = BN, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)(BM >> 4) << 24);
            uint32_t it = 0, ti = 0;
            int mt, nt;
            for (; next_tile(p, (int)ti, mt, nt); ++ti) {
                const uint32_t buf = ti & 1, aph = (ti >> 1) & 1;
                const int n0 = nt * BN;
                mbar_wait(&acc_empty[buf], aph ^ 1);     // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + buf * BN;
                for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&full[s], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES);
                    const uint32_t a_lo = a_hi + A_BYTES;
                    const uint32_t b_hi = a_hi + 2 * A_BYTES;
                    const uint32_t b_lo = b_hi + B_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint32_t off = k * 32;  // 16 bf16 = 32 bytes along K inside the swizzle row
                        const uint64_t dah = umma_desc(a_hi + off), dal = umma_desc(a_lo + off);
                        const uint64_t dbh = umma_desc(b_hi + off), dbl = umma_desc(b_lo + off);
                        umma_bf16(tmem_d, dal, dbh, idesc, (kb | k) != 0);   // small terms first
                        umma_bf16(tmem_d, dah, dbl, idesc, 1);
                        umma_bf16(tmem_d, dah, dbh, idesc, 1);
                    }
                    umma_commit(&empty[s]);          // stage reusable once these MMAs retire
                }
                for (int j = 0; j < res_kb; ++j) {    // acc += R_lo.I + R_hi.I
                    if ((int64_t)n0 + j * BK >= p.N) break;
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    ++it;
                    mbar_wait(&full[s], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = smem_u32(smem + s * STAGE_BYTES);
                    const uint32_t a_lo = a_hi + A_BYTES;
                    const uint32_t b_hi = a_hi + 2 * A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint32_t off = k * 32;
                        const uint64_t dbh = umma_desc(b_hi + off);
                        umma_bf16(tmem_d, umma_desc(a_lo + off), dbh, idesc, 1);
                        umma_bf16(tmem_d, umma_desc(a_hi + off), dbh, idesc, 1);
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&acc_full[buf]);         // accumulator complete
            }
        }
    } else {
        // ---------------- epilogue: warps 2..9 ------------------------------------------------
        // TMEM lane quarter q = warp % 4 (hardware rule); the two warps of a quarter take the even /
        // odd 32-column chunks.  thread = one accumulator row, 32 columns per tcgen05.ld.
        // A lone warp per scheduler cannot hide its own ALU / shared-memory latencies (measured IPC
        // ~0.2 with four epilogue warps), so the epilogue runs EIGHT warps and software-pipelines
        // the TMEM loads; the accumulator goes back to the MMA warp as soon as its last chunk is in
        // registers.
        // Row-per-thread 16-byte global stores reach only ~1.7 TB/s (every store instruction touches
        // 32 lines), so fp32 and split-bf16 outputs are staged in swizzled shared memory (4 KB per
        // warp: 32 x 32 fp32, or 32 x 32 bf16 hi | lo) and written by TMA as whole 32-row boxes
        // (clipped to the matrix by the tensor map); when both kinds are written they take turns in
        // the zone.  An fp32 residual box lands in the zone and is consumed in place (every thread
        // reads exactly the bytes it later overwrites).  Unaligned outputs and the sign-mask bytes
        // keep the direct path.
        const int q = warp & 3, half = (warp - 2) >> 2;
        uint8_t* stage = smem + STAGES * STAGE_BYTES + (warp - 2) * EPI_STAGE_BYTES;
        const uint32_t zone = smem_u32(stage);
        const bool tma_c = (p.tma_out & 1) != 0 && p.C, tma_p = (p.tma_out & 2) != 0 && p.C_hi;
        const bool tma_r = (p.tma_out & 4) != 0;
        const bool use_tma = tma_c || tma_p;
        uint64_t* my_r = &r_full[warp - 2];
        uint32_t rph = 0;
        uint32_t ti = 0;
        int mt, nt;
        for (; next_tile(p, (int)ti, mt, nt); ++ti) {
            const Tile tl = decode_tile(p, mt, nt, BN);
            const uint32_t buf = ti & 1, aph = (ti >> 1) & 1;
            const int nch = (int)((imin64_dev((int64_t)BN, p.N - tl.n0) + 31) >> 5);   // warp-uniform
            const int row_in_tile = q * 32 + lane;
            int64_t out_row;
            bool row_ok;
            if (p.conv) {
                const int oh = tl.oh0 + row_in_tile / PATCH_W, ow = tl.ow0 + row_in_tile % PATCH_W;
                row_ok = oh < p.OH && ow < p.OW;
                out_row = ((int64_t)tl.tb * p.OH + oh) * p.OW + ow;
            } else {
                out_row = tl.m0 + row_in_tile;
                row_ok = out_row < p.M;
                out_row += (int64_t)tl.tb * p.M;     // batched linear: problem tb's rows (tb = 0 otherwise)
            }
            mbar_wait(&acc_full[buf], aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tbase = tmem_base + buf * BN + ((uint32_t)(q * 32) << 16);
            uint32_t v[32];
            if (half < nch) tmem_ld32_issue(tbase + (uint32_t)(half * 32), v);
            else {   // nothing to do in this tile (narrow edge tile): just hand the accumulator back
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[buf]);
            }
            int open = 0;
#pragma unroll 1
            for (int c = half; c < nch; c += 2) {
                const int64_t n = (int64_t)tl.n0 + c * 32;
                if (tma_r) {
                    // fp32 residual box -> zone (after the zone's last store has been read out)
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        mbar_expect_tx(my_r, 32 * 32 * 4);
                        if (p.conv)
                            tma_load_4d(&tmR, my_r, stage, (int)n, tl.ow0, tl.oh0 + (q * 32) / PATCH_W, tl.tb);
                        else
                            tma_load_2d(&tmR, my_r, stage, (int)n, (int)(tl.m0 + q * 32));
                    }
                }
                // bias of this chunk: requested before the TMEM wait so the two latencies overlap
                const bool full_chunk = n + 32 <= p.N;
                const bool bias_vec = p.bias && full_chunk && ((n & 3) == 0);
                float4 b4[8];
                if (bias_vec) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) b4[j] = __ldg(reinterpret_cast<const float4*>(p.bias + n) + j);
                }
                float f[32];
                tmem_ld32_wait(v);
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                if (c + 2 < nch) {
                    tmem_ld32_issue(tbase + (uint32_t)((c + 2) * 32), v);
                } else {
                    // this warp's share of the accumulator is in registers: hand it back
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[buf]);
                }
                if (p.bias) {
                    if (bias_vec) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            f[4 * j] += b4[j].x; f[4 * j + 1] += b4[j].y; f[4 * j + 2] += b4[j].z; f[4 * j + 3] += b4[j].w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (n + j < p.N) f[j] += __ldg(p.bias + n + j);
                    }
                }
                if (tma_r) {
                    mbar_wait(my_r, rph);
                    rph ^= 1;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint4 t4 = lds_v4(zone + lane * 128 + ((j ^ (lane & 7)) << 4));
                        f[4 * j] += __uint_as_float(t4.x); f[4 * j + 1] += __uint_as_float(t4.y);
                        f[4 * j + 2] += __uint_as_float(t4.z); f[4 * j + 3] += __uint_as_float(t4.w);
                    }
                } else if (p.R_hi && !p.res_mma && row_ok) {
                    const __nv_bfloat16* rh = p.R_hi + out_row * p.ldr + n;
                    const __nv_bfloat16* rl = p.R_lo + out_row * p.ldr + n;
                    for (int j = 0; j < 32; ++j)
                        if (n + j < p.N) f[j] += __bfloat162float(rh[j]) + __bfloat162float(rl[j]);
                } else if (p.R && row_ok) {
                    const float* rr = p.R + out_row * p.ldr + n;
                    if (full_chunk && (p.ldr & 3) == 0) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 t4 = __ldg(reinterpret_cast<const float4*>(rr) + j);
                            f[4 * j] += t4.x; f[4 * j + 1] += t4.y; f[4 * j + 2] += t4.z; f[4 * j + 3] += t4.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (n + j < p.N) f[j] += __ldg(rr + j);
                    }
                }
                if (p.act == PVSG_ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
                } else if (p.act == PVSG_ACT_GELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
                }
                const int r0 = q * 32;
                if (p.C) {
                    if (tma_c) {
                        if (!tma_r) {   // (with a residual the zone was already waited for)
                            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                            __syncwarp();
                        }
                        // 128-byte rows, 16-byte pieces XOR-swizzled by (row & 7): conflict-free
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            sts_v4(zone + lane * 128 + ((j ^ (lane & 7)) << 4), v4u(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]));
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) {
                            if (p.conv) tma_store_4d(&tmC, zone, (int)n, tl.ow0, tl.oh0 + r0 / PATCH_W, tl.tb);
                            else if (p.batch > 1) tma_store_3d(&tmC, zone, (int)n, (int)(tl.m0 + r0), tl.tb);
                            else tma_store_2d(&tmC, zone, (int)n, (int)(tl.m0 + r0));
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                    } else if (row_ok) {
                        float* cp = p.C + out_row * p.ldc + n;
                        if (full_chunk && (p.ldc & 3) == 0) {
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                reinterpret_cast<float4*>(cp)[j] =
                                    make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (n + j < p.N) cp[j] = f[j];
                        }
                    }
                }
                if (p.C_hi) {
                    if (tma_p || (row_ok && full_chunk && (p.ldc & 7) == 0)) {
                        __nv_bfloat16* ch = p.C_hi + out_row * p.ldc + n;
                        __nv_bfloat16* cl = p.C_lo + out_row * p.ldc + n;
                        if (tma_p) {
                            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                            __syncwarp();
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint32_t hw[4], lw[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) split_pair_packed(f[8 * j + 2 * e], f[8 * j + 2 * e + 1], hw[e], lw[e]);
                            if (tma_p) {
                                // 64-byte rows, 16-byte pieces XOR-swizzled by ((row >> 1) & 3)
                                const uint32_t off = lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4);
                                sts_v4(zone + off, make_uint4(hw[0], hw[1], hw[2], hw[3]));
                                sts_v4(zone + 2048 + off, make_uint4(lw[0], lw[1], lw[2], lw[3]));
                            } else {
                                reinterpret_cast<uint4*>(ch)[j] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                                reinterpret_cast<uint4*>(cl)[j] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                            }
                        }
                        if (tma_p) {
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            __syncwarp();
                            if (lane == 0) {
                                if (p.conv) {
                                    const int oh = tl.oh0 + r0 / PATCH_W;
                                    tma_store_4d(&tmC_hi, zone, (int)n, tl.ow0, oh, tl.tb);
                                    tma_store_4d(&tmC_lo, zone + 2048, (int)n, tl.ow0, oh, tl.tb);
                                } else {
                                    tma_store_2d(&tmC_hi, zone, (int)n, (int)(tl.m0 + r0));
                                    tma_store_2d(&tmC_lo, zone + 2048, (int)n, (int)(tl.m0 + r0));
                                }
                                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                            }
                        }
                    } else if (row_ok) {
                        __nv_bfloat16* ch = p.C_hi + out_row * p.ldc + n;
                        __nv_bfloat16* cl = p.C_lo + out_row * p.ldc + n;
                        for (int j = 0; j < 32; ++j) {
                            if (n + j < p.N) {
                                const __nv_bfloat16 h = __float2bfloat16_rn(f[j]);
                                ch[j] = h;
                                cl[j] = __float2bfloat16_rn(f[j] - __bfloat162float(h));
                            }
                        }
                    }
                }
                if (p.mask && row_ok) {
                    uint8_t* mm = p.mask + out_row * p.ldc + n;
                    if (full_chunk && (p.ldc & 15) == 0) {
                        uint32_t w[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            w[j] = 0;
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const bool blocked = f[4 * j + e] < 0.f;
                                w[j] |= (blocked ? 1u : 0u) << (8 * e);
                                open += blocked ? 0 : 1;
                            }
                        }
                        reinterpret_cast<uint4*>(mm)[0] = make_uint4(w[0], w[1], w[2], w[3]);
                        reinterpret_cast<uint4*>(mm)[1] = make_uint4(w[4], w[5], w[6], w[7]);
                    } else {
                        for (int j = 0; j < 32; ++j) {
                            if (n + j < p.N) {
                                const bool blocked = f[j] < 0.f;
                                mm[j] = blocked ? 1 : 0;
                                open += blocked ? 0 : 1;
                            }
                        }
                    }
                }
            }
            if (p.mask && p.row_open && row_ok && open) atomicAdd(p.row_open + out_row, open);
        }
        (void)use_tma;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before exit
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                     : "memory");
    }
}

// Small-Cin convolutions (the 7x7/2 RGB stem): patches of the NHWC input are gathered straight into
// operand planes [M, Kpad] (k = (r*S + s)*Cin + c, zero beyond R*S*Cin and outside the image),
// which then go through the GEMM path.  One thread per (pixel, 4 consecutive k).
__global__ void __launch_bounds__(256) im2col_split_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                                           __nv_bfloat16* __restrict__ lo, int H, int W, int Cin,
                                                           int OH, int OW, int S, int stride, int pad, int Kreal,
                                                           int Kpad, int64_t total4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        const int kq = (int)(i % (Kpad / 4));
        int64_t m = i / (Kpad / 4);
        const int ow = (int)(m % OW);
        int64_t t = m / OW;
        const int oh = (int)(t % OH);
        const int64_t b = t / OH;
        uint16_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = kq * 4 + e;
            float v = 0.f;
            if (k < Kreal) {
                const int c = k % Cin, rs = k / Cin;
                const int r = rs / S, s = rs - r * S;
                const int ih = oh * stride - pad + r, iw = ow * stride - pad + s;
                if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = __ldg(x + ((b * H + ih) * W + iw) * (int64_t)Cin + c);
            }
            const __nv_bfloat16 hh = __float2bfloat16_rn(v);
            h[e] = __bfloat16_as_ushort(hh);
            l[e] = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(hh)));
        }
        reinterpret_cast<uint2*>(hi)[i] = make_uint2(h[0] | ((uint32_t)h[1] << 16), h[2] | ((uint32_t)h[3] << 16));
        reinterpret_cast<uint2*>(lo)[i] = make_uint2(l[0] | ((uint32_t)l[1] << 16), l[2] | ((uint32_t)l[3] << 16));
    }
}

// 7x7 stride-2 pad-3 stem (ResNet conv1) without im2col.  The frame (NCHW fp32, C <= 4) is packed
// ONCE into operand planes X2 [B, OH + 3, OW, 64]:
//     X2[b, yp, ox, par*32 + s*4 + c] = x[b, c, 2*yp + par - 3, 2*ox + s - 3]      (0 outside, s = 7, c >= C)
// i.e. a pair of input rows and the 8-pixel window of output column ox per 128-byte row.  Because the
// conv stride (2) equals the row-pair size, output row oy needs exactly the row pairs oy .. oy+3:
// the stem becomes a 4 x 1 convolution with 64 "channels" over X2 and runs through pvsg_conv2d_tc
// (weights re-laid to [Cout, 4, 1, 64]).  Bytes written: 2 planes x 128 B x B (OH+3) OW = 0.49 GB for
// 8 720p frames instead of the 1.45 GB of full im2col rows.
__global__ void __launch_bounds__(256) stem_pack_kernel(const float* __restrict__ x, uint4* __restrict__ hi,
                                                        uint4* __restrict__ lo, int C, int H, int W, int rows2,
                                                        int OW, int64_t total) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int chunk = (int)(i & 7);            // 16-byte piece of the 128-byte row: 2 pixels x 4 channels
        int64_t t = i >> 3;
        const int ox = (int)(t % OW);
        t /= OW;
        const int yp = (int)(t % rows2);
        const int64_t b = t / rows2;
        const int par = chunk >> 2, s0 = (chunk & 3) * 2;
        const int y = 2 * yp + par - 3;
        uint16_t h[8], l[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int s = s0 + (e >> 2), c = e & 3;
            const int xx = 2 * ox + s - 3;
            float v = 0.f;
            if (c < C && s < 7 && y >= 0 && y < H && xx >= 0 && xx < W) v = __ldg(x + ((b * C + c) * H + y) * (int64_t)W + xx);
            const __nv_bfloat16 hh = __float2bfloat16_rn(v);
            h[e] = __bfloat16_as_ushort(hh);
            l[e] = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(hh)));
        }
        hi[i] = make_uint4(h[0] | ((uint32_t)h[1] << 16), h[2] | ((uint32_t)h[3] << 16), h[4] | ((uint32_t)h[5] << 16),
                           h[6] | ((uint32_t)h[7] << 16));
        lo[i] = make_uint4(l[0] | ((uint32_t)l[1] << 16), l[2] | ((uint32_t)l[3] << 16), l[4] | ((uint32_t)l[5] << 16),
                           l[6] | ((uint32_t)l[7] << 16));
    }
}

// fp32 -> (hi, lo) bf16 planes, optionally of x + x2
__global__ void __launch_bounds__(256) split_kernel(const float* __restrict__ x, const float* __restrict__ x2,
                                                    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                    int64_t n4) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        if (x2) {
            const float4 u = __ldg(reinterpret_cast<const float4*>(x2) + i);
            v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
        }
        const float a[4] = {v.x, v.y, v.z, v.w};
        uint16_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const __nv_bfloat16 hh = __float2bfloat16_rn(a[e]);
            h[e] = __bfloat16_as_ushort(hh);
            l[e] = __bfloat16_as_ushort(__float2bfloat16_rn(a[e] - __bfloat162float(hh)));
        }
        reinterpret_cast<uint2*>(hi)[i] = make_uint2(h[0] | ((uint32_t)h[1] << 16), h[2] | ((uint32_t)h[3] << 16));
        reinterpret_cast<uint2*>(lo)[i] = make_uint2(l[0] | ((uint32_t)l[1] << 16), l[2] | ((uint32_t)l[3] << 16));
    }
}

// ------------------------------------------------------------------ host side ----------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int sm_count() {
    static int n[PVSG_MAX_DEVICES];   // per device: a process may drive several GPUs
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= PVSG_MAX_DEVICES) dev = 0;
    if (n[dev] == 0 && (cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n[dev] <= 0))
        n[dev] = 148;
    return n[dev];
}

// 2-D map over a row-major [rows, cols] bf16 matrix with row pitch ld (elements); box = [box_rows, 64]
bool make_map_2d(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 4-D map over NHWC bf16 [B, H, W, C]; box = [1, 8, 16, 64] output pixels; for a strided conv the
// box spans 8*stride x 16*stride input pixels traversed with elementStrides = stride
bool make_map_4d(CUtensorMap* m, const void* ptr, int B, int H, int W, int C, int stride) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)(PATCH_W * stride), (cuuint32_t)(PATCH_H * stride), 1};
    cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Output maps for the TMA-store epilogue: one epilogue warp stores a box of 32 rows x 32 columns
// (linear: 32 consecutive rows; conv: 2 x 16 output pixels), 128-byte (fp32) or 64-byte (bf16) rows.
bool make_out_map(CUtensorMap* m, const void* cptr, bool f32, const TcParams& p, int B, int64_t ld) {
    void* ptr = const_cast<void*>(cptr);
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    const cuuint64_t es = f32 ? 4 : 2;
    const CUtensorMapDataType dt = f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const CUtensorMapSwizzle sw = f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    cuuint32_t ones[4] = {1, 1, 1, 1};
    if (p.conv) {
        cuuint64_t dims[4] = {(cuuint64_t)p.N, (cuuint64_t)p.OW, (cuuint64_t)p.OH, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)ld * es, (cuuint64_t)p.OW * ld * es,
                                 (cuuint64_t)p.OH * p.OW * ld * es};
        cuuint32_t box[4] = {32, (cuuint32_t)PATCH_W, (cuuint32_t)(32 / PATCH_W), 1};
        return enc(m, dt, 4, ptr, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    cuuint64_t dims[2] = {(cuuint64_t)p.N, (cuuint64_t)p.M};
    cuuint64_t strides[1] = {(cuuint64_t)ld * es};
    cuuint32_t box[2] = {32, 32};
    return enc(m, dt, 2, ptr, dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// 3-D map over `batch` row-major [rows, cols] bf16 matrices (pitch ld, batch stride bs, elements)
bool make_map_3d(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int64_t bs, int batch,
                 int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)bs * 2};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)box_rows, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// fp32 output of the batched mode: [batch, M, N] with row pitch ld, 32 x 32 boxes
bool make_out_map_3d(CUtensorMap* m, const void* ptr, int64_t M, int64_t N, int64_t ld, int batch) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)M, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)M * ld * 4};
    cuuint32_t box[3] = {32, 32, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int BN>
int configure_tc() {
    static bool configured[PVSG_MAX_DEVICES];
    if (pvsg_first_use_on_device(configured) &&
        cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM_TOTAL) != cudaSuccess)
        return PVSG_ERR_LAUNCH;
    return PVSG_OK;
}

template <int BN>
int launch_tc(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& b_hi, const CUtensorMap& b_lo,
              const TcParams& p, int B, cudaStream_t st) {
    if (const int rc = configure_tc<BN>()) return rc;
    const int64_t tiles = (int64_t)p.tiles_m * p.tiles_n;
    const unsigned grid = (unsigned)imin64(tiles, sm_count());
    TcParams q = p;
    static const bool direct = getenv("PVSG_TC_DIRECT_STORE") != nullptr;
    CUtensorMap c{}, c_hi{}, c_lo{}, r{}, ra_hi{}, ra_lo{}, e{};
    q.tma_out = 0;
    q.res_mma = 0;
    if (p.batch > 1) {
        if (!direct && p.C && p.ldc % 4 == 0 && al16(p.C) && make_out_map_3d(&c, p.C, p.M, p.N, p.ldc, p.batch))
            q.tma_out |= 1;
    } else if (!direct && p.C && p.ldc % 4 == 0 && al16(p.C) && make_out_map(&c, p.C, true, p, B, p.ldc)) q.tma_out |= 1;
    if (p.batch <= 1 && !direct && p.C_hi && p.ldc % 8 == 0 && al16(p.C_hi) && al16(p.C_lo) &&
        make_out_map(&c_hi, p.C_hi, false, p, B, p.ldc) && make_out_map(&c_lo, p.C_lo, false, p, B, p.ldc))
        q.tma_out |= 2;
    if (p.batch <= 1 && !direct && p.R && p.ldr % 4 == 0 && al16(p.R) && make_out_map(&r, p.R, true, p, B, p.ldr))
        q.tma_out |= 4;
    // residual planes as extra k-blocks against the identity (linear mode; operand-style maps)
    if (p.R_hi && !p.conv && p.ident && p.ldr % 8 == 0 && al16(p.R_hi) && al16(p.R_lo) && al16(p.ident) &&
        make_map_2d(&ra_hi, p.R_hi, p.M, p.N, p.ldr, BM) && make_map_2d(&ra_lo, p.R_lo, p.M, p.N, p.ldr, BM) &&
        make_map_2d(&e, p.ident, 256, 256, 256, BN))
        q.res_mma = 1;
    gemm_tc_kernel<BN><<<grid, NTHREADS, Cfg<BN>::SMEM_TOTAL, st>>>(a_hi, a_lo, b_hi, b_lo, c, c_hi, c_lo, r, ra_hi, ra_lo,
                                                                    e, q);
    return pvsg_launch_status();
}

// Wide tiles when they add no padding along N and still leave >= 4 tiles per SM (wave quantisation).
int pick_bn(int64_t tiles_m, int64_t N) {
    static const int forced = [] { const char* e = getenv("PVSG_TC_BN"); return e ? atoi(e) : 0; }();
    if (forced == 128 || forced == 256) return forced;
    static const bool narrow = getenv("PVSG_TC_NO_BN64") == nullptr;
    if (narrow && N <= 64) return 64;
    const int64_t t256 = (N + 255) / 256;
    if (t256 * 256 != ((N + 127) / 128) * 128) return 128;
    return tiles_m * t256 >= 4 * (int64_t)sm_count() ? 256 : 128;
}

}  // namespace

int pvsg_internal::configure_gemm_tc() {
    int rc = configure_tc<64>();
    if (!rc) rc = configure_tc<128>();
    return rc ? rc : configure_tc<256>();
}

extern "C" int pvsg_im2col_split(const float* x, void* hi, void* lo, int B, int H, int W, int Cin, int R, int S,
                                 int stride, int pad, int Kpad, void* stream) {
    PVSG_CHECK_ARG(x && hi && lo && B > 0 && H > 0 && W > 0 && Cin > 0 && R > 0 && S > 0 && stride > 0 && pad >= 0);
    const int Kreal = R * S * Cin;
    PVSG_CHECK_ARG(Kpad >= Kreal && Kpad % 4 == 0);
    const int OH = (H + 2 * pad - R) / stride + 1, OW = (W + 2 * pad - S) / stride + 1;
    PVSG_CHECK_ARG(OH > 0 && OW > 0);
    const int64_t total4 = (int64_t)B * OH * OW * (Kpad / 4);
    im2col_split_kernel<<<(unsigned)imin64((total4 + 255) / 256, 148 * 32), 256, 0, as_stream(stream)>>>(
        x, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), H, W, Cin, OH, OW, S, stride,
        pad, Kreal, Kpad, total4);
    return pvsg_launch_status();
}

extern "C" int pvsg_stem7x7s2_pack(const float* x_nchw, void* hi, void* lo, int B, int C, int H, int W,
                                    void* stream) {
    PVSG_CHECK_ARG(x_nchw && hi && lo && B > 0 && C > 0 && C <= 4 && H > 0 && W > 0 && al16(hi) && al16(lo));
    const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
    const int64_t total = (int64_t)B * (OH + 3) * OW * 8;
    stem_pack_kernel<<<(unsigned)imin64((total + 255) / 256, 148 * 32), 256, 0, as_stream(stream)>>>(
        x_nchw, reinterpret_cast<uint4*>(hi), reinterpret_cast<uint4*>(lo), C, H, W, OH + 3, OW, total);
    return pvsg_launch_status();
}

extern "C" int pvsg_split_bf16(const float* x, const float* x2, void* hi, void* lo, int64_t n, void* stream) {
    PVSG_CHECK_ARG(x && hi && lo && n > 0 && n % 4 == 0 && al16(x) && (!x2 || al16(x2)));
    const int64_t n4 = n / 4;
    split_kernel<<<(unsigned)imin64((n4 + 255) / 256, 148 * 16), 256, 0, as_stream(stream)>>>(
        x, x2, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), n4);
    return pvsg_launch_status();
}

extern "C" int pvsg_linear_tc(const void* A_hi, const void* A_lo, int64_t lda, const void* W_hi, const void* W_lo,
                              int64_t ldw, const float* bias, const float* R, int64_t ldr, float* C, void* C_hi,
                              void* C_lo, uint8_t* mask, int32_t* row_open, int64_t ldc, int64_t M, int64_t N,
                              int64_t K, int act, const void* R_hi, const void* R_lo, const void* ident,
                              void* stream) {
    PVSG_CHECK_ARG(A_hi && A_lo && W_hi && W_lo && (C || C_hi || mask) && M > 0 && N > 0 && K > 0);
    PVSG_CHECK_ARG((C_hi == nullptr) == (C_lo == nullptr));
    PVSG_CHECK_ARG((R_hi == nullptr) == (R_lo == nullptr) && !(R && R_hi));
    PVSG_CHECK_ARG(lda >= K && ldw >= K && ldc >= N && ((!R && !R_hi) || ldr >= N));
    if (K % BK != 0 || lda % 8 != 0 || ldw % 8 != 0 || !al16(A_hi) || !al16(A_lo) || !al16(W_hi) || !al16(W_lo))
        return PVSG_ERR_UNSUPPORTED;
    if (M > 0x7fffffffLL || N > 0x7fffffffLL || ((M + BM - 1) / BM) * ((N + 127) / 128) > 0x7fffffffLL)
        return PVSG_ERR_UNSUPPORTED;
    const int bn = pick_bn((M + BM - 1) / BM, N);
    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
    if (!make_map_2d(&ta_hi, A_hi, M, K, lda, BM) || !make_map_2d(&ta_lo, A_lo, M, K, lda, BM) ||
        !make_map_2d(&tb_hi, W_hi, N, K, ldw, bn) || !make_map_2d(&tb_lo, W_lo, N, K, ldw, bn))
        return PVSG_ERR_LAUNCH;
    TcParams p{};
    p.bias = bias; p.R = R; p.C = C;
    p.R_hi = reinterpret_cast<const __nv_bfloat16*>(R_hi); p.R_lo = reinterpret_cast<const __nv_bfloat16*>(R_lo);
    p.ident = ident;
    p.C_hi = reinterpret_cast<__nv_bfloat16*>(C_hi); p.C_lo = reinterpret_cast<__nv_bfloat16*>(C_lo);
    p.mask = mask; p.row_open = row_open;
    p.M = M; p.N = N; p.ldc = ldc; p.ldr = ldr; p.num_kb = (int)(K / BK); p.act = act; p.conv = 0;
    p.tiles_m = (int)((M + BM - 1) / BM); p.tiles_n = (int)((N + bn - 1) / bn);
    if (bn == 64) return launch_tc<64>(ta_hi, ta_lo, tb_hi, tb_lo, p, 1, as_stream(stream));
    return bn == 256 ? launch_tc<256>(ta_hi, ta_lo, tb_hi, tb_lo, p, 1, as_stream(stream))
                     : launch_tc<128>(ta_hi, ta_lo, tb_hi, tb_lo, p, 1, as_stream(stream));
}

extern "C" int pvsg_linear_tc_batched(const void* A_hi, const void* A_lo, int64_t lda, int64_t a_bs, const void* W_hi,
                                      const void* W_lo, int64_t ldw, int64_t w_bs, float* C, uint8_t* mask,
                                      int32_t* row_open, int64_t ldc, int batch, int64_t M, int64_t N, int64_t K,
                                      void* stream) {
    PVSG_CHECK_ARG(A_hi && A_lo && W_hi && W_lo && (C || mask) && batch > 0 && M > 0 && N > 0 && K > 0);
    // batch strides: whole matrices ([batch, M, lda]) or K-chunks of ONE matrix (split-K views: stride = chunk length)
    PVSG_CHECK_ARG(lda >= K && ldw >= K && ldc >= N && (a_bs >= M * lda || a_bs >= K) && (w_bs >= N * ldw || w_bs >= K));
    if (K % BK != 0 || (lda | ldw | a_bs | w_bs) % 8 != 0 || !al16(A_hi) || !al16(A_lo) || !al16(W_hi) || !al16(W_lo))
        return PVSG_ERR_UNSUPPORTED;
    const int64_t tiles_mb = (M + BM - 1) / BM;
    if (M > 0x7fffffffLL || N > 0x7fffffffLL || batch * tiles_mb * ((N + 127) / 128) > 0x7fffffffLL)
        return PVSG_ERR_UNSUPPORTED;
    const int bn = pick_bn(batch * tiles_mb, N);
    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
    // rank-3 maps for a real batch; batch 1 is the plain 2-D problem
    if (batch > 1) {
        if (!make_map_3d(&ta_hi, A_hi, M, K, lda, a_bs, batch, BM) || !make_map_3d(&ta_lo, A_lo, M, K, lda, a_bs, batch, BM) ||
            !make_map_3d(&tb_hi, W_hi, N, K, ldw, w_bs, batch, bn) || !make_map_3d(&tb_lo, W_lo, N, K, ldw, w_bs, batch, bn))
            return PVSG_ERR_LAUNCH;
    } else if (!make_map_2d(&ta_hi, A_hi, M, K, lda, BM) || !make_map_2d(&ta_lo, A_lo, M, K, lda, BM) ||
               !make_map_2d(&tb_hi, W_hi, N, K, ldw, bn) || !make_map_2d(&tb_lo, W_lo, N, K, ldw, bn)) {
        return PVSG_ERR_LAUNCH;
    }
    TcParams p{};
    p.C = C; p.mask = mask; p.row_open = row_open;
    p.M = M; p.N = N; p.ldc = ldc; p.num_kb = (int)(K / BK); p.act = PVSG_ACT_NONE; p.conv = 0;
    p.batch = batch; p.tiles_mb = (int)tiles_mb;
    p.tiles_m = (int)(batch * tiles_mb); p.tiles_n = (int)((N + bn - 1) / bn);
    if (bn == 64) return launch_tc<64>(ta_hi, ta_lo, tb_hi, tb_lo, p, 1, as_stream(stream));
    return bn == 256 ? launch_tc<256>(ta_hi, ta_lo, tb_hi, tb_lo, p, 1, as_stream(stream))
                     : launch_tc<128>(ta_hi, ta_lo, tb_hi, tb_lo, p, 1, as_stream(stream));
}

extern "C" int pvsg_conv2d_tc(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo,
                              const float* bias, const float* residual, float* y, void* y_hi, void* y_lo, int B,
                              int H, int W, int Cin, int Cout, int R, int S, int stride, int pad, int act,
                              const void* res_hi, const void* res_lo, void* stream) {
    PVSG_CHECK_ARG(x_hi && x_lo && w_hi && w_lo && (y || y_hi) && B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0);
    PVSG_CHECK_ARG((res_hi == nullptr) == (res_lo == nullptr) && !(residual && res_hi));
    PVSG_CHECK_ARG((y_hi == nullptr) == (y_lo == nullptr) && R > 0 && S > 0 && pad >= 0);
    if (Cin % BK != 0 || !al16(x_hi) || !al16(x_lo) || !al16(w_hi) || !al16(w_lo)) return PVSG_ERR_UNSUPPORTED;
    if (stride < 1 || stride > 2) return PVSG_ERR_UNSUPPORTED;
    const int OH = (H + 2 * pad - R) / stride + 1, OW = (W + 2 * pad - S) / stride + 1;
    PVSG_CHECK_ARG(OH > 0 && OW > 0);
    const int64_t K = (int64_t)R * S * Cin;
    const int tiles_h = (OH + PATCH_H - 1) / PATCH_H, tiles_w = (OW + PATCH_W - 1) / PATCH_W;
    const int64_t tiles_m = (int64_t)B * tiles_h * tiles_w;
    const int bn = pick_bn(tiles_m, Cout);
    CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
    if (!make_map_4d(&ta_hi, x_hi, B, H, W, Cin, stride) || !make_map_4d(&ta_lo, x_lo, B, H, W, Cin, stride) ||
        !make_map_2d(&tb_hi, w_hi, Cout, K, K, bn) || !make_map_2d(&tb_lo, w_lo, Cout, K, K, bn))
        return PVSG_ERR_LAUNCH;
    TcParams p{};
    p.bias = bias; p.R = residual; p.C = y;
    p.R_hi = reinterpret_cast<const __nv_bfloat16*>(res_hi); p.R_lo = reinterpret_cast<const __nv_bfloat16*>(res_lo);
    p.C_hi = reinterpret_cast<__nv_bfloat16*>(y_hi); p.C_lo = reinterpret_cast<__nv_bfloat16*>(y_lo);
    p.M = (int64_t)B * OH * OW; p.N = Cout; p.ldc = Cout; p.ldr = Cout; p.act = act;
    p.num_kb = (int)(K / BK); p.conv = 1; p.OH = OH; p.OW = OW; p.cin_kb = Cin / BK; p.S = S; p.pad = pad; p.stride = stride;
    p.tiles_h = tiles_h; p.tiles_w = tiles_w;
    if (tiles_m * ((Cout + 127) / 128) > 0x7fffffffLL) return PVSG_ERR_UNSUPPORTED;
    p.tiles_m = (int)tiles_m; p.tiles_n = (Cout + bn - 1) / bn;
    if (bn == 64) return launch_tc<64>(ta_hi, ta_lo, tb_hi, tb_lo, p, B, as_stream(stream));
    return bn == 256 ? launch_tc<256>(ta_hi, ta_lo, tb_hi, tb_lo, p, B, as_stream(stream))
                     : launch_tc<128>(ta_hi, ta_lo, tb_hi, tb_lo, p, B, as_stream(stream));
}
