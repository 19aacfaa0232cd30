// Multi-scale deformable attention, region-tiled: TMA-staged value windows in shared memory.
//
// The encoder's queries ARE the pyramid tokens, and their sampling points stay within a few pixels of their own
// position on every level.  A CTA therefore owns one REGION of the image (16 x 8 pixels of the finest level) for
// one (frame, head) and serves every query whose reference point lies in it -- 128 tokens of the finest level,
// 32 of the middle one, 8 of the coarsest -- from three windows (region footprint + halo, one per level) of the
// head's value plane that ONE elected thread fetches with three 5-D `cp.async.bulk.tensor` loads
// (box = 32 channels x 1 head x WX x WY pixels; out-of-map pixels are zero-filled by TMA, which IS the op's zero
// padding, so in-window samples need no per-corner validity logic).  All bilinear gathers then hit shared memory:
// one lane group of 8 (a float4 of the 32 channels each) per (query, head), 12 samples x 4 corners = 48 LDS.128
// wavefronts per item, no L1 tag stage, no L1 misses.  Samples whose footprint leaves the window (large offsets;
// ~1 % with the halo below) take a compacted second pass through global memory.
//
// What bounds it: the shared-memory / L1 crossbar moves 128 B per clock per SM and an item needs 48 corner lines
// of 128 B (fp32 values), i.e. >= 48 wavefronts per (query, head) = 27 us per 720p frame and layer against 9.5 us
// for the algorithmic HBM bytes -- see DESIGN.md section 4.  The kernel removes everything else from that pipe:
// per-sample parameters travel through a per-warp staging area read with broadcast LDS.128 (the previous kernel
// broadcast them with 60 warp shuffles per item, which use the same crossbar), and window fills are 9 % of the
// gather traffic instead of the 24 % L1 miss rate.
#include <cuda.h>

#include "common.cuh"

namespace {

constexpr int RW = 16, RH = 8;            // region on the finest level
constexpr int NWARP = 7, NTHR = 32 * NWARP;   // 28 lane groups x 6 passes = 168 = 128 + 32 + 8 items per region
constexpr int ITEMS = RW * RH + RW * RH / 4 + RW * RH / 16;
constexpr int PASSES = 6;
static_assert(ITEMS == NWARP * 4 * PASSES, "items per region must fill the lane groups exactly");
constexpr int LP = 12, NP = 4, NH = 8;   // 3 levels x 4 points, 8 heads (the reference configuration)
constexpr int LPAD = 13;                  // float4 stride of an item's weight records: 52 words, the 4 groups hit 4 bank quads

template <int HX, int HY>
struct Win {
    // index = level (0 = coarsest ... 2 = finest); sizes cover offsets in [-H, +H] around every query of the region
    static constexpr int wx(int l) { return l == 2 ? RW + 2 * HX + 1 : (l == 1 ? RW / 2 + 2 * HX + 2 : RW / 4 + 2 * HX + 2); }
    static constexpr int wy(int l) { return l == 2 ? RH + 2 * HY + 1 : (l == 1 ? RH / 2 + 2 * HY + 2 : RH / 4 + 2 * HY + 2); }
    static constexpr int px(int l) { return wx(l) * wy(l); }
    static constexpr int off(int l) { return l == 0 ? 0 : (l == 1 ? px(0) : px(0) + px(1)); }   // in pixels
    static constexpr int PIXELS = px(0) + px(1) + px(2);
    static constexpr int WIN_BYTES = PIXELS * 128;
    // per warp: main records (12 x float4 weights + 12 bases) and fallback records (the same), for 4 items
    static constexpr int STAGE_BYTES_PER_WARP = 4 * (LPAD * 16 + LP * 4) + 4 * (LP * 16 + LP * 4);
    static constexpr int SMEM = WIN_BYTES + NWARP * STAGE_BYTES_PER_WARP + 64;
};

struct TileLevels {
    int h[3], w[3];
    int start[3];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint64_t* bar, uint32_t dst, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

struct Sample {   // one bilinear sample, prepared by the lane that owns it
    float w[4];   // corner weights (attention weight folded in) for the shared-memory pass; 0 if it goes elsewhere
    int base;     // float4 index of corner (y0, x0) inside the CTA's windows
    float gw[4];  // global-memory fallback: fixed-corner weights ...
    int gbase;    // ... and pixel index of the clamped corner in the value tensor; -1 = no fallback needed
};

template <int HX, int HY>
__device__ __forceinline__ void prep_sample(const TileLevels& lv, int l, int ox, int oy, float x, float y, float aw,
                                            Sample& s) {
    using W = Win<HX, HY>;
    const int hgt = lv.h[l], wid = lv.w[l];
    x = fminf(fmaxf(x, -2.f), (float)wid + 1.f);
    y = fminf(fmaxf(y, -2.f), (float)hgt + 1.f);
    const float fy = floorf(y), fx = floorf(x);
    const int y0 = (int)fy, x0 = (int)fx;
    const float ly = y - fy, lx = x - fx;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const int wxp = x0 - ox, wyp = y0 - oy;
    const int WX = l == 2 ? W::wx(2) : (l == 1 ? W::wx(1) : W::wx(0));
    const int WY = l == 2 ? W::wy(2) : (l == 1 ? W::wy(1) : W::wy(0));
    const int woff = l == 2 ? W::off(2) : (l == 1 ? W::off(1) : W::off(0));
    const bool touches = x0 >= -1 && x0 < wid && y0 >= -1 && y0 < hgt;     // some corner lies inside the map
    const bool inwin = wxp >= 0 && wxp <= WX - 2 && wyp >= 0 && wyp <= WY - 2;
    const bool smem = touches && inwin;
    // shared-memory pass: plain bilinear weights, out-of-map corners read TMA's zero fill
    s.w[0] = smem ? hy * hx * aw : 0.f; s.w[1] = smem ? hy * lx * aw : 0.f;
    s.w[2] = smem ? ly * hx * aw : 0.f; s.w[3] = smem ? ly * lx * aw : 0.f;
    s.base = smem ? (woff + wyp * WX + wxp) * 8 : 0;
    s.gbase = -1;
    if (touches && !inwin) {
        // fixed corners (yb, xb) .. (yb + 1, xb + 1) clamped into the map, padding folded into the weights
        const int xb = min(max(x0, 0), wid - 2), yb = min(max(y0, 0), hgt - 2);
        const float wxa = (x0 == xb) ? hx : ((x0 + 1 == xb) ? lx : 0.f);
        const float wxb = (x0 == xb) ? lx : ((x0 == xb + 1) ? hx : 0.f);
        const float wya = (y0 == yb) ? hy : ((y0 + 1 == yb) ? ly : 0.f);
        const float wyb = (y0 == yb) ? ly : ((y0 == yb + 1) ? hy : 0.f);
        s.gbase = lv.start[l] + yb * wid + xb;
        s.gw[0] = wya * wxa * aw; s.gw[1] = wya * wxb * aw;
        s.gw[2] = wyb * wxa * aw; s.gw[3] = wyb * wxb * aw;
    }
}

__device__ __forceinline__ void fma4(float4& acc, float w, const float4& v) {
    acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
    acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
}

template <int HX, int HY>
__global__ void __launch_bounds__(NTHR, 2)
msda_tile_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                 const __grid_constant__ CUtensorMap map2, const float* __restrict__ value, TileLevels lv,
                 const float* __restrict__ proj, const float* __restrict__ ref, float* __restrict__ out,
                 uint2* __restrict__ out_hi, uint2* __restrict__ out_lo, int B, int N, int regions_x, int regions_y) {
    using W = Win<HX, HY>;
    extern __shared__ __align__(128) uint8_t smem[];
    const float4* win = reinterpret_cast<const float4*>(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 3, j = lane & 7;
    uint8_t* stage = smem + W::WIN_BYTES + warp * W::STAGE_BYTES_PER_WARP;
    float4* st_w = reinterpret_cast<float4*>(stage);                       // [4 items][LPAD]
    int* st_b = reinterpret_cast<int*>(stage + 4 * LPAD * 16);             // [4 items][12]
    float4* fb_w = reinterpret_cast<float4*>(stage + 4 * LPAD * 16 + 4 * LP * 4);   // [4 items][12] compacted
    int* fb_b = reinterpret_cast<int*>(stage + 4 * LPAD * 16 + 4 * LP * 4 + 4 * LP * 16);   // [4 items][12]
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + W::WIN_BYTES + NWARP * W::STAGE_BYTES_PER_WARP);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map2) : "memory");
    }
    __syncthreads();
    uint32_t phase = 0;
    const int regions = regions_x * regions_y;
    const int64_t total = (int64_t)B * regions * NH;
    const float4* vglob = reinterpret_cast<const float4*>(value);
    for (int64_t wi = blockIdx.x; wi < total; wi += gridDim.x) {
        const int head = (int)(wi % NH);
        const int64_t t = wi / NH;
        const int reg = (int)(t % regions);
        const int b = (int)(t / regions);
        const int ry = reg / regions_x, rx = reg - ry * regions_x;
        // window origins per level (exact 2x pyramid; see Win)
        const int ox2 = RW * rx - HX, oy2 = RH * ry - HY;
        const int ox1 = RW / 2 * rx - 1 - HX, oy1 = RH / 2 * ry - 1 - HY;
        const int ox0 = RW / 4 * rx - 1 - HX, oy0 = RH / 4 * ry - 1 - HY;
        if (threadIdx.x == 0) {
            // every warp has left the previous region (the __syncthreads at the loop end); order those generic-proxy
            // reads before the async-proxy writes of the new windows
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bar, (uint32_t)W::WIN_BYTES);
            tma_load_5d(&map0, bar, smem_u32(smem) + W::off(0) * 128, 0, head, ox0, oy0, b);
            tma_load_5d(&map1, bar, smem_u32(smem) + W::off(1) * 128, 0, head, ox1, oy1, b);
            tma_load_5d(&map2, bar, smem_u32(smem) + W::off(2) * 128, 0, head, ox2, oy2, b);
        }
        const float4* vb = vglob + ((int64_t)b * N * NH + head) * 8 + j;     // global fallback base of this lane
        // Query of a lane group in a pass, and its raw projections.  Lane i < 6 of the group owns samples 2i and
        // 2i + 1 (one level: l = i / 2): ONE 16-byte load brings both offset pairs, one 8-byte load both logits, so a
        // warp touches 4 + 4 lines per pass.  The loads of pass p + 1 are issued before the gathers of pass p
        // (their L2 / HBM latency was the top stall of the first version of this kernel).
        struct Raw { float4 off; float2 lg; float2 rf; int nq; bool valid; };
        auto fetch = [&](int pass) {
            Raw r;
            const int it = pass * (NWARP * 4) + warp * 4 + g;
            int ql, qy, qx;
            if (it < RW * RH) { ql = 2; qy = RH * ry + it / RW; qx = RW * rx + it % RW; }
            else if (it < RW * RH + RW * RH / 4) { const int k = it - RW * RH; ql = 1; qy = RH / 2 * ry + k / (RW / 2); qx = RW / 2 * rx + k % (RW / 2); }
            else { const int k = it - RW * RH - RW * RH / 4; ql = 0; qy = RH / 4 * ry + k / (RW / 4); qx = RW / 4 * rx + k % (RW / 4); }
            r.valid = qy < lv.h[ql] && qx < lv.w[ql];
            r.nq = r.valid ? lv.start[ql] + qy * lv.w[ql] + qx : 0;
            const float* prow = proj + ((int64_t)b * N + r.nq) * (NH * LP * 3);
            const int i = j < 6 ? j : 0;
            r.off = __ldg(reinterpret_cast<const float4*>(prow + head * LP * 2) + i);
            r.lg = __ldg(reinterpret_cast<const float2*>(prow + NH * LP * 2 + head * LP) + i);
            r.rf = __ldg(reinterpret_cast<const float2*>(ref) + r.nq);
            return r;
        };
        Raw nxt = fetch(0);
        bool waited = false;
#pragma unroll 1
        for (int pass = 0; pass < PASSES; ++pass) {
            const Raw cur = nxt;
            if (pass + 1 < PASSES) nxt = fetch(pass + 1);
            const bool valid = cur.valid;
            const int nq = cur.nq;
            const bool has = j < 6;
            // ---- softmax over the 12 logits (two per lane), locations, per-sample records
            const float lg0 = has ? cur.lg.x : -INFINITY, lg1 = has ? cur.lg.y : -INFINITY;
            float mx = fmaxf(lg0, lg1);
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const float e0 = has ? __expf(lg0 - mx) : 0.f, e1 = has ? __expf(lg1 - mx) : 0.f;
            float sum = e0 + e1;
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float inv = 1.f / sum;
            const int l = min(j >> 1, 2);
            const int ox = l == 0 ? ox0 : (l == 1 ? ox1 : ox2), oy = l == 0 ? oy0 : (l == 1 ? oy1 : oy2);
            const float fw = (float)lv.w[l], fh = (float)lv.h[l];
            Sample s0, s1;
            // loc = ref + off / (w, h); pixel = loc * size - 0.5 = ref * size + off - 0.5
            prep_sample<HX, HY>(lv, l, ox, oy, fmaf(cur.rf.x, fw, cur.off.x) - 0.5f, fmaf(cur.rf.y, fh, cur.off.y) - 0.5f,
                                (valid && has) ? e0 * inv : 0.f, s0);
            prep_sample<HX, HY>(lv, l, ox, oy, fmaf(cur.rf.x, fw, cur.off.z) - 0.5f, fmaf(cur.rf.y, fh, cur.off.w) - 0.5f,
                                (valid && has) ? e1 * inv : 0.f, s1);
            if (!valid || !has) s0.gbase = s1.gbase = -1;
            // compaction slots of the fallback samples inside the item
            const unsigned m0 = __ballot_sync(0xffffffffu, s0.gbase >= 0), m1 = __ballot_sync(0xffffffffu, s1.gbase >= 0);
            const unsigned gm0 = (m0 >> (8 * g)) & 0xffu, gm1 = (m1 >> (8 * g)) & 0xffu;
            const int nfb = __popc(gm0) + __popc(gm1);
            __syncwarp();      // the previous pass has finished reading the staging area
            if (has) {
                // slot t of the record array holds sample 2t (t < 6) or 2(t - 6) + 1: the lanes of a group store
                // consecutive 16-byte records (the sample-major order cost 8 wavefronts per store: 2-way conflicts)
                st_w[g * LPAD + j] = make_float4(s0.w[0], s0.w[1], s0.w[2], s0.w[3]);
                st_w[g * LPAD + 6 + j] = make_float4(s1.w[0], s1.w[1], s1.w[2], s1.w[3]);
                st_b[g * LP + j] = s0.base;
                st_b[g * LP + 6 + j] = s1.base;
            }
            const unsigned below = (1u << j) - 1u;
            int k = __popc(gm0 & below) + __popc(gm1 & below);
            if (s0.gbase >= 0) {
                fb_w[g * LP + k] = make_float4(s0.gw[0], s0.gw[1], s0.gw[2], s0.gw[3]);
                fb_b[g * LP + k] = s0.gbase * 4 + l;     // low bits: level of the sample
                ++k;
            }
            if (s1.gbase >= 0) {
                fb_w[g * LP + k] = make_float4(s1.gw[0], s1.gw[1], s1.gw[2], s1.gw[3]);
                fb_b[g * LP + k] = s1.gbase * 4 + l;
            }
            __syncwarp();
            if (!waited) {     // the first pass's preparation overlapped the window fill
                mbar_wait(bar, phase);
                phase ^= 1;
                waited = true;
            }
            // ---- shared-memory pass: 12 samples x 4 corners
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            const int4* bq = reinterpret_cast<const int4*>(st_b + g * LP);
            const int4 b0 = bq[0], b1 = bq[1], b2 = bq[2];
            const int bases[LP] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w};
#pragma unroll
            for (int s = 0; s < LP; ++s) {
                const int l = (s < 6 ? 2 * s : 2 * (s - 6) + 1) / NP;     // level of the sample in slot s
                const int rowstride = (l == 2 ? W::wx(2) : (l == 1 ? W::wx(1) : W::wx(0))) * 8;
                const float4 w = st_w[g * LPAD + s];
                const float4* p = win + bases[s] + j;
                const float4 v00 = p[0], v01 = p[8], v10 = p[rowstride], v11 = p[rowstride + 8];
                fma4(acc, w.x, v00); fma4(acc, w.y, v01); fma4(acc, w.z, v10); fma4(acc, w.w, v11);
            }
            // ---- samples whose footprint left the window: compacted, straight from global memory
            const int nmax = max(max(__shfl_sync(0xffffffffu, nfb, 0), __shfl_sync(0xffffffffu, nfb, 8)),
                                 max(__shfl_sync(0xffffffffu, nfb, 16), __shfl_sync(0xffffffffu, nfb, 24)));
            for (int k = 0; k < nmax; ++k) {
                if (k < nfb) {
                    const float4 w = fb_w[g * LP + k];
                    const int code = fb_b[g * LP + k];
                    const int pix = code >> 2;
                    const int rowstride = lv.w[code & 3] * (NH * 8);
                    const float4* p = vb + (int64_t)pix * (NH * 8);
                    const float4 v00 = __ldg(p), v01 = __ldg(p + NH * 8), v10 = __ldg(p + rowstride),
                                 v11 = __ldg(p + rowstride + NH * 8);
                    fma4(acc, w.x, v00); fma4(acc, w.y, v01); fma4(acc, w.z, v10); fma4(acc, w.w, v11);
                }
            }
            if (valid) {
                const int64_t oi = (((int64_t)b * N + nq) * NH + head) * 8 + j;
                if (out) reinterpret_cast<float4*>(out)[oi] = acc;
                if (out_hi) {   // operand planes for the output projection GEMM
                    const float a4[4] = {acc.x, acc.y, acc.z, acc.w};
                    uint32_t hh[4], ll[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const uint32_t u = __float_as_uint(a4[e]);
                        const uint32_t hb = (u + 0x7fffu + ((u >> 16) & 1u)) & 0xffff0000u;   // RNE bf16
                        const float rem = a4[e] - __uint_as_float(hb);
                        const uint32_t ur = __float_as_uint(rem);
                        hh[e] = hb >> 16;
                        ll[e] = ((ur + 0x7fffu + ((ur >> 16) & 1u)) >> 16) & 0xffffu;
                    }
                    out_hi[oi] = make_uint2(hh[0] | (hh[1] << 16), hh[2] | (hh[3] << 16));
                    out_lo[oi] = make_uint2(ll[0] | (ll[1] << 16), ll[2] | (ll[3] << 16));
                }
            }
        }
        __syncthreads();   // all gathers of this region are done before its windows are overwritten
    }
}

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

// 5-D map over one level of value [B, N, 8 heads, 32] fp32: (channel, head, x, y, frame); box = one head's window
bool make_level_map(CUtensorMap* m, const float* value, int B, int64_t N, int start, int h, int w, int wx, int wy) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[5] = {32, (cuuint64_t)NH, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)B};
    cuuint64_t strides[4] = {128, 128 * NH, (cuuint64_t)w * 128 * NH, (cuuint64_t)N * 128 * NH};
    cuuint32_t box[5] = {32, 1, (cuuint32_t)wx, (cuuint32_t)wy, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    void* base = const_cast<float*>(value + (int64_t)start * NH * 32);
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

constexpr int kHX = 4, kHY = 3;

}  // namespace

int pvsg_internal::configure_msda_tile() {
    static bool configured[PVSG_MAX_DEVICES];
    if (pvsg_first_use_on_device(configured) &&
        (cudaFuncSetAttribute(msda_tile_kernel<kHX, kHY>, cudaFuncAttributeMaxDynamicSharedMemorySize, Win<kHX, kHY>::SMEM) != cudaSuccess ||
         cudaFuncSetAttribute(msda_tile_kernel<kHX, kHY>, cudaFuncAttributePreferredSharedMemoryCarveout,
                              cudaSharedmemCarveoutMaxShared) != cudaSuccess))
        return PVSG_ERR_LAUNCH;
    return PVSG_OK;
}

// Region-tiled fused MSDeformAttn (queries = pyramid tokens, 3 levels x 4 points, 8 heads x 32 channels, levels
// forming an exact 2x pyramid).  Returns PVSG_ERR_UNSUPPORTED when the shape does not qualify (the caller then
// uses the lane-group kernel of msda.cu).
int pvsg_msda_tile_launch(const float* value, const int* hs, const int* ws, const int* starts, const float* proj,
                          const float* ref, float* out, void* out_hi, void* out_lo, int B, int64_t N, void* stream) {
    if (hs[1] != 2 * hs[0] || hs[2] != 2 * hs[1] || ws[1] != 2 * ws[0] || ws[2] != 2 * ws[1]) return PVSG_ERR_UNSUPPORTED;
    if (hs[0] < 2 || ws[0] < 2 || N * NH * 8 >= (1LL << 30) || (reinterpret_cast<uintptr_t>(value) & 15)) return PVSG_ERR_UNSUPPORTED;
    using W = Win<kHX, kHY>;
    CUtensorMap maps[3];
    for (int l = 0; l < 3; ++l)
        if (!make_level_map(&maps[l], value, B, N, starts[l], hs[l], ws[l], W::wx(l), W::wy(l))) return PVSG_ERR_UNSUPPORTED;
    if (const int rc = pvsg_internal::configure_msda_tile()) return rc;
    TileLevels lv;
    for (int l = 0; l < 3; ++l) { lv.h[l] = hs[l]; lv.w[l] = ws[l]; lv.start[l] = starts[l]; }
    const int regions_x = (ws[2] + RW - 1) / RW, regions_y = (hs[2] + RH - 1) / RH;
    const int64_t total = (int64_t)B * regions_x * regions_y * NH;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        sms = 148;
    const unsigned grid = (unsigned)imin64(total, 2 * (int64_t)sms);
    msda_tile_kernel<kHX, kHY><<<grid, NTHR, W::SMEM, as_stream(stream)>>>(
        maps[0], maps[1], maps[2], value, lv, proj, ref, out, reinterpret_cast<uint2*>(out_hi),
        reinterpret_cast<uint2*>(out_lo), B, (int)N, regions_x, regions_y);
    return pvsg_launch_status();
}
