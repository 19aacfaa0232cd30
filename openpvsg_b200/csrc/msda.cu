// Multi-scale deformable attention (sampling + aggregation), D = 32 channels per head.
//
// One warp owns one (query, head).  The warp is split in 4 groups of 8 lanes; a group
// covers the 32 channels with one float4 per lane (one 128-byte line per bilinear corner)
// and walks the samples s = g, g+4, g+8, ... of the L*P samples; the four partial sums
// are combined with two warp shuffles.  In the fused variant the warp also does the
// softmax over the L*P attention logits and the location arithmetic from the raw
// projections, so sampling_locations / attention_weights never exist in memory.
#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int MAX_LEVELS = 8;

struct MsdaLevels {
    int h[MAX_LEVELS];
    int w[MAX_LEVELS];
    int64_t start[MAX_LEVELS];
    int64_t tile_start[MAX_LEVELS + 1];   // prefix sums of ceil(h/8) * ceil(w/8)
};

// Bilinear sample of value[b, start + (y, x), head, 4*q4 .. 4*q4+3] with zero padding,
// pixel coordinates (x, y) already in the align_corners=False frame (loc * size - 0.5).
__device__ __forceinline__ void sample_acc(const float4* __restrict__ vbase, int hgt, int wid, int pix_stride,
                                           float x, float y, float aw, float4& acc) {
    // Branch-free: corner indices are clamped into the map and out-of-range corners get weight 0
    // (identical to zero padding, and to mmcv's "skip samples outside (-1, size)" rule), so the four
    // loads of every sample are unconditional and the compiler can keep all of them in flight.
    x = fminf(fmaxf(x, -2.f), (float)wid + 1.f);
    y = fminf(fmaxf(y, -2.f), (float)hgt + 1.f);
    const float fy = floorf(y), fx = floorf(x);
    const int y0 = (int)fy, x0 = (int)fx;
    const float ly = y - fy, lx = x - fx;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const bool y0v = y0 >= 0 && y0 < hgt, y1v = y0 + 1 >= 0 && y0 + 1 < hgt;
    const bool x0v = x0 >= 0 && x0 < wid, x1v = x0 + 1 >= 0 && x0 + 1 < wid;
    const int yc0 = min(max(y0, 0), hgt - 1), yc1 = min(max(y0 + 1, 0), hgt - 1);
    const int xc0 = min(max(x0, 0), wid - 1), xc1 = min(max(x0 + 1, 0), wid - 1);
    // 32-bit indices: N * H * 8 float4 < 2^31 (checked on the host)
    const float4 v00 = __ldg(vbase + (yc0 * wid + xc0) * pix_stride);
    const float4 v01 = __ldg(vbase + (yc0 * wid + xc1) * pix_stride);
    const float4 v10 = __ldg(vbase + (yc1 * wid + xc0) * pix_stride);
    const float4 v11 = __ldg(vbase + (yc1 * wid + xc1) * pix_stride);
    const float w00 = (y0v && x0v) ? hy * hx * aw : 0.f, w01 = (y0v && x1v) ? hy * lx * aw : 0.f;
    const float w10 = (y1v && x0v) ? ly * hx * aw : 0.f, w11 = (y1v && x1v) ? ly * lx * aw : 0.f;
    acc.x += w00 * v00.x + w01 * v01.x + w10 * v10.x + w11 * v11.x;
    acc.y += w00 * v00.y + w01 * v01.y + w10 * v10.y + w11 * v11.y;
    acc.z += w00 * v00.z + w01 * v01.z + w10 * v10.z + w11 * v11.z;
    acc.w += w00 * v00.w + w01 * v01.w + w10 * v10.w + w11 * v11.w;
}

// Per-sample part of the bilinear lookup, done ONCE by the lane that owns the sample: clamped
// corner pixel indices (relative to the value tensor, level start included) and the four corner
// weights (attention weight folded in, 0 for corners outside the map).
__device__ __forceinline__ void sample_prep(int hgt, int wid, int start, float x, float y, float aw, int (&idx)[4],
                                            float (&w)[4]) {
    x = fminf(fmaxf(x, -2.f), (float)wid + 1.f);
    y = fminf(fmaxf(y, -2.f), (float)hgt + 1.f);
    const float fy = floorf(y), fx = floorf(x);
    const int y0 = (int)fy, x0 = (int)fx;
    const float ly = y - fy, lx = x - fx;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const bool y0v = y0 >= 0 && y0 < hgt, y1v = y0 + 1 >= 0 && y0 + 1 < hgt;
    const bool x0v = x0 >= 0 && x0 < wid, x1v = x0 + 1 >= 0 && x0 + 1 < wid;
    const int yc0 = min(max(y0, 0), hgt - 1), yc1 = min(max(y0 + 1, 0), hgt - 1);
    const int xc0 = min(max(x0, 0), wid - 1), xc1 = min(max(x0 + 1, 0), wid - 1);
    idx[0] = start + yc0 * wid + xc0; idx[1] = start + yc0 * wid + xc1;
    idx[2] = start + yc1 * wid + xc0; idx[3] = start + yc1 * wid + xc1;
    w[0] = (y0v && x0v) ? hy * hx * aw : 0.f; w[1] = (y0v && x1v) ? hy * lx * aw : 0.f;
    w[2] = (y1v && x0v) ? ly * hx * aw : 0.f; w[3] = (y1v && x1v) ? ly * lx * aw : 0.f;
}

template <bool FUSED, bool TILED>
__global__ void __launch_bounds__(256, 4) msda_kernel(const float* __restrict__ value, MsdaLevels lv,
                                                   const float* __restrict__ loc_or_proj,
                                                   const float* __restrict__ aw_or_ref, float* __restrict__ out,
                                                   int64_t N, int64_t Nq, int H, int L, int P, int64_t total) {
    const int lane = threadIdx.x & 31;
    const int g = lane >> 3, q4 = lane & 7;
    const int LP = L * P;
    // work item = (b, query, head); consecutive warps of a CTA take consecutive QUERIES of one
    // head so that their sampling neighbourhoods overlap in L1.
    // TILED (queries are the pyramid tokens themselves): a CTA owns an 8 x 8 tile of queries
    // of one level, warp w walks row w -- the 64 sampling neighbourhoods overlap in 2-D, which is
    // what keeps the bilinear corner lines in L1.
    const int wpb = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5;
    const int64_t qblocks = TILED ? lv.tile_start[L] : (Nq + wpb - 1) / wpb;
    for (int64_t blk = blockIdx.x; blk < total; blk += gridDim.x) {
        const int64_t qb = blk % qblocks;
        const int64_t t = blk / qblocks;
        const int head = (int)(t % H);
        const int64_t b = t / H;
        int tl = 0, ty = 0, tx = 0;
        if (TILED) {
            while (tl + 1 < L && qb >= lv.tile_start[tl + 1]) ++tl;
            const int local = (int)(qb - lv.tile_start[tl]);
            const int tiles_x = (lv.w[tl] + 7) >> 3;
            ty = local / tiles_x;
            tx = local - ty * tiles_x;
        }
      for (int qx = 0; qx < (TILED ? 8 : 1); ++qx) {
        int64_t nq;
        if (TILED) {
            const int y = ty * 8 + warp, x = tx * 8 + qx;
            if (y >= lv.h[tl] || x >= lv.w[tl]) continue;
            nq = lv.start[tl] + (int64_t)y * lv.w[tl] + x;
        } else {
            nq = qb * wpb + warp;
            if (nq >= Nq) continue;
        }
        const int pix_stride = H * 8;  // float4 per pixel (H heads * 32 ch / 4)
        const float4* vb = reinterpret_cast<const float4*>(value) + (b * N * H + head) * 8 + q4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        if (FUSED) {
            // proj row: [H*LP*2 offsets | H*LP logits]
            const float* prow = loc_or_proj + (b * Nq + nq) * (int64_t)(H * LP * 3);
            const float* offp = prow + (int64_t)head * LP * 2;
            const float* lgp = prow + (int64_t)H * LP * 2 + (int64_t)head * LP;
            const float lg = lane < LP ? __ldg(lgp + lane) : -INFINITY;
            const float mx = warp_max(lg);
            const float e = lane < LP ? __expf(lg - mx) : 0.f;
            const float wgt = e / warp_sum(e);
            const float o0 = lane < LP ? __ldg(offp + 2 * lane) : 0.f;      // x offset of sample `lane`
            const float o1 = lane < LP ? __ldg(offp + 2 * lane + 1) : 0.f;  // y offset
            const float rx = __ldg(aw_or_ref + nq * 2), ry = __ldg(aw_or_ref + nq * 2 + 1);
            // lane s owns sample s: location arithmetic once per sample, not once per channel group
            int idx[4];
            float cw[4];
            {
                const int l = min(lane, LP - 1) / P;
                const int hgt = lv.h[l], wid = lv.w[l];
                // loc = ref + off / (w, h); pixel = loc * size - 0.5 = ref * size + off - 0.5
                const float x = fmaf(rx, (float)wid, o0) - 0.5f;
                const float y = fmaf(ry, (float)hgt, o1) - 0.5f;
                sample_prep(hgt, wid, (int)lv.start[l], x, y, wgt, idx, cw);   // wgt = 0 for lane >= LP
            }
#pragma unroll 3
            for (int s0 = 0; s0 < LP; s0 += 4) {  // trip count is warp-uniform (full-mask shuffles)
                const int s = min(s0 + g, 31);    // lanes >= LP carry weight 0 and in-range indices
                float4 v[4];
                float w[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int i = __shfl_sync(0xffffffffu, idx[c], s);
                    w[c] = __shfl_sync(0xffffffffu, cw[c], s);
                    v[c] = __ldg(vb + i * pix_stride);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    acc.x = fmaf(w[c], v[c].x, acc.x); acc.y = fmaf(w[c], v[c].y, acc.y);
                    acc.z = fmaf(w[c], v[c].z, acc.z); acc.w = fmaf(w[c], v[c].w, acc.w);
                }
            }
        } else {
            const float* locp = loc_or_proj + ((b * Nq + nq) * H + head) * (int64_t)LP * 2;
            const float* awp = aw_or_ref + ((b * Nq + nq) * H + head) * (int64_t)LP;
            for (int s = g; s < LP; s += 4) {
                const int l = s / P;
                const int hgt = lv.h[l], wid = lv.w[l];
                const float x = __ldg(locp + 2 * s) * (float)wid - 0.5f;
                const float y = __ldg(locp + 2 * s + 1) * (float)hgt - 0.5f;
                sample_acc(vb + (int)lv.start[l] * pix_stride, hgt, wid, pix_stride, x, y, __ldg(awp + s), acc);
            }
        }
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
            acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
            acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
        }
        if (g == 0) reinterpret_cast<float4*>(out)[((b * Nq + nq) * H + head) * 8 + q4] = acc;
      }
    }
}

// ---------------------------------------------------------------------------------------
// Fused variant, L*P <= 16 (the reference configuration has 3 x 4 = 12 samples).
//
// ncu on the warp-per-item kernel above showed the LSU data pipe as the limiter (77 % busy) and
// most of its wavefronts were not the bilinear gathers but warp shuffles (which travel through
// the same pipe) and register spills.  Here a group of 8 LANES owns one (query, head) item and a
// warp carries four items (4 consecutive queries of one head):
//   * lane j of the group owns samples j and j + 8: softmax = 3-step butterfly inside the group,
//     location arithmetic twice per lane instead of once per item-warp;
//   * per sample the owner broadcasts 5 registers (packed corner index + 4 corner weights) with
//     width-8 shuffles; there is no cross-group reduction at the end.
// Shuffles per item: 12 x 5 + 6 (softmax) over FOUR items per instruction, i.e. ~16 per item
// instead of 42.
// Sample preparation with FIXED corner offsets: the four corners are always the pixels
// (yb, xb), (yb, xb+1), (yb+1, xb), (yb+1, xb+1) with xb = clamp(x0, 0, wid-2), yb likewise, so
// their addresses are base + {0, S, R, R+S} for strides that are uniform per level -- no per-corner
// index arithmetic.  Zero padding and the clamp are folded into the weights: a corner that the
// bilinear footprint does not touch (outside the map, or shifted by the clamp) gets weight 0.
// Needs wid, hgt >= 2.
__device__ __forceinline__ void sample_prep_fixed(int hgt, int wid, int start, float x, float y, float aw,
                                                  int& base, float (&w)[4]) {
    x = fminf(fmaxf(x, -2.f), (float)wid + 1.f);
    y = fminf(fmaxf(y, -2.f), (float)hgt + 1.f);
    const float fy = floorf(y), fx = floorf(x);
    const int y0 = (int)fy, x0 = (int)fx;
    const float ly = y - fy, lx = x - fx;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const int xb = min(max(x0, 0), wid - 2), yb = min(max(y0, 0), hgt - 2);
    const float wxa = (x0 == xb) ? hx : ((x0 + 1 == xb) ? lx : 0.f);
    const float wxb = (x0 == xb) ? lx : ((x0 == xb + 1) ? hx : 0.f);
    const float wya = (y0 == yb) ? hy : ((y0 + 1 == yb) ? ly : 0.f);
    const float wyb = (y0 == yb) ? ly : ((y0 == yb + 1) ? hy : 0.f);
    base = start + yb * wid + xb;
    w[0] = wya * wxa * aw; w[1] = wya * wxb * aw;
    w[2] = wyb * wxa * aw; w[3] = wyb * wxb * aw;
}

// LPC / PC / HC: compile-time L*P, P and head count (12 / 4 / 8 in the reference configuration) so
// that the sample loop is fully unrolled (owner lane, register slot and level of every sample are
// constants) and the pixel stride is an immediate address offset; 0 = runtime.
template <bool TILED, int LPC, int PC, int HC>
__global__ void __launch_bounds__(256, 4) msda_group_kernel(const float* __restrict__ value, MsdaLevels lv,
                                                            const float* __restrict__ proj,
                                                            const float* __restrict__ ref, float* __restrict__ out,
                                                            int64_t N, int64_t Nq, int H, int L, int P_rt,
                                                            int64_t total, uint2* __restrict__ out_hi = nullptr,
                                                            uint2* __restrict__ out_lo = nullptr) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 3, j = lane & 7;
    const int P = PC > 0 ? PC : P_rt;
    const int LP = LPC > 0 ? LPC : L * P;
    const int pix_stride = HC > 0 ? HC * 8 : H * 8;   // float4 per pixel
    // TILED: CTA = 8 x 8 query tile of one level and head, warp = one row, two passes of 4 queries.
    // otherwise: CTA = 32 consecutive queries of one head, warp = 4 of them.
    const int64_t qblocks = TILED ? lv.tile_start[L] : (Nq + 31) / 32;
    for (int64_t blk = blockIdx.x; blk < total; blk += gridDim.x) {
        const int64_t qb = blk % qblocks;
        const int64_t t = blk / qblocks;
        const int head = (int)(t % H);
        const int64_t b = t / H;
        int tl = 0, ty = 0, tx = 0;
        if (TILED) {
            while (tl + 1 < L && qb >= lv.tile_start[tl + 1]) ++tl;
            const int local = (int)(qb - lv.tile_start[tl]);
            const int tiles_x = (lv.w[tl] + 7) >> 3;
            ty = local / tiles_x;
            tx = local - ty * tiles_x;
        }
        const float4* vb = reinterpret_cast<const float4*>(value) + (b * N * H + head) * 8 + j;
        for (int pass = 0; pass < (TILED ? 2 : 1); ++pass) {
            int64_t nq;
            bool valid;
            if (TILED) {
                const int y = ty * 8 + warp, x = tx * 8 + pass * 4 + g;
                if (y >= lv.h[tl] || tx * 8 + pass * 4 >= lv.w[tl]) continue;   // warp-uniform
                valid = x < lv.w[tl];
                nq = lv.start[tl] + (int64_t)y * lv.w[tl] + min(x, lv.w[tl] - 1);
            } else {
                const int64_t q0 = qb * 32 + warp * 4;
                if (q0 >= Nq) continue;                                         // warp-uniform
                valid = q0 + g < Nq;
                nq = min(q0 + g, Nq - 1);
            }
            const float* prow = proj + (b * Nq + nq) * (int64_t)(H * LP * 3);
            const float2* offp = reinterpret_cast<const float2*>(prow + (int64_t)head * LP * 2);
            const float* lgp = prow + (int64_t)H * LP * 2 + (int64_t)head * LP;
            const bool has0 = j < LP, has1 = j + 8 < LP;
            const float lg0 = has0 ? __ldg(lgp + j) : -INFINITY, lg1 = has1 ? __ldg(lgp + j + 8) : -INFINITY;
            float mx = fmaxf(lg0, lg1);
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const float e0 = has0 ? __expf(lg0 - mx) : 0.f, e1 = has1 ? __expf(lg1 - mx) : 0.f;
            float sum = e0 + e1;
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float inv = 1.f / sum;
            const float2 rf = __ldg(reinterpret_cast<const float2*>(ref) + nq);
            int base0, base1;
            float w0[4], w1[4];
            {
                const float2 o = has0 ? __ldg(offp + j) : make_float2(0.f, 0.f);
                const int l = min(j, LP - 1) / P;
                const int hgt = lv.h[l], wid = lv.w[l];
                // loc = ref + off / (w, h); pixel = loc * size - 0.5 = ref * size + off - 0.5
                sample_prep_fixed(hgt, wid, (int)lv.start[l], fmaf(rf.x, (float)wid, o.x) - 0.5f,
                                  fmaf(rf.y, (float)hgt, o.y) - 0.5f, e0 * inv, base0, w0);
            }
            {
                const float2 o = has1 ? __ldg(offp + j + 8) : make_float2(0.f, 0.f);
                const int l = min(j + 8, LP - 1) / P;
                const int hgt = lv.h[l], wid = lv.w[l];
                sample_prep_fixed(hgt, wid, (int)lv.start[l], fmaf(rf.x, (float)wid, o.x) - 0.5f,
                                  fmaf(rf.y, (float)hgt, o.y) - 0.5f, e1 * inv, base1, w1);
            }
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            auto one_sample = [&](int s, int l) {
                const int rowstride = lv.w[l] * pix_stride;      // uniform per level
                const bool second = s >= 8;                      // uniform (compile-time when unrolled)
                const int src = s & 7;
                const int base = __shfl_sync(0xffffffffu, second ? base1 : base0, src, 8);
                float w[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) w[c] = __shfl_sync(0xffffffffu, second ? w1[c] : w0[c], src, 8);
                const float4* p00 = vb + (int64_t)base * pix_stride;
                const float4* p10 = p00 + rowstride;
                const float4 v00 = __ldg(p00);
                const float4 v01 = __ldg(p00 + pix_stride);     // immediate offset when HC is set
                const float4 v10 = __ldg(p10);
                const float4 v11 = __ldg(p10 + pix_stride);
                acc.x = fmaf(w[0], v00.x, acc.x); acc.y = fmaf(w[0], v00.y, acc.y);
                acc.z = fmaf(w[0], v00.z, acc.z); acc.w = fmaf(w[0], v00.w, acc.w);
                acc.x = fmaf(w[1], v01.x, acc.x); acc.y = fmaf(w[1], v01.y, acc.y);
                acc.z = fmaf(w[1], v01.z, acc.z); acc.w = fmaf(w[1], v01.w, acc.w);
                acc.x = fmaf(w[2], v10.x, acc.x); acc.y = fmaf(w[2], v10.y, acc.y);
                acc.z = fmaf(w[2], v10.z, acc.z); acc.w = fmaf(w[2], v10.w, acc.w);
                acc.x = fmaf(w[3], v11.x, acc.x); acc.y = fmaf(w[3], v11.y, acc.y);
                acc.z = fmaf(w[3], v11.z, acc.z); acc.w = fmaf(w[3], v11.w, acc.w);
            };
            if (LPC > 0) {
#pragma unroll
                for (int s = 0; s < (LPC > 0 ? LPC : 1); ++s) one_sample(s, s / (PC > 0 ? PC : 1));
            } else {
                int l = 0, pcount = 0;
#pragma unroll 2
                for (int s = 0; s < LP; ++s) {
                    one_sample(s, l);
                    if (++pcount == P) { pcount = 0; ++l; }
                }
            }
            if (valid) {
                const int64_t oi = ((b * Nq + nq) * H + head) * 8 + j;
                if (out) reinterpret_cast<float4*>(out)[oi] = acc;
                if (out_hi) {   // operand planes for the output projection GEMM
                    const float a4[4] = {acc.x, acc.y, acc.z, acc.w};
                    uint32_t hh[4], ll[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const uint32_t u = __float_as_uint(a4[e]);
                        // round-to-nearest-even bf16 of a4[e], then of the remainder
                        const uint32_t hb = (u + 0x7fffu + ((u >> 16) & 1u)) & 0xffff0000u;
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
    }
}

}  // namespace

// csrc/msda_tile.cu: region-tiled kernel with TMA-staged value windows (queries = pyramid tokens, reference configuration)
int pvsg_msda_tile_launch(const float* value, const int* hs, const int* ws, const int* starts, const float* proj,
                          const float* ref, float* out, void* out_hi, void* out_lo, int B, int64_t N, void* stream);

namespace {

int fill_levels(MsdaLevels& lv, const int64_t* spatial_shapes, const int64_t* level_start_index, int L,
                int64_t N) {
    if (L <= 0 || L > MAX_LEVELS) return PVSG_ERR_UNSUPPORTED;
    int64_t tot = 0;
    for (int l = 0; l < L; ++l) {
        lv.h[l] = (int)spatial_shapes[2 * l];
        lv.w[l] = (int)spatial_shapes[2 * l + 1];
        lv.start[l] = level_start_index[l];
        if (lv.h[l] <= 0 || lv.w[l] <= 0 || lv.start[l] != tot) return PVSG_ERR_INVALID_ARG;
        tot += (int64_t)lv.h[l] * lv.w[l];
        lv.tile_start[l + 1] = lv.tile_start[l] + (int64_t)((lv.h[l] + 7) / 8) * ((lv.w[l] + 7) / 8);
    }
    return tot == N ? PVSG_OK : PVSG_ERR_INVALID_ARG;
}

template <bool FUSED>
int launch(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
           const float* a, const float* b2, float* out, int B, int64_t N, int64_t Nq, int H, int D, int L,
           int P, void* stream, void* out_hi = nullptr, void* out_lo = nullptr) {
    PVSG_CHECK_ARG(value && spatial_shapes && level_start_index && a && b2 && (out || out_hi));
    PVSG_CHECK_ARG((out_hi == nullptr) == (out_lo == nullptr));
    PVSG_CHECK_ARG(B > 0 && N > 0 && Nq > 0 && H > 0 && P > 0);
    if (D != 32 || L * P > 32 || N * H * 8 >= (1LL << 31)) return PVSG_ERR_UNSUPPORTED;
    MsdaLevels lv{};
    int rc = fill_levels(lv, spatial_shapes, level_start_index, L, N);
    if (rc != PVSG_OK) return rc;
    const int wpb = 8;
    bool wide = true;   // the fixed-corner kernel needs every level to be at least 2 x 2
    for (int l = 0; l < L; ++l) wide = wide && lv.h[l] >= 2 && lv.w[l] >= 2;
    if (FUSED && L * P <= 16 && N < (1LL << 30) && wide) {
        uint2* oh = reinterpret_cast<uint2*>(out_hi);
        uint2* ol = reinterpret_cast<uint2*>(out_lo);
        const bool ref_cfg = L == 3 && P == 4 && H == 8;     // fully unrolled instance
        const char* impl = getenv("PVSG_MSDA_IMPL");      // debug switch: "group" forces the lane-group kernel
        const bool force_group = impl != nullptr && impl[0] == 'g';
        if (ref_cfg && Nq == N && !force_group) {
            // the encoder's case: TMA-staged windows in shared memory (msda_tile.cu); other shapes use the lane-group kernel
            const int hs[3] = {lv.h[0], lv.h[1], lv.h[2]}, ws[3] = {lv.w[0], lv.w[1], lv.w[2]};
            const int starts[3] = {(int)lv.start[0], (int)lv.start[1], (int)lv.start[2]};
            rc = pvsg_msda_tile_launch(value, hs, ws, starts, a, b2, out, out_hi, out_lo, B, N, stream);
            if (rc != PVSG_ERR_UNSUPPORTED) return rc;
        }
        if (Nq == N) {
            const int64_t total = (int64_t)B * H * lv.tile_start[L];
            const unsigned grid = (unsigned)imin64(total, 148 * 64);
            if (ref_cfg)
                msda_group_kernel<true, 12, 4, 8><<<grid, 256, 0, as_stream(stream)>>>(value, lv, a, b2, out, N, Nq, H, L, P,
                                                                                  total, oh, ol);
            else
                msda_group_kernel<true, 0, 0, 0><<<grid, 256, 0, as_stream(stream)>>>(value, lv, a, b2, out, N, Nq, H, L, P,
                                                                                 total, oh, ol);
        } else {
            const int64_t total = (int64_t)B * H * ((Nq + 31) / 32);
            const unsigned grid = (unsigned)imin64(total, 148 * 64);
            if (ref_cfg)
                msda_group_kernel<false, 12, 4, 8><<<grid, 256, 0, as_stream(stream)>>>(value, lv, a, b2, out, N, Nq, H, L, P,
                                                                                   total, oh, ol);
            else
                msda_group_kernel<false, 0, 0, 0><<<grid, 256, 0, as_stream(stream)>>>(value, lv, a, b2, out, N, Nq, H, L, P,
                                                                                  total, oh, ol);
        }
        return pvsg_launch_status();
    }
    if (out_hi || !out) return PVSG_ERR_UNSUPPORTED;   // planes only from the group kernel
    if (Nq == N) {   // queries = pyramid tokens: 2-D tiled query order
        const int64_t total = (int64_t)B * H * lv.tile_start[L];
        const unsigned grid = (unsigned)imin64(total, 148 * 64);
        msda_kernel<FUSED, true><<<grid, wpb * 32, 0, as_stream(stream)>>>(value, lv, a, b2, out, N, Nq, H, L, P, total);
    } else {
        const int64_t total = (int64_t)B * H * ((Nq + wpb - 1) / wpb);
        const unsigned grid = (unsigned)imin64(total, 148 * 64);
        msda_kernel<FUSED, false><<<grid, wpb * 32, 0, as_stream(stream)>>>(value, lv, a, b2, out, N, Nq, H, L, P, total);
    }
    return pvsg_launch_status();
}

}  // namespace

extern "C" int pvsg_msda_forward(const float* value, const int64_t* spatial_shapes,
                                 const int64_t* level_start_index, const float* sampling_locations,
                                 const float* attention_weights, float* out, int B, int64_t N, int64_t Nq,
                                 int H, int D, int L, int P, void* stream) {
    return launch<false>(value, spatial_shapes, level_start_index, sampling_locations, attention_weights,
                         out, B, N, Nq, H, D, L, P, stream);
}

extern "C" int pvsg_msda_fused_forward(const float* value, const int64_t* spatial_shapes,
                                       const int64_t* level_start_index, const float* proj,
                                       const float* ref, float* out, int B, int64_t N, int64_t Nq, int H,
                                       int D, int L, int P, void* stream) {
    return launch<true>(value, spatial_shapes, level_start_index, proj, ref, out, B, N, Nq, H, D, L, P, stream);
}

extern "C" int pvsg_msda_fused_forward_split(const float* value, const int64_t* spatial_shapes,
                                             const int64_t* level_start_index, const float* proj,
                                             const float* ref, float* out, void* out_hi, void* out_lo, int B,
                                             int64_t N, int64_t Nq, int H, int D, int L, int P, void* stream) {
    PVSG_CHECK_ARG(out_hi && out_lo);
    return launch<true>(value, spatial_shapes, level_start_index, proj, ref, out, B, N, Nq, H, D, L, P, stream,
                        out_hi, out_lo);
}
