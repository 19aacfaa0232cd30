// Fused panoptic / instance post-processing (MaskFormerFusionHeadCustom, reference
// models/mask2former/mask2former_fusion_head.py:96-242, :325-404) working directly on the
// low-resolution mask logits: the x4 bilinear upsample, crop, optional rescale, sigmoid,
// score-weighted argmax and the per-segment area tests all happen in registers.
#include "common.cuh"

namespace {

constexpr int NONE = 0x7fff;
constexpr int INS_STRIPS = 8;   // 256-pixel strips per CTA in the instance statistics kernel

struct UpGeom {
    int h, w;            // low-res logits
    int in_h, in_w;      // batch_input_shape (upsample target)
    int img_h, img_w;    // crop
    int out_h, out_w;    // final size
    float sh, sw;        // h / in_h, w / in_w
    float rh, rw;        // img_h / out_h, img_w / out_w (rescale stage)
    int rescale;         // out != img
};

// logits upsampled to (in_h, in_w) at integer pixel (y, x): same operation order as ATen's
// upsample_bilinear2d: (a*wx0 + b*wx1)*wy0 + (c*wx0 + d*wx1)*wy1.
__device__ __forceinline__ float up_logit(const float* __restrict__ lg, const UpGeom& g, int y, int x) {
    int y0, y1, x0, x1;
    float wy0, wy1, wx0, wx1;
    bilinear_coord(y, g.sh, g.h, y0, y1, wy0, wy1);
    bilinear_coord(x, g.sw, g.w, x0, x1, wx0, wx1);
    const float a = __ldg(lg + y0 * g.w + x0), b = __ldg(lg + y0 * g.w + x1);
    const float c = __ldg(lg + y1 * g.w + x0), d = __ldg(lg + y1 * g.w + x1);
    return (a * wx0 + b * wx1) * wy0 + (c * wx0 + d * wx1) * wy1;
}

// value of the final (cropped, optionally rescaled) logit map at output pixel (oy, ox)
__device__ __forceinline__ float final_logit(const float* __restrict__ lg, const UpGeom& g, int oy, int ox) {
    if (!g.rescale) return up_logit(lg, g, oy, ox);
    int y0, y1, x0, x1;
    float wy0, wy1, wx0, wx1;
    bilinear_coord(oy, g.rh, g.img_h, y0, y1, wy0, wy1);
    bilinear_coord(ox, g.rw, g.img_w, x0, x1, wx0, wx1);
    const float a = up_logit(lg, g, y0, x0), b = up_logit(lg, g, y0, x1);
    const float c = up_logit(lg, g, y1, x0), d = up_logit(lg, g, y1, x1);
    return (a * wx0 + b * wx1) * wy0 + (c * wx0 + d * wx1) * wy1;
}

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

// Step 1 (one CTA, one warp per query): softmax max / argmax over the classes, keep test,
// compaction in query order.
__global__ void __launch_bounds__(1024) pan_select_kernel(const float* __restrict__ cls, int Q, int NC, float thr,
                                                          int32_t* __restrict__ seg_info, int32_t* __restrict__ work,
                                                          float* __restrict__ scores) {
    __shared__ int keep[1024];
    __shared__ float sc[1024];
    __shared__ int lb[1024];
    {   // one CTA per frame of the batch
        const int bz = blockIdx.x;
        cls += (int64_t)bz * Q * (NC + 1); seg_info += (int64_t)bz * (1 + 4 * Q);
        work += (int64_t)bz * 4 * Q; scores += (int64_t)bz * Q;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int i = threadIdx.x; i < 4 * Q; i += blockDim.x) work[i] = 0;
    for (int q = warp; q < Q; q += nwarp) {
        const float* row = cls + (int64_t)q * (NC + 1);
        float mx = -INFINITY;
        int arg = 0x7fffffff;
        for (int c = lane; c <= NC; c += 32) {
            const float v = __ldg(row + c);
            if (v > mx) { mx = v; arg = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {   // first maximum wins (torch.max)
            const float om = __shfl_xor_sync(0xffffffffu, mx, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
        }
        float sum = 0.f;
        for (int c = lane; c <= NC; c += 32) sum += expf(__ldg(row + c) - mx);
        sum = warp_sum(sum);
        if (lane == 0) {
            const float score = 1.f / sum;  // softmax value of the arg-max class
            sc[q] = score;
            lb[q] = arg;
            keep[q] = (arg != NC && score > thr) ? 1 : 0;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = 0;
        for (int q = 0; q < Q; ++q) {
            if (keep[q]) {
                seg_info[1 + 4 * n + 0] = q;
                seg_info[1 + 4 * n + 1] = lb[q];
                seg_info[1 + 4 * n + 2] = -1;
                seg_info[1 + 4 * n + 3] = 0;
                scores[n] = sc[q];
                ++n;
            }
        }
        seg_info[0] = n;
    }
}

// Step 2: one thread per output pixel.  work[0..Q) = mask_area (argmax == k), work[Q..2Q) =
// original_area (sigmoid >= 0.5), work[2Q..3Q) = final area (argmax == k and sigmoid >= 0.5).
__global__ void __launch_bounds__(256) pan_pixel_kernel(const float* __restrict__ mask_logits, UpGeom g, int Q,
                                                        const int32_t* __restrict__ seg_info,
                                                        const float* __restrict__ scores, int32_t* __restrict__ work,
                                                        uint16_t* __restrict__ pix) {
    extern __shared__ int sh_cnt[];  // [3 * n_kept]
    const int64_t npix = (int64_t)g.out_h * g.out_w;
    {   // blockIdx.y = frame of the batch
        const int bz = blockIdx.y;
        mask_logits += (int64_t)bz * Q * g.h * g.w; seg_info += (int64_t)bz * (1 + 4 * Q);
        scores += (int64_t)bz * Q; work += (int64_t)bz * 4 * Q; pix += (int64_t)bz * npix;
    }
    const int n = seg_info[0];
    for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) sh_cnt[i] = 0;
    __syncthreads();
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = p < npix;
    const int oy = valid ? (int)(p / g.out_w) : 0, ox = valid ? (int)(p % g.out_w) : 0;
    const int lane = threadIdx.x & 31;
    float best = -INFINITY;
    int bestk = NONE;
    bool best_hi = false;
    const int64_t lstride = (int64_t)g.h * g.w;
    for (int k = 0; k < n; ++k) {
        const int q = seg_info[1 + 4 * k];
        float sig = 0.f;
        bool hi = false;
        if (valid) {
            sig = sigmoidf_(final_logit(mask_logits + q * lstride, g, oy, ox));
            hi = sig >= 0.5f;
            const float prob = scores[k] * sig;
            if (prob > best) { best = prob; bestk = k; best_hi = hi; }  // first max wins (torch argmax)
        }
        const unsigned bal = __ballot_sync(0xffffffffu, hi);
        if (lane == 0 && bal) atomicAdd(&sh_cnt[n + k], __popc(bal));
    }
    // histogram of winners, warp-aggregated
    {
        const unsigned peers = __match_any_sync(0xffffffffu, bestk);
        const unsigned hi_peers = __ballot_sync(0xffffffffu, best_hi) & peers;
        if (valid && bestk != NONE && lane == (__ffs(peers) - 1)) {
            atomicAdd(&sh_cnt[bestk], __popc(peers));
            if (hi_peers) atomicAdd(&sh_cnt[2 * n + bestk], __popc(hi_peers));
        }
    }
    if (valid) pix[p] = (uint16_t)(bestk | (best_hi ? 0x8000 : 0));
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) {
        const int v = sh_cnt[i];
        if (v) atomicAdd(&work[(i / n) * Q + (i % n)], v);
    }
}

// Step 3 (one thread): the sequential accept / id-assignment loop,
// mask2former_fusion_head.py:140-169.
__global__ void pan_decide_kernel(int Q, int num_things, double iou_thr, int filter_low_score, int instance_offset,
                                  int32_t* __restrict__ seg_info, const int32_t* __restrict__ work) {
    if (threadIdx.x != 0) return;
    seg_info += (int64_t)blockIdx.x * (1 + 4 * Q);   // one CTA per frame of the batch
    work += (int64_t)blockIdx.x * 4 * Q;
    const int n = seg_info[0];
    int instance_id = 1;
    for (int k = 0; k < n; ++k) {
        const int cls = seg_info[1 + 4 * k + 1];
        const int mask_area = work[k], original_area = work[Q + k];
        const int final_area = filter_low_score ? work[2 * Q + k] : mask_area;
        int seg = -1;
        if (mask_area > 0 && original_area > 0) {
            if (!((double)mask_area / (double)original_area < iou_thr) && final_area > 0) {
                if (cls < num_things) {
                    seg = cls + instance_id * instance_offset;
                    ++instance_id;
                } else {
                    seg = cls;
                }
            }
        }
        seg_info[1 + 4 * k + 2] = seg;
        seg_info[1 + 4 * k + 3] = seg >= 0 ? final_area : 0;
    }
}

// Step 4: winner index -> segment id.
__global__ void __launch_bounds__(256) pan_write_kernel(const uint16_t* __restrict__ pix,
                                                        const int32_t* __restrict__ seg_info, int NC,
                                                        int filter_low_score, int32_t* __restrict__ pan, int64_t npix,
                                                        int Q) {
    pix += (int64_t)blockIdx.y * npix; pan += (int64_t)blockIdx.y * npix;   // blockIdx.y = frame
    seg_info += (int64_t)blockIdx.y * (1 + 4 * Q);
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix;
         p += (int64_t)gridDim.x * blockDim.x) {
        const int v = pix[p];
        const int k = v & NONE;
        int id = NC;
        if (k != NONE && (!filter_low_score || (v & 0x8000))) {
            const int seg = seg_info[1 + 4 * k + 2];
            if (seg >= 0) id = seg;
        }
        pan[p] = id;
    }
}

__global__ void ins_init_kernel(float* stats, int32_t* boxes, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        stats[2 * i] = 0.f;
        stats[2 * i + 1] = 0.f;
        boxes[4 * i] = INT_MAX;
        boxes[4 * i + 1] = INT_MAX;
        boxes[4 * i + 2] = -1;
        boxes[4 * i + 3] = -1;
    }
}

__global__ void __launch_bounds__(256) ins_pixel_kernel(const float* __restrict__ mask_logits,
                                                        const int32_t* __restrict__ query_idx, UpGeom g,
                                                        float* __restrict__ stats, int32_t* __restrict__ boxes,
                                                        uint8_t* __restrict__ masks_out, int Q, int n) {
    // one CTA = INS_STRIPS consecutive strips of 256 pixels of one candidate; statistics are
    // reduced in registers -> warp shuffles -> shared memory, then ONE set of atomics per CTA
    const int i = blockIdx.y;          // candidate i of frame i / n
    const int q = query_idx[i];
    const int64_t npix = (int64_t)g.out_h * g.out_w;
    const float* lg = mask_logits + ((int64_t)(i / n) * Q + q) * g.h * g.w;
    float s = 0.f;
    int cnt = 0, xmin = INT_MAX, ymin = INT_MAX, xmax = -1, ymax = -1;
    for (int st = 0; st < INS_STRIPS; ++st) {
        const int64_t p = ((int64_t)blockIdx.x * INS_STRIPS + st) * blockDim.x + threadIdx.x;
        if (p >= npix) break;
        const int oy = (int)(p / g.out_w), ox = (int)(p % g.out_w);
        const float v = final_logit(lg, g, oy, ox);
        const bool on = v > 0.f;
        if (on) {
            s += sigmoidf_(v);
            cnt += 1;
            xmin = min(xmin, ox); xmax = max(xmax, ox);
            ymin = min(ymin, oy); ymax = max(ymax, oy);
        }
        if (masks_out) masks_out[(int64_t)i * npix + p] = on ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    __shared__ float sh_s[8];
    __shared__ int sh_i[8][5];
    const int wid = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        sh_s[wid] = s;
        sh_i[wid][0] = cnt; sh_i[wid][1] = xmin; sh_i[wid][2] = ymin; sh_i[wid][3] = xmax; sh_i[wid][4] = ymax;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            s += sh_s[w];
            cnt += sh_i[w][0];
            xmin = min(xmin, sh_i[w][1]); ymin = min(ymin, sh_i[w][2]);
            xmax = max(xmax, sh_i[w][3]); ymax = max(ymax, sh_i[w][4]);
        }
        if (cnt) {
            atomicAdd(&stats[2 * i], s);
            atomicAdd(&stats[2 * i + 1], (float)cnt);
            atomicMin(&boxes[4 * i], xmin);
            atomicMin(&boxes[4 * i + 1], ymin);
            atomicMax(&boxes[4 * i + 2], xmax);
            atomicMax(&boxes[4 * i + 3], ymax);
        }
    }
}

__global__ void ins_finish_kernel(int32_t* boxes, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        if (boxes[4 * i + 2] < 0) {
            boxes[4 * i] = boxes[4 * i + 1] = boxes[4 * i + 2] = boxes[4 * i + 3] = 0;
        } else {
            boxes[4 * i + 2] += 1;
            boxes[4 * i + 3] += 1;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Fast paths for the usual geometry (no rescale stage: out == img, i.e. only upsample + crop).
// A thread owns one output COLUMN and walks ROWS consecutive output rows: the horizontal source
// coordinates are computed once, the vertical ones are warp-uniform, and the horizontally
// interpolated values of the two source rows,  top = a*wx0 + b*wx1,  bot = c*wx0 + d*wx1,  are
// kept in registers while consecutive output rows fall between the same pair of source rows
// (x4 upsampling: 4 loads per 4 rows instead of 16).  v = top*wy0 + bot*wy1 is the same
// operation order as up_logit() above / ATen's upsample_bilinear2d.
__device__ __forceinline__ float fast_sigmoid(float v) { return __fdividef(1.f, 1.f + __expf(-v)); }

constexpr int PAN_ROWS = 4;

__global__ void __launch_bounds__(256) pan_pixel_fast_kernel(const float* __restrict__ mask_logits, UpGeom g, int Q,
                                                             const int32_t* __restrict__ seg_info,
                                                             const float* __restrict__ scores,
                                                             int32_t* __restrict__ work, uint16_t* __restrict__ pix) {
    extern __shared__ int sh_cnt[];  // [3 * n_kept]
    {   // blockIdx.z = frame of the batch
        const int bz = blockIdx.z;
        mask_logits += (int64_t)bz * Q * g.h * g.w; seg_info += (int64_t)bz * (1 + 4 * Q);
        scores += (int64_t)bz * Q; work += (int64_t)bz * 4 * Q; pix += (int64_t)bz * g.out_h * g.out_w;
    }
    const int n = seg_info[0];
    for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) sh_cnt[i] = 0;
    __syncthreads();
    const int ox = blockIdx.x * blockDim.x + threadIdx.x;
    const int oy0 = blockIdx.y * PAN_ROWS;
    const bool xvalid = ox < g.out_w;
    const int lane = threadIdx.x & 31;
    int x0, x1;
    float wx0, wx1;
    bilinear_coord(xvalid ? ox : 0, g.sw, g.w, x0, x1, wx0, wx1);
    int y0[PAN_ROWS], y1[PAN_ROWS];
    float wy0[PAN_ROWS], wy1[PAN_ROWS];
    float best[PAN_ROWS];
    int bestk[PAN_ROWS];
    bool best_hi[PAN_ROWS];
#pragma unroll
    for (int r = 0; r < PAN_ROWS; ++r) {
        bilinear_coord(min(oy0 + r, g.out_h - 1), g.sh, g.h, y0[r], y1[r], wy0[r], wy1[r]);
        best[r] = -INFINITY;
        bestk[r] = NONE;
        best_hi[r] = false;
    }
    const int64_t lstride = (int64_t)g.h * g.w;
    for (int k = 0; k < n; ++k) {
        const float* lg = mask_logits + seg_info[1 + 4 * k] * lstride;
        const float sck = scores[k];
        float top = 0.f, bot = 0.f;
        int cy0 = -1, cy1 = -1, hi_cnt = 0;
#pragma unroll
        for (int r = 0; r < PAN_ROWS; ++r) {
            if (y0[r] != cy0 || y1[r] != cy1) {   // warp-uniform
                cy0 = y0[r]; cy1 = y1[r];
                top = __ldg(lg + cy0 * g.w + x0) * wx0 + __ldg(lg + cy0 * g.w + x1) * wx1;
                bot = __ldg(lg + cy1 * g.w + x0) * wx0 + __ldg(lg + cy1 * g.w + x1) * wx1;
            }
            const bool valid = xvalid && oy0 + r < g.out_h;
            const float sig = fast_sigmoid(top * wy0[r] + bot * wy1[r]);
            const bool hi = valid && sig >= 0.5f;
            const float prob = sck * sig;
            if (valid && prob > best[r]) { best[r] = prob; bestk[r] = k; best_hi[r] = hi; }  // first max wins
            hi_cnt += __popc(__ballot_sync(0xffffffffu, hi));
        }
        if (lane == 0 && hi_cnt) atomicAdd(&sh_cnt[n + k], hi_cnt);
    }
#pragma unroll
    for (int r = 0; r < PAN_ROWS; ++r) {
        const bool valid = xvalid && oy0 + r < g.out_h;
        const int bk = valid ? bestk[r] : NONE;
        const unsigned peers = __match_any_sync(0xffffffffu, bk);
        const unsigned hi_peers = __ballot_sync(0xffffffffu, best_hi[r]) & peers;
        if (bk != NONE && lane == (__ffs(peers) - 1)) {
            atomicAdd(&sh_cnt[bk], __popc(peers));
            if (hi_peers) atomicAdd(&sh_cnt[2 * n + bk], __popc(hi_peers));
        }
        if (valid) pix[(int64_t)(oy0 + r) * g.out_w + ox] = (uint16_t)(bk | (best_hi[r] ? 0x8000 : 0));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) {
        const int v = sh_cnt[i];
        if (v) atomicAdd(&work[(i / n) * Q + (i % n)], v);
    }
}

constexpr int INS_ROWS = 16;

template <bool MASKS>
__global__ void __launch_bounds__(256) ins_pixel_fast_kernel(const float* __restrict__ mask_logits,
                                                             const int32_t* __restrict__ query_idx, UpGeom g,
                                                             float* __restrict__ stats, int32_t* __restrict__ boxes,
                                                             uint8_t* __restrict__ masks_out, int Q, int n) {
    const int i = blockIdx.z;          // candidate i of frame i / n
    const float* lg = mask_logits + ((int64_t)(i / n) * Q + query_idx[i]) * g.h * g.w;
    const int ox = blockIdx.x * blockDim.x + threadIdx.x;
    const int oy0 = blockIdx.y * INS_ROWS;
    // vertical source coordinates of the CTA's rows: computed once, read as broadcasts
    __shared__ int sy0[INS_ROWS], sy1[INS_ROWS];
    __shared__ float swy0[INS_ROWS], swy1[INS_ROWS];
    if (threadIdx.x < INS_ROWS)
        bilinear_coord(min(oy0 + (int)threadIdx.x, g.out_h - 1), g.sh, g.h, sy0[threadIdx.x], sy1[threadIdx.x],
                       swy0[threadIdx.x], swy1[threadIdx.x]);
    __syncthreads();
    const bool xvalid = ox < g.out_w;
    int x0, x1;
    float wx0, wx1;
    bilinear_coord(xvalid ? ox : 0, g.sw, g.w, x0, x1, wx0, wx1);
    float s = 0.f, top = 0.f, bot = 0.f;
    int cnt = 0, ymin = INT_MAX, ymax = -1, cy0 = -1, cy1 = -1;
    const int rows = min(INS_ROWS, g.out_h - oy0);
    for (int r = 0; r < rows; ++r) {
        const int y0 = sy0[r], y1 = sy1[r];
        const float wy0 = swy0[r], wy1 = swy1[r];
        if (y0 != cy0 || y1 != cy1) {   // warp-uniform
            cy0 = y0; cy1 = y1;
            top = __ldg(lg + cy0 * g.w + x0) * wx0 + __ldg(lg + cy0 * g.w + x1) * wx1;
            bot = __ldg(lg + cy1 * g.w + x0) * wx0 + __ldg(lg + cy1 * g.w + x1) * wx1;
        }
        const float v = top * wy0 + bot * wy1;
        const bool on = xvalid && v > 0.f;
        if (on) {
            s += fast_sigmoid(v);
            cnt += 1;
            ymin = min(ymin, oy0 + r);
            ymax = oy0 + r;
        }
        if (MASKS && xvalid) masks_out[((int64_t)i * g.out_h + oy0 + r) * g.out_w + ox] = on ? 1 : 0;
    }
    int xmin = cnt ? ox : INT_MAX, xmax = cnt ? ox : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    }
    __shared__ float sh_s[8];
    __shared__ int sh_i[8][5];
    const int wid = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        sh_s[wid] = s;
        sh_i[wid][0] = cnt; sh_i[wid][1] = xmin; sh_i[wid][2] = ymin; sh_i[wid][3] = xmax; sh_i[wid][4] = ymax;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            s += sh_s[w];
            cnt += sh_i[w][0];
            xmin = min(xmin, sh_i[w][1]); ymin = min(ymin, sh_i[w][2]);
            xmax = max(xmax, sh_i[w][3]); ymax = max(ymax, sh_i[w][4]);
        }
        if (cnt) {
            atomicAdd(&stats[2 * i], s);
            atomicAdd(&stats[2 * i + 1], (float)cnt);
            atomicMin(&boxes[4 * i], xmin);
            atomicMin(&boxes[4 * i + 1], ymin);
            atomicMax(&boxes[4 * i + 2], xmax);
            atomicMax(&boxes[4 * i + 3], ymax);
        }
    }
}

// Instance candidates (mask2former_fusion_head.py:214-222): softmax over the classes, drop the
// void column, top-k over the flattened [Q * NC] scores.  One CTA: scores in shared memory, the
// k-th largest value found by a 4 x 8-bit radix select on the float bit patterns (scores > 0, so
// the unsigned order is the float order), then compaction in index order (torch.topk with
// sorted=False leaves the order unspecified; ties at the threshold go to the lowest indices).
__global__ void __launch_bounds__(1024) ins_select_kernel(const float* __restrict__ cls, int Q, int NC, int k,
                                                          float* __restrict__ top_scores,
                                                          int32_t* __restrict__ top_labels,
                                                          int32_t* __restrict__ top_query) {
    extern __shared__ float sc[];            // [Q * NC]
    __shared__ unsigned hist[256];
    __shared__ unsigned sel_prefix, sel_remaining, out_gt, out_eq;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int total = Q * NC;
    {   // one CTA per frame of the batch
        const int bz = blockIdx.x;
        cls += (int64_t)bz * Q * (NC + 1);
        top_scores += (int64_t)bz * k; top_labels += (int64_t)bz * k; top_query += (int64_t)bz * k;
    }
    for (int q = warp; q < Q; q += nwarp) {
        const float* row = cls + (int64_t)q * (NC + 1);
        float mx = -INFINITY;
        for (int c = lane; c <= NC; c += 32) mx = fmaxf(mx, __ldg(row + c));
        mx = warp_max(mx);
        float sum = 0.f;
        for (int c = lane; c <= NC; c += 32) sum += expf(__ldg(row + c) - mx);
        sum = warp_sum(sum);
        for (int c = lane; c < NC; c += 32) sc[q * NC + c] = expf(__ldg(row + c) - mx) / sum;
    }
    if (threadIdx.x == 0) { sel_prefix = 0; sel_remaining = (unsigned)k; out_gt = 0; out_eq = 0; }
    __syncthreads();
    // radix select, most significant byte first
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        const unsigned prefix = sel_prefix;
        const unsigned pmask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const unsigned u = __float_as_uint(sc[i]);
            const bool in = (u & pmask) == prefix;
            // softmax scores share their leading bytes: aggregate equal bins inside the warp first
            const unsigned bin = in ? ((u >> shift) & 255u) : 256u;
            const unsigned peers = __match_any_sync(__activemask(), bin);
            if (in && lane == (__ffs(peers) - 1)) atomicAdd(&hist[bin], (unsigned)__popc(peers));
        }
        __syncthreads();
        if (warp == 0) {
            // bin holding the rem-th largest element: lane L owns bins 255 - 8L .. 248 - 8L (descending),
            // exclusive prefix over lanes by shuffles, then a scan of the owning lane's 8 bins
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
            const bool owner = excl < rem0 && rem0 <= incl;       // exactly one lane (total >= rem0)
            const unsigned ball = __ballot_sync(0xffffffffu, owner);
            if (ball == 0) {            // cannot happen (k <= total); keep the state consistent
                if (lane == 0) { sel_prefix = prefix; sel_remaining = rem0; }
            } else if (owner) {
                unsigned rem = rem0 - excl;
                int b = 255 - 8 * lane;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    if (mine[e] >= rem) { b = 255 - 8 * lane - e; break; }
                    rem -= mine[e];
                }
                sel_prefix = prefix | ((unsigned)b << shift);
                sel_remaining = rem;   // how many elements equal to the final threshold are taken
            }
        }
        __syncthreads();
    }
    const unsigned thr = sel_prefix;
    const unsigned n_eq = sel_remaining;
    // compaction in index order: warp-strided chunks keep the order deterministic
    // (chunk c = indices [c*1024, c*1024+1024) handled by all threads, prefix by ballot + smem base)
    __shared__ unsigned wcount_gt[32], wcount_eq[32];
    for (int base = 0; base < total; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const unsigned u = i < total ? __float_as_uint(sc[i]) : 0u;
        const bool gt = i < total && u > thr, eq = i < total && u == thr;
        const unsigned bg = __ballot_sync(0xffffffffu, gt), be = __ballot_sync(0xffffffffu, eq);
        if (lane == 0) { wcount_gt[warp] = __popc(bg); wcount_eq[warp] = __popc(be); }
        __syncthreads();
        unsigned og = out_gt, oe = out_eq;
        for (int w = 0; w < warp; ++w) { og += wcount_gt[w]; oe += wcount_eq[w]; }
        const unsigned lower = (1u << lane) - 1u;
        // elements greater than the threshold: slots [0, k - n_eq); equal ones: the next n_eq slots
        if (gt) {
            const unsigned slot = og + __popc(bg & lower);
            top_scores[slot] = sc[i]; top_labels[slot] = i % NC; top_query[slot] = i / NC;
        } else if (eq) {
            const unsigned r = oe + __popc(be & lower);
            if (r < n_eq) {
                const unsigned slot = (unsigned)k - n_eq + r;
                top_scores[slot] = sc[i]; top_labels[slot] = i % NC; top_query[slot] = i / NC;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned tg = 0, te = 0;
            for (int w = 0; w < nwarp; ++w) { tg += wcount_gt[w]; te += wcount_eq[w]; }
            out_gt += tg; out_eq += te;
        }
        __syncthreads();
    }
}

// The detector's instance selection (models/mask2former_vps/mask2former.py:192-201) on the device,
// static shapes: det_score = class score * mask score for thing candidates, 1-based id = rank of the
// candidate among the thing candidates (candidate order), descending sort by det_score, keep `topk`.
// Outputs: boxes6 [topk,6] = (id, x0, y0, x1, y1, det_score), labels [topk], sel_query [topk] (input
// of the mask pass), count[0] = number of thing candidates.  One CTA.
__global__ void __launch_bounds__(128) ins_finalize_kernel(const float* __restrict__ scores,
                                                           const int32_t* __restrict__ labels,
                                                           const int32_t* __restrict__ query,
                                                           const float* __restrict__ stats,
                                                           const int32_t* __restrict__ boxes, int n, int num_things,
                                                           int topk, float* __restrict__ boxes6,
                                                           int32_t* __restrict__ out_labels,
                                                           int32_t* __restrict__ sel_query,
                                                           int32_t* __restrict__ count) {
    __shared__ float ds[1024];
    __shared__ int thing[1024];
    {   // one CTA per frame of the batch
        const int bz = blockIdx.x;
        scores += (int64_t)bz * n; labels += (int64_t)bz * n; query += (int64_t)bz * n;
        stats += (int64_t)bz * 2 * n; boxes += (int64_t)bz * 4 * n;
        boxes6 += (int64_t)bz * 6 * topk; out_labels += (int64_t)bz * topk; sel_query += (int64_t)bz * topk;
        count += bz;
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int t = labels[i] < num_things ? 1 : 0;
        thing[i] = t;
        ds[i] = t ? scores[i] * stats[2 * i] / (stats[2 * i + 1] + 1e-6f) : -1.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int rank = 0, id = 0;
        const float si = ds[i];
        for (int j = 0; j < n; ++j) {
            const float sj = ds[j];
            rank += (sj > si || (sj == si && j < i)) ? 1 : 0;
            id += (j <= i) ? thing[j] : 0;
        }
        if (rank < topk) {
            boxes6[6 * rank + 0] = (float)id;
            boxes6[6 * rank + 1] = (float)boxes[4 * i + 0];
            boxes6[6 * rank + 2] = (float)boxes[4 * i + 1];
            boxes6[6 * rank + 3] = (float)boxes[4 * i + 2];
            boxes6[6 * rank + 4] = (float)boxes[4 * i + 3];
            boxes6[6 * rank + 5] = si;
            out_labels[rank] = labels[i];
            sel_query[rank] = query[i];
        }
    }
    if (threadIdx.x == 0) {
        int c = 0;
        for (int j = 0; j < n; ++j) c += thing[j];
        count[0] = c;
    }
}

int make_geom(UpGeom& g, int h, int w, int in_h, int in_w, int img_h, int img_w, int out_h, int out_w) {
    if (h <= 0 || w <= 0 || in_h <= 0 || in_w <= 0 || img_h <= 0 || img_w <= 0 || out_h <= 0 || out_w <= 0)
        return PVSG_ERR_INVALID_ARG;
    if (img_h > in_h || img_w > in_w) return PVSG_ERR_INVALID_ARG;
    g.h = h; g.w = w; g.in_h = in_h; g.in_w = in_w; g.img_h = img_h; g.img_w = img_w;
    g.out_h = out_h; g.out_w = out_w;
    g.sh = (float)h / (float)in_h; g.sw = (float)w / (float)in_w;
    g.rh = (float)img_h / (float)out_h; g.rw = (float)img_w / (float)out_w;
    g.rescale = (out_h != img_h || out_w != img_w) ? 1 : 0;
    return PVSG_OK;
}

}  // namespace

int pvsg_internal::configure_panoptic() {
    static bool configured[PVSG_MAX_DEVICES];
    if (pvsg_first_use_on_device(configured) &&
        cudaFuncSetAttribute(ins_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
        return PVSG_ERR_LAUNCH;
    return PVSG_OK;
}

extern "C" int pvsg_panoptic_fuse_batched(const float* cls_logits, const float* mask_logits, int B, int Q, int NC,
                                          int num_things, int h, int w, int in_h, int in_w, int img_h, int img_w,
                                          int out_h, int out_w, float object_mask_thr, double iou_thr,
                                          int filter_low_score, int instance_offset, int32_t* pan_out,
                                          int32_t* seg_info, int32_t* work, float* scores, uint16_t* pix_ws,
                                          void* stream) {
    PVSG_CHECK_ARG(cls_logits && mask_logits && pan_out && seg_info && work && scores && pix_ws);
    PVSG_CHECK_ARG(B > 0 && B <= 65535 && Q > 0 && Q <= 1024 && NC > 0 && num_things >= 0 && num_things <= NC);
    UpGeom g{};
    int rc = make_geom(g, h, w, in_h, in_w, img_h, img_w, out_h, out_w);
    if (rc != PVSG_OK) return rc;
    cudaStream_t st = as_stream(stream);
    const int64_t npix = (int64_t)out_h * out_w;
    pan_select_kernel<<<B, 1024, 0, st>>>(cls_logits, Q, NC, object_mask_thr, seg_info, work, scores);
    if (!g.rescale) {
        dim3 grid((unsigned)((out_w + 255) / 256), (unsigned)((out_h + PAN_ROWS - 1) / PAN_ROWS), (unsigned)B);
        pan_pixel_fast_kernel<<<grid, 256, sizeof(int) * 3 * Q, st>>>(mask_logits, g, Q, seg_info, scores, work, pix_ws);
    } else {
        dim3 grid((unsigned)((npix + 255) / 256), (unsigned)B);
        pan_pixel_kernel<<<grid, 256, sizeof(int) * 3 * Q, st>>>(mask_logits, g, Q, seg_info, scores, work, pix_ws);
    }
    pan_decide_kernel<<<B, 32, 0, st>>>(Q, num_things, iou_thr, filter_low_score, instance_offset, seg_info, work);
    dim3 gw((unsigned)imin64((npix + 255) / 256, 148 * 16), (unsigned)B);
    pan_write_kernel<<<gw, 256, 0, st>>>(pix_ws, seg_info, NC, filter_low_score, pan_out, npix, Q);
    return pvsg_launch_status();
}

extern "C" int pvsg_panoptic_fuse(const float* cls_logits, const float* mask_logits, int Q, int NC,
                                  int num_things, int h, int w, int in_h, int in_w, int img_h, int img_w,
                                  int out_h, int out_w, float object_mask_thr, double iou_thr,
                                  int filter_low_score, int instance_offset, int32_t* pan_out,
                                  int32_t* seg_info, int32_t* work, float* scores, uint16_t* pix_ws,
                                  void* stream) {
    return pvsg_panoptic_fuse_batched(cls_logits, mask_logits, 1, Q, NC, num_things, h, w, in_h, in_w, img_h, img_w,
                                      out_h, out_w, object_mask_thr, iou_thr, filter_low_score, instance_offset,
                                      pan_out, seg_info, work, scores, pix_ws, stream);
}

extern "C" int pvsg_instance_masks_batched(const float* mask_logits, const int32_t* query_idx, int B, int Q, int n,
                                           int h, int w, int in_h, int in_w, int img_h, int img_w, int out_h,
                                           int out_w, float* stats, int32_t* boxes, uint8_t* masks_out,
                                           void* stream) {
    PVSG_CHECK_ARG(mask_logits && query_idx && stats && boxes && B > 0 && Q > 0 && n > 0 && (int64_t)B * n <= 65535);
    UpGeom g{};
    int rc = make_geom(g, h, w, in_h, in_w, img_h, img_w, out_h, out_w);
    if (rc != PVSG_OK) return rc;
    cudaStream_t st = as_stream(stream);
    const int64_t npix = (int64_t)out_h * out_w;
    const int nb = B * n;
    ins_init_kernel<<<(nb + 127) / 128, 128, 0, st>>>(stats, boxes, nb);
    if (!g.rescale) {
        dim3 grid((unsigned)((out_w + 255) / 256), (unsigned)((out_h + INS_ROWS - 1) / INS_ROWS), (unsigned)nb);
        if (masks_out)
            ins_pixel_fast_kernel<true><<<grid, 256, 0, st>>>(mask_logits, query_idx, g, stats, boxes, masks_out, Q, n);
        else
            ins_pixel_fast_kernel<false><<<grid, 256, 0, st>>>(mask_logits, query_idx, g, stats, boxes, masks_out, Q, n);
    } else {
        dim3 grid((unsigned)((npix + 256 * INS_STRIPS - 1) / (256 * INS_STRIPS)), (unsigned)nb);
        ins_pixel_kernel<<<grid, 256, 0, st>>>(mask_logits, query_idx, g, stats, boxes, masks_out, Q, n);
    }
    ins_finish_kernel<<<(nb + 127) / 128, 128, 0, st>>>(boxes, nb);
    return pvsg_launch_status();
}

extern "C" int pvsg_instance_masks(const float* mask_logits, const int32_t* query_idx, int n, int h, int w,
                                   int in_h, int in_w, int img_h, int img_w, int out_h, int out_w,
                                   float* stats, int32_t* boxes, uint8_t* masks_out, void* stream) {
    // a single frame: the query count only enters as the per-frame stride, which is unused for B = 1
    return pvsg_instance_masks_batched(mask_logits, query_idx, 1, 1 << 20, n, h, w, in_h, in_w, img_h, img_w, out_h,
                                       out_w, stats, boxes, masks_out, stream);
}

extern "C" int pvsg_instance_select_batched(const float* cls_logits, int B, int Q, int NC, int k, float* top_scores,
                                            int32_t* top_labels, int32_t* top_query, void* stream) {
    PVSG_CHECK_ARG(cls_logits && top_scores && top_labels && top_query && B > 0 && Q > 0 && NC > 0 && k > 0);
    PVSG_CHECK_ARG((int64_t)k <= (int64_t)Q * NC);
    const size_t smem = sizeof(float) * (size_t)Q * NC;
    if (smem > 200 * 1024) return PVSG_ERR_UNSUPPORTED;
    if (const int rc = pvsg_internal::configure_panoptic()) return rc;
    ins_select_kernel<<<B, 1024, smem, as_stream(stream)>>>(cls_logits, Q, NC, k, top_scores, top_labels, top_query);
    return pvsg_launch_status();
}

extern "C" int pvsg_instance_select(const float* cls_logits, int Q, int NC, int k, float* top_scores,
                                    int32_t* top_labels, int32_t* top_query, void* stream) {
    return pvsg_instance_select_batched(cls_logits, 1, Q, NC, k, top_scores, top_labels, top_query, stream);
}

extern "C" int pvsg_instance_finalize_batched(const float* scores, const int32_t* labels, const int32_t* query,
                                              const float* stats, const int32_t* boxes, int B, int n, int num_things,
                                              int topk, float* boxes6, int32_t* out_labels, int32_t* sel_query,
                                              int32_t* count, void* stream) {
    PVSG_CHECK_ARG(scores && labels && query && stats && boxes && boxes6 && out_labels && sel_query && count);
    PVSG_CHECK_ARG(B > 0 && n > 0 && n <= 1024 && topk > 0 && topk <= n && num_things >= 0);
    ins_finalize_kernel<<<B, 128, 0, as_stream(stream)>>>(scores, labels, query, stats, boxes, n, num_things, topk,
                                                          boxes6, out_labels, sel_query, count);
    return pvsg_launch_status();
}

extern "C" int pvsg_instance_finalize(const float* scores, const int32_t* labels, const int32_t* query,
                                      const float* stats, const int32_t* boxes, int n, int num_things, int topk,
                                      float* boxes6, int32_t* out_labels, int32_t* sel_query, int32_t* count,
                                      void* stream) {
    return pvsg_instance_finalize_batched(scores, labels, query, stats, boxes, 1, n, num_things, topk, boxes6,
                                          out_labels, sel_query, count, stream);
}

// ------------------------------------------------------------------------------------------
// Device-side run-length events for the tube wire format (reference: concat_seq,
// models/mask2former_vps/utils.py:38-54 -> pycocotools RLE of `pan == id` per kept segment, written as
// `masks.txt` rows by models/unitrack/utils/io.py:14-37).  COCO RLE walks the mask COLUMN-major; a
// run boundary of segment s sits wherever the label changes to or from s along that walk.  One pass
// over the panoptic map finds the boundaries of ALL kept segments at once:
//   rle_count_kernel : thread = image column, counts its boundary events
//   rle_scan_kernel  : exclusive scan over the columns of a frame (events stay in walk order)
//   rle_emit_kernel  : same walk, writes (position, segment slot) at the scanned offset
// The host gets ~10^4 events per frame instead of the 3.7 MB map; counts = differences of a
// segment's positions (first run = zeros, so a segment that owns pixel 0 starts with a 0 count).
namespace {

__device__ __forceinline__ int rle_slot(const int* __restrict__ ids, int n, int label) {
    for (int k = 0; k < n; ++k)
        if (ids[k] == label) return k;
    return -1;
}

template <bool EMIT>
__global__ void __launch_bounds__(256) rle_walk_kernel(const int32_t* __restrict__ pan, const int32_t* __restrict__ seg_info,
                                                       int Q, int H, int W, int32_t* __restrict__ col_events,
                                                       const int32_t* __restrict__ col_offset, uint32_t* __restrict__ ev_pos,
                                                       int16_t* __restrict__ ev_slot, int cap) {
    __shared__ int ids[1024];
    __shared__ int nseg;
    const int b = blockIdx.y;
    pan += (int64_t)b * H * W;
    seg_info += (int64_t)b * (1 + 4 * Q);
    if (threadIdx.x == 0) {
        int n = 0;
        const int kept = seg_info[0];
        for (int k = 0; k < kept; ++k) {
            const int seg = seg_info[1 + 4 * k + 2];
            // a stuff class kept by several queries is ONE segment: first occurrence defines the slot
            bool dup = false;
            for (int j = 0; j < n; ++j) dup = dup || ids[j] == seg;
            if (seg >= 0 && !dup) ids[n++] = seg;
        }
        nseg = n;
    }
    __syncthreads();
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    const int n = nseg;
    int prev = x == 0 ? INT_MIN : pan[(int64_t)(H - 1) * W + x - 1];
    int count = 0;
    int64_t off = EMIT ? col_offset[(int64_t)b * W + x] : 0;
    for (int y = 0; y < H; ++y) {
        const int cur = __ldg(pan + (int64_t)y * W + x);
        if (cur != prev) {
            const int sp = prev == INT_MIN ? -1 : rle_slot(ids, n, prev);
            const int sc = rle_slot(ids, n, cur);
            const uint32_t pos = (uint32_t)x * (uint32_t)H + (uint32_t)y;
            if (sp >= 0) {      // segment `prev` ends here
                if (EMIT && off + count < cap) { ev_pos[(int64_t)b * cap + off + count] = pos; ev_slot[(int64_t)b * cap + off + count] = (int16_t)sp; }
                ++count;
            }
            if (sc >= 0) {      // segment `cur` starts here
                if (EMIT && off + count < cap) { ev_pos[(int64_t)b * cap + off + count] = pos; ev_slot[(int64_t)b * cap + off + count] = (int16_t)sc; }
                ++count;
            }
            prev = cur;
        }
    }
    if (!EMIT) col_events[(int64_t)b * W + x] = count;
}

// exclusive scan of the per-column event counts of one frame (one CTA per frame), total -> n_events[b]
__global__ void __launch_bounds__(1024) rle_scan_kernel(const int32_t* __restrict__ col_events, int32_t* __restrict__ col_offset,
                                                        int32_t* __restrict__ n_events, int W) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    const int b = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < W; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = i < W ? col_events[(int64_t)b * W + i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        int woff = 0;
        for (int w = 0; w < warp; ++w) woff += warp_tot[w];
        if (i < W) col_offset[(int64_t)b * W + i] = carry + woff + incl - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry += woff + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) n_events[b] = carry;
}

}  // namespace

extern "C" int pvsg_rle_events(const int32_t* pan, const int32_t* seg_info, int B, int Q, int H, int W,
                               int32_t* col_ws, uint32_t* ev_pos, int16_t* ev_slot, int32_t* n_events, int cap,
                               void* stream) {
    PVSG_CHECK_ARG(pan && seg_info && col_ws && ev_pos && ev_slot && n_events);
    PVSG_CHECK_ARG(B > 0 && B <= 65535 && Q > 0 && Q <= 1024 && H > 0 && W > 0 && cap > 0 && (int64_t)H * W < (1LL << 32));
    cudaStream_t st = as_stream(stream);
    int32_t* col_events = col_ws;                       // [B, W]
    int32_t* col_offset = col_ws + (int64_t)B * W;      // [B, W]
    dim3 grid((unsigned)((W + 255) / 256), (unsigned)B);
    rle_walk_kernel<false><<<grid, 256, 0, st>>>(pan, seg_info, Q, H, W, col_events, nullptr, nullptr, nullptr, cap);
    rle_scan_kernel<<<B, 1024, 0, st>>>(col_events, col_offset, n_events, W);
    rle_walk_kernel<true><<<grid, 256, 0, st>>>(pan, seg_info, Q, H, W, nullptr, col_offset, ev_pos, ev_slot, cap);
    return pvsg_launch_status();
}
