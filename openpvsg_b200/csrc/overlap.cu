// Ground-truth / predicted segment overlap counts for the relation-set builder (sm_100a).
//
// Reference: match_and_process_gt_tubes, utils/relation_matching.py:205-260 -- for every frame,
// every GT object and every predicted tube of the same class it decodes the tube's RLE mask and
// computes calculate_iou (:156-165) = |gt & pred| / |gt | pred| with two full-frame numpy passes
// (O(frames x objects x tubes x H x W) bytes).  Both sides are PARTITIONS of the frame (GT: the
// instance-id PNG; pred: the panoptic id map the RLE rows were cut from), so every IoU of a frame
// follows from ONE joint histogram count[g][s] = #{pixels : gt == g and pred slot == s}:
//   inter = count[g][s], |gt g| = sum_s count[g][s], |pred s| = sum_g count[g][s].
// One HBM pass over two int32 maps per frame (7.5 MB at 720p) instead of ~GBs; HBM-bound.
//
// A thread walks a strip of one image row and flushes one shared-memory atomic per RUN of equal
// (gt, pred) labels (segments are spatially coherent: a 720p frame has ~10^4 runs for 9.4e5
// pixels); the CTA histogram is merged into HBM with one atomic per non-zero cell.  The id -> slot
// lookup at a run boundary is a shared-memory hash probe (a linear search over the kept ids cost
// ~160 instructions per strip on maps with 40 segments and capped the kernel at 21 % of HBM peak).
#include "common.cuh"
#include <limits.h>

namespace {

constexpr int kStrip = 16;   // pixels per thread (four int4 loads per map)
constexpr int kHash = 2048;  // >= 2 x the 1024 segment slots a frame can have

__device__ __forceinline__ int hash_id(int id) { return (int)(((uint32_t)id * 2654435761u) >> 21) & (kHash - 1); }

// slot of a panoptic id, or `none`; INT_MIN marks an empty cell (no label uses it: it is also the "no run yet" marker)
__device__ __forceinline__ int find_slot(const int* hkey, const short* hval, int id, int none) {
    if (id == INT_MIN) return none;
    int h = hash_id(id);
    while (true) {
        const int k = hkey[h];
        if (k == id) return hval[h];
        if (k == INT_MIN) return none;
        h = (h + 1) & (kHash - 1);
    }
}

__global__ void __launch_bounds__(256) overlap_kernel(const int32_t* __restrict__ gt, const int32_t* __restrict__ pan,
                                                      const int32_t* __restrict__ seg_info, int Q, int64_t HW, int G,
                                                      int ctas_per_frame, int vec, int32_t* __restrict__ counts) {
    extern __shared__ int32_t hist[];     // [(G + 1), (Q + 1)]
    __shared__ int hkey[kHash];            // open-addressing table: panoptic id -> slot (load factor <= 0.5)
    __shared__ short hval[kHash];
    const int b = blockIdx.y;
    const int cols = Q + 1;
    const int cells = (G + 1) * cols;
    gt += (int64_t)b * HW;
    pan += (int64_t)b * HW;
    seg_info += (int64_t)b * (1 + 4 * Q);
    counts += (int64_t)b * cells;
    for (int i = threadIdx.x; i < cells; i += blockDim.x) hist[i] = 0;
    for (int i = threadIdx.x; i < kHash; i += blockDim.x) hkey[i] = INT_MIN;
    __syncthreads();
    if (threadIdx.x == 0) {   // slot = order of first appearance among the kept rows (as pvsg_rle_events)
        int n = 0;
        const int kept = seg_info[0];
        for (int k = 0; k < kept; ++k) {
            const int seg = seg_info[1 + 4 * k + 2];
            if (seg >= 0 && find_slot(hkey, hval, seg, -1) < 0) {   // a stuff class kept by several queries is one slot
                int h = hash_id(seg);
                while (hkey[h] != INT_MIN) h = (h + 1) & (kHash - 1);
                hkey[h] = seg;
                hval[h] = (short)n;
                ++n;
            }
        }
    }
    __syncthreads();
    const int64_t strips = (HW + kStrip - 1) / kStrip;
    // the trip count is uniform over the CTA (strips beyond the map just carry len = 0), so the warp-level
    // aggregation of the closing run below sees all 32 lanes
    for (int64_t sb = (int64_t)blockIdx.x * blockDim.x; sb < strips; sb += (int64_t)ctas_per_frame * blockDim.x) {
        const int64_t base = (sb + threadIdx.x) * kStrip;
        int g_run = INT_MIN, p_run = INT_MIN, cell = 0, len = 0;
        if (base >= HW) {
            // nothing to read
        } else if (vec && base + kStrip <= HW) {
            int4 gv[kStrip / 4], pv[kStrip / 4];
#pragma unroll
            for (int v = 0; v < kStrip / 4; ++v) {
                gv[v] = __ldg(reinterpret_cast<const int4*>(gt + base) + v);
                pv[v] = __ldg(reinterpret_cast<const int4*>(pan + base) + v);
            }
            const int* ga = reinterpret_cast<const int*>(gv);
            const int* pa = reinterpret_cast<const int*>(pv);
#pragma unroll
            for (int i = 0; i < kStrip; ++i) {
                const int g = ga[i], p = pa[i];
                if (g != g_run || p != p_run) {
                    if (len) atomicAdd(&hist[cell], len);
                    const int slot = find_slot(hkey, hval, p, Q);   // column Q: not a kept segment
                    const int row = (g >= 0 && g < G) ? g : G;   // row G: ids outside [0, G)
                    cell = row * cols + slot;
                    g_run = g; p_run = p; len = 0;
                }
                ++len;
            }
        } else {
            for (int64_t i = base; i < base + kStrip && i < HW; ++i) {
                const int g = __ldg(gt + i), p = __ldg(pan + i);
                if (g != g_run || p != p_run) {
                    if (len) atomicAdd(&hist[cell], len);
                    const int slot = find_slot(hkey, hval, p, Q);
                    const int row = (g >= 0 && g < G) ? g : G;
                    cell = row * cols + slot;
                    g_run = g; p_run = p; len = 0;
                }
                ++len;
            }
        }
        // closing run of every lane: neighbouring strips mostly end in the same (gt, segment) cell -- one atomic per
        // group of equal cells instead of up to 32 same-address atomics
        const unsigned grp = __match_any_sync(0xffffffffu, len ? cell : -1 - (int)(threadIdx.x & 31));
        const int total = __reduce_add_sync(grp, len);
        if (total && (threadIdx.x & 31) == __ffs(grp) - 1) atomicAdd(&hist[cell], total);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cells; i += blockDim.x) {
        const int v = hist[i];
        if (v) atomicAdd(&counts[i], v);
    }
}

}  // namespace

int pvsg_internal::configure_overlap() {
    static bool configured[PVSG_MAX_DEVICES];   // idempotent, value-independent of the call
    if (pvsg_first_use_on_device(configured) &&
        cudaFuncSetAttribute(overlap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
        return PVSG_ERR_LAUNCH;
    return PVSG_OK;
}

extern "C" int pvsg_tube_overlap(const int32_t* gt, const int32_t* pan, const int32_t* seg_info, int B, int Q,
                                 int H, int W, int G, int32_t* counts, void* stream) {
    PVSG_CHECK_ARG(gt && pan && seg_info && counts);
    PVSG_CHECK_ARG(B > 0 && B <= 65535 && Q > 0 && Q <= 1024 && H > 0 && W > 0 && G > 0);
    const int64_t cells = (int64_t)(G + 1) * (Q + 1);
    const size_t smem = (size_t)cells * sizeof(int32_t);
    PVSG_CHECK_ARG(smem <= 200 * 1024);
    cudaStream_t st = as_stream(stream);
    if (const int rc = pvsg_internal::configure_overlap()) return rc;
    if (cudaMemsetAsync(counts, 0, (size_t)B * cells * sizeof(int32_t), st) != cudaSuccess) return PVSG_ERR_LAUNCH;
    const int64_t HW = (int64_t)H * W;
    const int64_t strips = (HW + kStrip - 1) / kStrip;
    // ~8 CTAs per SM (first version: 2 per SM = 16 warps per SM, 21 % of HBM peak -- too little in flight for a
    // streaming kernel whose threads serialise on their run atomics), but no more than the strips need
    int per_frame = (int)imin64((strips + 255) / 256, (int64_t)((8 * 148 + B - 1) / B));
    if (per_frame < 1) per_frame = 1;
    dim3 grid((unsigned)per_frame, (unsigned)B);
    // int4 loads need 16-byte aligned frames
    const int vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(gt) | reinterpret_cast<uintptr_t>(pan)) & 15) == 0;
    overlap_kernel<<<grid, 256, smem, st>>>(gt, pan, seg_info, Q, HW, G, per_frame, vec, counts);
    return pvsg_launch_status();
}
