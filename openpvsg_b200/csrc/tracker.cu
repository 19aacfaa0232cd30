// IPS tracker path (reference models/unitrack): the two numeric kernels of the association step.
//
//  pvsg_reconsdot   UniTrack's reconstruction-similarity distance between track and detection embeddings
//                   (core/association/matching.py:194-238): mask-pooled appearance features [n, positions, d], zero
//                   padded to the longest; affinity of every (track position, detection position) pair, softmax over
//                   ALL positions of the other side (temperature 100, padded positions included, as the reference),
//                   per-pair reconstruction of one side from the other, cosine of the reconstruction with the original.
//  pvsg_lap_assign  lap.lapjv(cost, extend_cost=True, cost_limit=thresh) (matching.py:29-40): exact linear assignment
//                   on the (n + m) square extension, Jonker-Volgenant shortest augmenting paths in fp64, ONE warp --
//                   the problems are tiny (tens of tracks), what matters is that the step is stream-ordered.
#include "common.cuh"

namespace {

constexpr float kEps = 1e-12f;   // F.normalize eps

// y[p, :] = x[p, :] / max(||x[p, :]||, eps) for every position p (one warp per position)
__global__ void __launch_bounds__(256) rd_normalize_kernel(const float* __restrict__ x, float* __restrict__ y, int P, int d) {
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= P) return;
    float s = 0.f;
    for (int c = lane; c < d; c += 32) { const float v = x[(int64_t)p * d + c]; s = fmaf(v, v, s); }
    s = warp_sum(s);
    const float inv = 1.f / fmaxf(sqrtf(s), kEps);
    for (int c = lane; c < d; c += 32) y[(int64_t)p * d + c] = x[(int64_t)p * d + c] * inv;
}

// aff[i, j] = <a[i, :], b[j, :]>
__global__ void __launch_bounds__(256) rd_affinity_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                          float* __restrict__ aff, int T, int D, int d) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)T * D) return;
    const int r = (int)(i / D), c = (int)(i % D);
    float s = 0.f;
    for (int k = 0; k < d; ++k) s = fmaf(a[(int64_t)r * d + k], b[(int64_t)c * d + k], s);
    aff[i] = s;
}

// log-sum-exp of tmp * aff along rows (mode 0: one warp per row) or columns (mode 1: one warp per column)
__global__ void __launch_bounds__(256) rd_lse_kernel(const float* __restrict__ aff, float* __restrict__ lse, int T, int D,
                                                     float tmp, int mode) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int n = mode == 0 ? T : D, len = mode == 0 ? D : T;
    if (w >= n) return;
    float mx = -INFINITY;
    for (int k = lane; k < len; k += 32) mx = fmaxf(mx, tmp * (mode == 0 ? aff[(int64_t)w * D + k] : aff[(int64_t)k * D + w]));
    mx = warp_max(mx);
    float s = 0.f;
    for (int k = lane; k < len; k += 32) s += expf(tmp * (mode == 0 ? aff[(int64_t)w * D + k] : aff[(int64_t)k * D + w]) - mx);
    s = warp_sum(s);
    if (lane == 0) lse[w] = mx + logf(s);
}

// One CTA per (track t, detection e) pair and direction:
//   dir 0: reconstruct the track's positions from the detection's:  r[s, :] = sum_s' softmax_row(t, s)[(e, s')] fdet[e, s', :]
//   dir 1: reconstruct the detection's positions from the track's:  r[s', :] = sum_s softmax_col(e, s')[(t, s)] ftrk[t, s, :]
// and return cos(r, original) over the flattened (position, channel) vector.
__global__ void __launch_bounds__(128) rd_dot_kernel(const float* __restrict__ ftrk, const float* __restrict__ fdet,
                                                     const float* __restrict__ aff, const float* __restrict__ row_lse,
                                                     const float* __restrict__ col_lse, float* __restrict__ dots, int ntrk,
                                                     int nst, int ndet, int nsd, int d, float tmp) {
    const int t = blockIdx.x, e = blockIdx.y, dir = blockIdx.z;
    const int D = ndet * nsd;
    const int na = dir == 0 ? nst : nsd, nb = dir == 0 ? nsd : nst;     // positions of the reconstructed / source side
    const float* fa = dir == 0 ? ftrk + (int64_t)t * nst * d : fdet + (int64_t)e * nsd * d;
    const float* fb = dir == 0 ? fdet + (int64_t)e * nsd * d : ftrk + (int64_t)t * nst * d;
    float num = 0.f, rr = 0.f, ff = 0.f;
    for (int i = threadIdx.x; i < na * d; i += blockDim.x) {
        const int s = i / d, m = i - s * d;
        float r = 0.f;
        for (int k = 0; k < nb; ++k) {
            const float a = dir == 0 ? aff[(int64_t)(t * nst + s) * D + e * nsd + k] : aff[(int64_t)(t * nst + k) * D + e * nsd + s];
            const float l = dir == 0 ? row_lse[t * nst + s] : col_lse[e * nsd + s];
            r = fmaf(expf(tmp * a - l), fb[(int64_t)k * d + m], r);
        }
        const float f = fa[i];
        num = fmaf(r, f, num);
        rr = fmaf(r, r, rr);
        ff = fmaf(f, f, ff);
    }
    __shared__ float sh[3][4];
    num = warp_sum(num); rr = warp_sum(rr); ff = warp_sum(ff);
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = num; sh[1][threadIdx.x >> 5] = rr; sh[2][threadIdx.x >> 5] = ff; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const float n2 = sh[0][0] + sh[0][1] + sh[0][2] + sh[0][3], r2 = sh[1][0] + sh[1][1] + sh[1][2] + sh[1][3],
                    f2 = sh[2][0] + sh[2][1] + sh[2][2] + sh[2][3];
        dots[((int64_t)dir * ntrk + t) * ndet + e] = n2 / (fmaxf(sqrtf(r2), kEps) * fmaxf(sqrtf(f2), kEps));
    }
}

__global__ void rd_cost_kernel(const float* __restrict__ dots, float* __restrict__ cost, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cost[i] = 1.f - 0.5f * (dots[i] + dots[n + i]);
}

// ------------------------------------------------------------------------------------- assignment
constexpr int LAP_MAX = 256;          // n + m
constexpr double LAP_BIG = 1e9;       // stands for +inf entries (never chosen: the extension offers cost_limit instead)

__device__ __forceinline__ double lap_cost(const float* __restrict__ c, int n, int m, double half, int i, int j) {   // 1-based
    if (i <= n && j <= m) { const float v = c[(int64_t)(i - 1) * m + (j - 1)]; return isfinite(v) ? (double)v : LAP_BIG; }
    return (i > n && j > m) ? 0.0 : half;
}

// one warp = one problem (blockIdx.x); extend = 0: plain square problem (n == m), no unmatched option
__global__ void __launch_bounds__(32) lap_kernel(const float* __restrict__ c_all, int n, int m, double cost_limit, int extend,
                                                 int32_t* __restrict__ x_all, int32_t* __restrict__ y_all) {
    __shared__ double u[LAP_MAX + 1], v[LAP_MAX + 1], minv[LAP_MAX + 1];
    __shared__ int p[LAP_MAX + 1], way[LAP_MAX + 1];
    __shared__ unsigned char used[LAP_MAX + 1];
    const float* c = c_all + (int64_t)blockIdx.x * n * m;
    int32_t* x = x_all + (int64_t)blockIdx.x * n;
    int32_t* y = y_all + (int64_t)blockIdx.x * m;
    const int lane = threadIdx.x, N = extend ? n + m : n;
    const double half = isfinite(cost_limit) ? cost_limit * 0.5 : 0.0;
    for (int j = lane; j <= N; j += 32) { u[j] = 0.0; v[j] = 0.0; p[j] = 0; way[j] = 0; }
    __syncwarp();
    for (int i = 1; i <= N; ++i) {
        if (lane == 0) p[0] = i;
        for (int j = lane; j <= N; j += 32) { minv[j] = INFINITY; used[j] = 0; }
        __syncwarp();
        int j0 = 0;
        do {
            if (lane == 0) used[j0] = 1;
            __syncwarp();
            const int i0 = p[j0];
            double delta = INFINITY;
            int j1 = 0x7fffffff;
            for (int j = 1 + lane; j <= N; j += 32) {
                if (!used[j]) {
                    const double cur = lap_cost(c, n, m, half, i0, j) - u[i0] - v[j];
                    if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
                    if (minv[j] < delta) { delta = minv[j]; j1 = j; }       // ascending j: first minimum of the lane
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {      // warp arg-min, ties to the smaller column
                const double od = __shfl_xor_sync(0xffffffffu, delta, o);
                const int oj = __shfl_xor_sync(0xffffffffu, j1, o);
                if (od < delta || (od == delta && oj < j1)) { delta = od; j1 = oj; }
            }
            __syncwarp();
            for (int j = lane; j <= N; j += 32) {
                if (used[j]) { u[p[j]] += delta; v[j] -= delta; }           // p[] is injective on used columns
                else minv[j] -= delta;
            }
            __syncwarp();
            j0 = j1;
        } while (p[j0] != 0);
        if (lane == 0) {
            do { const int j1 = way[j0]; p[j0] = p[j1]; j0 = j1; } while (j0);
        }
        __syncwarp();
    }
    for (int i = lane; i < n; i += 32) x[i] = -1;
    for (int j = lane; j < m; j += 32) y[j] = -1;
    __syncwarp();
    for (int j = 1 + lane; j <= m; j += 32) {
        const int i = p[j];
        if (i >= 1 && i <= n && isfinite(c[(int64_t)(i - 1) * m + (j - 1)])) { x[i - 1] = j - 1; y[j - 1] = i - 1; }
    }
}

}  // namespace

extern "C" int64_t pvsg_reconsdot_workspace_bytes(int ntrk, int nst, int ndet, int nsd, int d) {
    const int64_t T = (int64_t)ntrk * nst, D = (int64_t)ndet * nsd;
    return 4 * (T * d + D * d + T * D + T + D + 2 * (int64_t)ntrk * ndet) + 256;
}

extern "C" int pvsg_reconsdot(const float* trk, const float* det, float* cost, void* workspace, int ntrk, int nst, int ndet,
                              int nsd, int d, float tmp, void* stream) {
    PVSG_CHECK_ARG(trk && det && cost && workspace && ntrk > 0 && nst > 0 && ndet > 0 && nsd > 0 && d > 0);
    if (ndet > 65535 || (int64_t)ntrk * nst * ndet * nsd > (1LL << 30)) return PVSG_ERR_UNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    const int T = ntrk * nst, D = ndet * nsd;
    float* ft = reinterpret_cast<float*>(workspace);
    float* fd = ft + (int64_t)T * d;
    float* aff = fd + (int64_t)D * d;
    float* row_lse = aff + (int64_t)T * D;
    float* col_lse = row_lse + T;
    float* dots = col_lse + D;
    rd_normalize_kernel<<<(T + 7) / 8, 256, 0, st>>>(trk, ft, T, d);
    rd_normalize_kernel<<<(D + 7) / 8, 256, 0, st>>>(det, fd, D, d);
    rd_affinity_kernel<<<(unsigned)(((int64_t)T * D + 255) / 256), 256, 0, st>>>(ft, fd, aff, T, D, d);
    rd_lse_kernel<<<(T + 7) / 8, 256, 0, st>>>(aff, row_lse, T, D, tmp, 0);
    rd_lse_kernel<<<(D + 7) / 8, 256, 0, st>>>(aff, col_lse, T, D, tmp, 1);
    rd_dot_kernel<<<dim3(ntrk, ndet, 2), 128, 0, st>>>(ft, fd, aff, row_lse, col_lse, dots, ntrk, nst, ndet, nsd, d, tmp);
    rd_cost_kernel<<<(ntrk * ndet + 255) / 256, 256, 0, st>>>(dots, cost, ntrk * ndet);
    return pvsg_launch_status();
}

extern "C" int pvsg_lap_assign(const float* cost, int n, int m, double cost_limit, int32_t* x, int32_t* y, void* stream) {
    PVSG_CHECK_ARG(cost && x && y && n > 0 && m > 0);
    if (n + m > LAP_MAX) return PVSG_ERR_UNSUPPORTED;
    lap_kernel<<<1, 32, 0, as_stream(stream)>>>(cost, n, m, cost_limit, 1, x, y);
    return pvsg_launch_status();
}

extern "C" int pvsg_lap_square_batched(const float* cost, int batch, int n, int32_t* x, int32_t* y, void* stream) {
    PVSG_CHECK_ARG(cost && x && y && batch > 0 && n > 0);
    if (n > LAP_MAX) return PVSG_ERR_UNSUPPORTED;
    lap_kernel<<<batch, 32, 0, as_stream(stream)>>>(cost, n, n, INFINITY, 0, x, y);
    return pvsg_launch_status();
}

// cost[t, i, j] = 1 - <a_t[i] / |a_t[i]|, b_t[j] / |b_t[j]|>, a_t = E[t], b_t = E[t + 1]  (one warp per (t, i, j))
namespace {
__global__ void __launch_bounds__(256) cosine_chain_kernel(const float* __restrict__ E, float* __restrict__ cost, int Q, int C,
                                                           int64_t total) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= total) return;
    const int j = (int)(w % Q);
    const int i = (int)((w / Q) % Q);
    const int64_t t = w / ((int64_t)Q * Q);
    const float* a = E + (t * Q + i) * C;
    const float* b = E + ((t + 1) * Q + j) * C;
    float ab = 0.f, aa = 0.f, bb = 0.f;
    for (int k = lane; k < C; k += 32) { const float x = a[k], y = b[k]; ab = fmaf(x, y, ab); aa = fmaf(x, x, aa); bb = fmaf(y, y, bb); }
    ab = warp_sum(ab); aa = warp_sum(aa); bb = warp_sum(bb);
    if (lane == 0) cost[w] = 1.f - ab / (sqrtf(aa) * sqrtf(bb));
}
}  // namespace

namespace {
// perms[0] = identity, perms[t][i] = sigma[t-1][perms[t-1][i]]  (one CTA, Q threads, T-1 dependent steps)
__global__ void perm_chain_kernel(const int32_t* __restrict__ sigma, int32_t* __restrict__ perms, int T, int Q) {
    for (int i = threadIdx.x; i < Q; i += blockDim.x) {
        int p = i;
        perms[i] = p;
        for (int t = 1; t < T; ++t) { p = sigma[(int64_t)(t - 1) * Q + p]; perms[(int64_t)t * Q + i] = p; }
    }
}
}  // namespace

extern "C" int pvsg_perm_chain(const int32_t* sigma, int32_t* perms, int T, int Q, void* stream) {
    PVSG_CHECK_ARG(perms && T > 0 && Q > 0 && (sigma || T == 1));
    perm_chain_kernel<<<1, 128, 0, as_stream(stream)>>>(sigma, perms, T, Q);
    return pvsg_launch_status();
}

extern "C" int pvsg_cosine_chain_cost(const float* embeds, float* cost, int T, int Q, int C, void* stream) {
    PVSG_CHECK_ARG(embeds && cost && T > 1 && Q > 0 && C > 0);
    const int64_t total = (int64_t)(T - 1) * Q * Q;
    cosine_chain_kernel<<<(unsigned)((total + 7) / 8), 256, 0, as_stream(stream)>>>(embeds, cost, Q, C, total);
    return pvsg_launch_status();
}
