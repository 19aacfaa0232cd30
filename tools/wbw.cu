// Store-pattern microbenchmark: how fast can a [M, N] fp32 matrix be written when
//  (a) each thread owns a row and writes 16 B pieces (the tcgen05.ld-natural epilogue pattern),
//  (b) a warp writes whole 128 B lines (coalesced),
//  (c) cudaMemsetAsync.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/wbw tools/wbw.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128) rows16(float* c, long M, int N) {
    // block = 128 rows, thread = one row; loops over the row in 32-column chunks like the GEMM epilogue
    for (long m0 = (long)blockIdx.x * 128; m0 < M; m0 += (long)gridDim.x * 128) {
        const long row = m0 + threadIdx.x;
        if (row >= M) continue;
        float* cc = c + row * N;
        for (int n = 0; n < N; n += 32) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                reinterpret_cast<float4*>(cc + n)[j] = make_float4(1.f, 2.f, 3.f, (float)j);
        }
    }
}

__global__ void __launch_bounds__(128) lines128(float* c, long M, int N) {
    // 8 lanes x 16 B = one 128 B line segment of a row; a warp writes 4 rows per instruction
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long m0 = (long)blockIdx.x * 128; m0 < M; m0 += (long)gridDim.x * 128) {
        for (int n = 0; n < N; n += 32) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const long row = m0 + warp * 32 + j * 4 + (lane >> 3);
                if (row < M)
                    reinterpret_cast<float4*>(c + row * N + n)[lane & 7] = make_float4(1.f, 2.f, 3.f, (float)j);
            }
        }
    }
}

__global__ void __launch_bounds__(256) flat(float4* c, long n4) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x)
        c[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}

int main() {
    const long M = 154560;
    for (int N : {256, 1024}) {
        float* c;
        const size_t bytes = (size_t)M * N * 4;
        cudaMalloc(&c, bytes);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int mode = 0; mode < 5; ++mode) {
            float best = 1e9;
            for (int rep = 0; rep < 6; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) rows16<<<148, 128>>>(c, M, N);
                if (mode == 1) rows16<<<148 * 4, 128>>>(c, M, N);
                if (mode == 2) lines128<<<148 * 4, 128>>>(c, M, N);
                if (mode == 3) flat<<<148 * 8, 256>>>(reinterpret_cast<float4*>(c), (long)(bytes / 16));
                if (mode == 4) cudaMemsetAsync(c, 0, bytes);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep > 0 && ms < best) best = ms;
            }
            const char* names[] = {"rows16 1cta/sm (4 warps)", "rows16 4cta/sm", "lines128 4cta/sm", "flat float4", "memset"};
            printf("N=%d %-26s %8.1f us  %7.0f GB/s\n", N, names[mode], best * 1e3, bytes / best / 1e6);
        }
        cudaFree(c);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
