// DMMA (mma.sync m8n8k4 f64) issue rate as a function of resident warps per SM and independent accumulators per warp.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/dmma_lat tools/dmma_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int NACC>
__global__ void k(double* out, int iters, double a, double b, long long* cyc)
{
    double c[NACC][2];
#pragma unroll
    for (int q = 0; q < NACC; ++q) { c[q][0] = threadIdx.x * 1e-3; c[q][1] = q; }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < NACC; ++q)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[q][0]), "+d"(c[q][1]) : "d"(a), "d"(b));
    }
    const long long t1 = clock64();
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < NACC; ++q) s += c[q][0] + c[q][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NACC> void run(int threads, double* out, long long* cyc)
{
    const int iters = 2000;
    k<NACC><<<148, threads>>>(out, iters, 1.0000001, 1e-9, cyc);
    cudaDeviceSynchronize();
    k<NACC><<<148, threads>>>(out, iters, 1.0000001, 1e-9, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double per_warp = (double)h / (iters * NACC), warps_per_smsp = threads / 32 / 4.0;
    printf("threads/SM %4d acc/warp %2d: %.1f cycles per DMMA per warp, %.1f cycles per DMMA per SMSP -> %.1f TFLOP/s\n", threads, NACC, per_warp,
           per_warp / warps_per_smsp, 512.0 / (per_warp / warps_per_smsp) * 4 * 148 * 1.965e9 / 1e12);
}
int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(double)); cudaMalloc(&cyc, 8);
    for (int threads : {128, 384, 512, 1024}) { run<1>(threads, out, cyc); run<4>(threads, out, cyc); run<10>(threads, out, cyc); run<21>(threads, out, cyc); }
    return 0;
}
