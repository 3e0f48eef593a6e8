// FP64 micro-benchmarks for the second roofline of the bundle-adjustment kernels (SURVEY.md 8d asks the builder to
// measure the FP64 peak; MEASURED_PEAKS.json only holds HBM and bf16 numbers).
//   dfma  : dependent-chain-free DFMA stream, 8 independent accumulators per thread
//   dmma  : mma.sync.aligned.m8n8k4.f64 stream, 4 independent accumulator fragments per warp
//   lds   : shared-memory read-modify-write of doubles (LDS.64 + DADD + STS.64), conflict-free, per SM rate
//   atoms : atomicAdd(double) on shared memory (CAS loop on sm_100a), conflict-free addresses
// Build + run (GPU box):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_peak tools/fp64_peak.cu && /tmp/fp64_peak
// Output: one JSON object on stdout.
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b)
{
    double x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = threadIdx.x * 1e-3 + k;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = fma(x[k], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double a, double b)
{
    double c[4][2];
#pragma unroll
    for (int k = 0; k < 4; ++k) { c[k][0] = threadIdx.x * 1e-3; c[k][1] = k; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) s += c[k][0] + c[k][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_lds_rmw(double* out, int iters, double a)
{
    __shared__ double s[256 * 8];
    for (int i = threadIdx.x; i < 256 * 8; i += 256) s[i] = i;
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {                                   // conflict-free: consecutive lanes, consecutive doubles
            volatile double* q = &s[k * 256 + threadIdx.x];              // volatile: one LDS.64 + DADD + STS.64 per update
            *q = *q + a;
        }
    }
    __syncthreads();
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s[threadIdx.x];
}

__global__ void __launch_bounds__(256) k_atoms(double* out, int iters, double a)
{
    __shared__ double s[256 * 8];
    for (int i = threadIdx.x; i < 256 * 8; i += 256) s[i] = i;
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(&s[k * 256 + threadIdx.x], a);
    }
    __syncthreads();
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s[threadIdx.x];
}

template <typename F>
static float time_ms(F launch, int reps)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();                       // warm-up
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return best;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const int blocks = sms * 8, threads = 256, iters = 4096;
    double* out;
    CK(cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)));
    const float t_fma = time_ms([&] { k_dfma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
    const float t_mma = time_ms([&] { k_dmma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
    const float t_lds = time_ms([&] { k_lds_rmw<<<blocks, threads>>>(out, iters / 4, 1e-9); }, 5);
    const float t_atm = time_ms([&] { k_atoms<<<blocks, threads>>>(out, iters / 16, 1e-9); }, 5);
    CK(cudaGetLastError());
    const double n_fma = (double)blocks * threads * iters * 8;                        // DFMA lane-operations
    const double n_mma = (double)blocks * (threads / 32) * iters * 4 * 256;           // FMA per m8n8k4 = 8*8*4
    const double n_lds = (double)blocks * threads * (iters / 4) * 8;
    const double n_atm = (double)blocks * threads * (iters / 16) * 8;
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_attr\": %d, "
           "\"dfma_tflops\": %.2f, \"dfma_gfma_per_s\": %.1f, \"dmma_m8n8k4_tflops\": %.2f, "
           "\"smem_rmw_f64_per_s_G\": %.1f, \"smem_atomicadd_f64_per_s_G\": %.1f, "
           "\"ms\": {\"dfma\": %.4f, \"dmma\": %.4f, \"lds_rmw\": %.4f, \"atoms\": %.4f}}\n",
           prop.name, sms, clk, 2.0 * n_fma / (t_fma * 1e-3) / 1e12, n_fma / (t_fma * 1e-3) / 1e9,
           2.0 * n_mma / (t_mma * 1e-3) / 1e12, n_lds / (t_lds * 1e-3) / 1e9, n_atm / (t_atm * 1e-3) / 1e9,
           t_fma, t_mma, t_lds, t_atm);
    cudaFree(out);
    return 0;
}
