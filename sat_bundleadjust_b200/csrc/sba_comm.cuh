// One-shot all-reduce over NVLink peer memory for the small, latency-bound exchanges of the multi-GPU solve
// ([U | g_c], [S | rhs], a few scalars: 29 KB at 10 cameras x 6 parameters).
//
// Every rank owns one symmetric buffer (cudaMalloc, exported with cudaIpcGetMemHandle and mapped by its peers):
//     data[2][cap] doubles | flag[2][MAX_RANKS] uint64
// Exchange number s (the same monotonically increasing counter on every rank; parity = s & 1):
//   push:  copy the local contribution into the own data[parity], __threadfence_system(), and -- by the last block to
//          finish -- store s with release/system scope into flag[parity][me] of EVERY peer (writes over NVLink);
//   pull:  spin (acquire/system, with a time-out) until the own flag[parity][r] >= s for all r, then add up the R
//          contributions in rank order with volatile peer loads (reads over NVLink) -- the same order and the same
//          operands on every rank, hence bit-identical results -- and store the sum into local memory.
// A slot is rewritten at exchange s + 2; a peer signals s + 1 only after its pull of s has completed (stream order),
// and nobody starts s + 2 before it has seen every peer's s + 1, so two parities suffice.
#pragma once
#include "sba_internal.cuh"

namespace sba {

constexpr int COMM_MAX_RANKS = 16;

struct CommView {            // passed to the kernels by value
    double* data[COMM_MAX_RANKS];                 // data[r]: base of rank r's buffer (own buffer for r == me)
    unsigned long long* flag[COMM_MAX_RANKS];     // flag[r]: flags of rank r
    long long cap;                                // doubles per parity
    long long timeout_cycles;                     // a peer that has not arrived after this many clocks is reported, not waited for
    int me, world;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_volatile_f64(const double* p)
{
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256)
k_comm_push(CommView c, const double* __restrict__ src, long long count, unsigned long long seq, unsigned* counter)
{
    const int parity = (int)(seq & 1ull);
    double* mine = c.data[c.me] + (long long)parity * c.cap;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
        mine[i] = src[i];
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    __threadfence_system();
    if (threadIdx.x < c.world && threadIdx.x != c.me)
        st_release_sys(c.flag[threadIdx.x] + parity * COMM_MAX_RANKS + c.me, seq);
    if (threadIdx.x == 0) *counter = 0u;
}

__global__ void __launch_bounds__(256)
k_comm_pull(CommView c, double* __restrict__ dst, long long count, unsigned long long seq, double* scal)
{
    const int parity = (int)(seq & 1ull);
    __shared__ int ok;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    if (threadIdx.x < c.world && threadIdx.x != c.me) {
        const unsigned long long* f = c.flag[c.me] + parity * COMM_MAX_RANKS + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) < seq) {
            if (clock64() - t0 > c.timeout_cycles) { ok = 0; break; }  // a peer is gone: report instead of hanging
        }
    }
    __syncthreads();
    if (!ok) {
        if (threadIdx.x == 0 && blockIdx.x == 0) scal[SC_COMM_FAIL] = 1.0;
        return;
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int r = 0; r < c.world; ++r) acc += ld_volatile_f64(c.data[r] + (long long)parity * c.cap + i);
        dst[i] = acc;
    }
}

// push + pull in ONE launch (the exchanges are latency: every launch saved is ~5 us of a ~20 us exchange).  The grid
// is at most 16 CTAs, all resident, so a CTA may spin on the peers' flags while its siblings are still pushing; the
// peers publish their flags independently of us, so there is no circular wait.  In place: buf -> own slot -> sum -> buf.
__global__ void __launch_bounds__(256)
k_comm_allreduce(CommView c, double* __restrict__ buf, long long count, unsigned long long seq, unsigned* counter, double* scal)
{
    const int parity = (int)(seq & 1ull);
    double* mine = c.data[c.me] + (long long)parity * c.cap;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
        mine[i] = buf[i];
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    __shared__ int ok;
    if (threadIdx.x == 0) { last = (atomicAdd(counter, 1u) == gridDim.x - 1); ok = 1; }
    __syncthreads();
    if (last) {
        __threadfence_system();
        if (threadIdx.x < c.world && threadIdx.x != c.me)
            st_release_sys(c.flag[threadIdx.x] + parity * COMM_MAX_RANKS + c.me, seq);
        if (threadIdx.x == 0) *counter = 0u;
    }
    if (threadIdx.x < c.world && threadIdx.x != c.me) {
        const unsigned long long* f = c.flag[c.me] + parity * COMM_MAX_RANKS + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) < seq) {
            if (clock64() - t0 > c.timeout_cycles) { ok = 0; break; }  // a peer is gone: report instead of hanging
        }
    }
    __syncthreads();
    if (!ok) {
        if (threadIdx.x == 0) scal[SC_COMM_FAIL] = 1.0;
        return;
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        double acc = 0.0;
        for (int r = 0; r < c.world; ++r)
            acc += (r == c.me) ? mine[i] : ld_volatile_f64(c.data[r] + (long long)parity * c.cap + i);
        buf[i] = acc;
    }
}

// The same exchange done by ONE CTA from inside another kernel -- the epilogue of the kernel that produced `buf` (its last CTA
// to finish, after the grid-wide sums are complete) -- so that the small exchanges of an iteration cost no launch of their own:
// push, flags, spin, pull, all by the calling CTA.  Every thread of the CTA must call it; `buf` must be visible to the CTA
// (written by it, or published by other CTAs with __threadfence() before the counter that elected this CTA).
struct CommFused {           // kernel argument: on == 0 -> no exchange (single GPU, or the exchange is done by separate launches)
    CommView c;
    unsigned long long seq;
    int on;
};

__device__ __forceinline__ void cta_allreduce(const CommView& c, double* buf, int count, unsigned long long seq, double* scal)
{
    const int parity = (int)(seq & 1ull);
    double* mine = c.data[c.me] + (long long)parity * c.cap;
    for (int i = threadIdx.x; i < count; i += blockDim.x) mine[i] = __ldcg(buf + i);
    __threadfence_system();
    __shared__ int ok;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    if (threadIdx.x < c.world && threadIdx.x != c.me) {
        st_release_sys(c.flag[threadIdx.x] + parity * COMM_MAX_RANKS + c.me, seq);
        const unsigned long long* f = c.flag[c.me] + parity * COMM_MAX_RANKS + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) < seq) {
            if (clock64() - t0 > c.timeout_cycles) { ok = 0; break; }  // a peer is gone: report instead of hanging
        }
    }
    __syncthreads();
    if (!ok) {
        if (threadIdx.x == 0) scal[SC_COMM_FAIL] = 1.0;
        __syncthreads();
        return;
    }
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        double acc = 0.0;
        for (int r = 0; r < c.world; ++r)
            acc += (r == c.me) ? mine[i] : ld_volatile_f64(c.data[r] + (long long)parity * c.cap + i);
        buf[i] = acc;
    }
    __threadfence();
    __syncthreads();
}

}  // namespace sba
