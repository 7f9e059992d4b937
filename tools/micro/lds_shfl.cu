// micro-benchmark: shared-memory load and shuffle throughput per SM for the address patterns the Jacobian kernel uses
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_shfl lds_shfl.cu
#include <cstdio>
#include <string>
#include <cuda_runtime.h>
template <int W, int MODE>
__global__ void k_lds(unsigned* out, int iters, long long* cyc) {
    __shared__ __align__(16) unsigned buf[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) buf[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int idx;   // in units of 4 bytes
    if (MODE == 0) idx = 0;                       // all lanes the same address
    else if (MODE == 1) idx = (lane & 3) * 116;   // 4 distinct records 464 B apart (the kernel's pattern)
    else if (MODE == 2) idx = (lane >> 3) * 116;  // quarter-warp uniform
    else if (MODE == 3) idx = lane * W;           // fully distinct, conflict-free
    else idx = (lane & 7) * 116;                  // 8 distinct records
    unsigned a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    const unsigned base = (unsigned)__cvta_generic_to_shared(buf) + idx * 4;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const unsigned addr = base + ((u * 16 * W * 4) % 4096) + ((it & 1) << 12);
            unsigned x, y, z, w;
            if (W == 4) { asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(addr) : "memory"); a0 += x; a1 += y; a2 += z; a3 += w; }
            else if (W == 2) { asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "r"(addr) : "memory"); a0 += x; a1 += y; }
            else { asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(addr) : "memory"); a0 += x; }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void k_shfl(unsigned* out, int iters, long long* cyc) {
    unsigned a = threadIdx.x, b = a * 3;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) { a += __shfl_xor_sync(0xffffffffu, a, 1); b += __shfl_xor_sync(0xffffffffu, b, 2); }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// FP64 FMA stream with a shuffle every K FMAs: does the shuffle steal FP64 issue slots?
template <int K>
__global__ void k_mix(double* out, int iters, long long* cyc) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    unsigned s = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int k = 0; k < K; ++k) { x0 = fma(x0, 0.999, 1e-9); x1 = fma(x1, 0.999, 1e-9); x2 = fma(x2, 0.999, 1e-9); x3 = fma(x3, 0.999, 1e-9); }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <class F> void run(const char* name, F launch, double ops_per_warp_iter, int warps) {
    long long* cyc; cudaMallocManaged(&cyc, 8);
    launch(cyc); cudaDeviceSynchronize(); launch(cyc); cudaDeviceSynchronize();
    printf("%-34s cycles %lld  => %.2f clk per warp-instruction per SM (%d warps/SM)\n", name, *cyc, (double)*cyc / (ops_per_warp_iter * warps), warps);
    cudaFree(cyc);
}
int main() {
    unsigned* out; cudaMalloc(&out, 1 << 24);
    const int iters = 2000, T = 512, warps = T / 32;
    const char* mn[5] = {"same", "4 records", "quarter-uniform", "distinct", "8 records"};
#define RUN(W, M) run((std::string("LDS.") + std::to_string(32 * W) + " " + mn[M]).c_str(), [&](long long* c) { k_lds<W, M><<<148, T>>>(out, iters, c); }, 16.0 * iters, warps);
    RUN(4, 0) RUN(4, 1) RUN(4, 2) RUN(4, 3) RUN(4, 4) RUN(2, 0) RUN(2, 1) RUN(2, 2) RUN(2, 3) RUN(2, 4) RUN(1, 0) RUN(1, 1) RUN(1, 2) RUN(1, 3) RUN(1, 4)
    run("SHFL.BFLY", [&](long long* c) { k_shfl<<<148, T>>>(out, iters, c); }, 16.0 * iters, warps);
    run("DFMA x16 + 1 SHFL (per 16 DFMA)", [&](long long* c) { k_mix<4><<<148, T>>>((double*)out, iters, c); }, 4.0 * 16 * iters, warps);
    run("DFMA x4 + 1 SHFL (per 4 DFMA)", [&](long long* c) { k_mix<1><<<148, T>>>((double*)out, iters, c); }, 4.0 * 4 * iters, warps);
    return 0;
}
