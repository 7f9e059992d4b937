// micro-benchmark: flush of one sliding-window element step (1008 FP64 values into 12 matrix-column streams)
//   MODE 0: 1008 RED.E.ADD.F64 per element (what k_jacobian_sw does)
//   MODE 1: direct entries staged in shared memory as runs of 7 and added by the TMA engine (cp.reduce.async.bulk .add.f64, the
//           16-byte aligned 48 bytes of every run), run edges and the transposed entries stay RED: 432 RED + 96 bulk ops per element
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_bulk red_bulk.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ void bulk_red_add(double* gdst, const double* ssrc, unsigned bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(gdst), "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(64, 5) k_flush(double* __restrict__ val, int nel1, int nrows, double seed) {
    __shared__ __align__(16) double s_run[64][6][8];
    const int tid = threadIdx.x, i2 = (tid >> 3) & 3, jj = (tid & 7) + 8 * (tid >> 5), jcls = jj & 3, b2 = jj >> 2;
    const int row = blockIdx.x % nrows, seg = blockIdx.x / nrows, seg_len = 16;
    const int n1 = nel1 + 3, ncol = n1 * (nrows + 3);
    double v = seed + tid;
    for (int e1 = seg * seg_len; e1 < min(nel1, (seg + 1) * seg_len); ++e1) {
        const int b = (jcls - e1) & 3;
        const int J = (e1 + b) + n1 * (row + b2), I = e1 + n1 * (row + i2);
        const int st = (3 - b) + 7 * (i2 - b2 + 3);
        v = v * 1.0000001 + 1e-9;
        if (MODE == 0) {
#pragma unroll
            for (int dd = 0; dd < 3; ++dd)
#pragma unroll
                for (int c = 0; c < 3; ++c) atomicAdd(val + ((size_t)(dd * ncol + J) * 147 + st + c * 49), v);
            if (b == 0) {
#pragma unroll
                for (int a = 1; a < 4; ++a)
#pragma unroll
                    for (int dd = 0; dd < 3; ++dd)
#pragma unroll
                        for (int c = 0; c < 3; ++c) atomicAdd(val + ((size_t)(dd * ncol + J) * 147 + st + a + c * 49), v);
            }
        } else {
            // transposed entries of the three pairs c < dd: column (I, c), as before
#pragma unroll
            for (int k = 0; k < 3; ++k) atomicAdd(val + ((size_t)(k * ncol + I) * 147 + (48 - st) + (k + 1) % 3 * 49), v);
            // direct entries of the six pairs c <= dd: run position st1 = 3 - b of this thread's six runs
            size_t g[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const int dd = k < 1 ? 0 : (k < 3 ? 1 : 2), c = k - dd * (dd + 1) / 2;
                g[k] = (size_t)(dd * ncol + J) * 147 + (7 * (i2 - b2 + 3)) + c * 49;     // start of the run (st1 = 0)
            }
            if (b == 3) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the runs of the previous column have been read
#pragma unroll
            for (int k = 0; k < 6; ++k) s_run[tid][k][(3 - b) + (g[k] & 1)] = v;
            if (b == 0) {
#pragma unroll
                for (int a = 1; a < 4; ++a) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) atomicAdd(val + ((size_t)(k * ncol + I + a) * 147 + (48 - st - a) + (k + 1) % 3 * 49), v);
#pragma unroll
                    for (int k = 0; k < 6; ++k) s_run[tid][k][3 + a + (g[k] & 1)] = v;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    const int par = (int)(g[k] & 1);
                    // aligned 48 bytes by the TMA engine, the odd element (first or last of the run) by RED
                    bulk_red_add(val + g[k] + par, &s_run[tid][k][2 * par], 48);
                    atomicAdd(val + g[k] + (par ? 0 : 6), v);
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char** argv) {
    const int nel1 = 576, nrows = 576;
    const size_t n = (size_t)3 * (nel1 + 3) * (nrows + 3) * 147 + 1024;
    double* val;
    cudaMalloc(&val, n * sizeof(double));
    cudaMemset(val, 0, n * sizeof(double));
    const int grid = nrows * ((nel1 + 15) / 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaMemsetAsync(val, 0, n * sizeof(double));
            cudaEventRecord(e0);
            if (mode == 0) k_flush<0><<<grid, 64>>>(val, nel1, nrows, 1.0);
            else k_flush<1><<<grid, 64>>>(val, nel1, nrows, 1.0);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        // checksum: both modes add the same number of values
        printf("mode %d: %.3f ms  (%s)  err=%s\n", mode, best, mode ? "432 RED + 96 bulk reduce per element" : "1008 RED per element", cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
