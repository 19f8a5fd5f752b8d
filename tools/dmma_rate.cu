// Development probe: latency / throughput of fp64 tensor-core MMA (mma.sync m8n8k4 / m16n8k8 / m16n8k16 f64) and of
// dependent DFMA chains on sm_100a - evidence for moving the Riccati stage products of the QP kernel to DMMA.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/dmma_rate tools/dmma_rate.cu && tools/bin/dmma_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma1684(double* c, const double* a, double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}

template <int CHAINS, int KIND>
__global__ void probe(double* out, long long* cyc, int iters) {
    double c[CHAINS][4];
    double a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = 1e-3 * (threadIdx.x + i);
    b[0] = 1e-3; b[1] = 2e-3;
    for (int ch = 0; ch < CHAINS; ++ch) for (int i = 0; i < 4; ++i) c[ch][i] = ch + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ch = 0; ch < CHAINS; ++ch) {
            if (KIND == 0) dmma884(c[ch][0], c[ch][1], a[0], b[0]);
            else if (KIND == 1) dmma1688(c[ch], a, b);
            else if (KIND == 2) dmma1684(c[ch], a, b[0]);
            else { c[ch][0] = fma(c[ch][0], a[0], b[0]); }
        }
    }
    const long long t1 = clock64();
    double s = 0;
    for (int ch = 0; ch < CHAINS; ++ch) for (int i = 0; i < 4; ++i) s += c[ch][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int CHAINS, int KIND>
void run(const char* name, int warps_per_sm, double macs_per_instr) {
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    const int threads = warps_per_sm * 32;
    probe<CHAINS, KIND><<<148, threads>>>(out, cyc, iters);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<CHAINS, KIND><<<148, threads>>>(out, cyc, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double instr = (double)iters * CHAINS;
    printf("%-10s chains %d warps/SM %2d: %7.1f cycles per instr per warp (chain step %6.1f), %8.2f TFLOP/s\n", name, CHAINS, warps_per_sm,
           (double)c / instr, (double)c / iters, 2.0 * macs_per_instr * instr * warps_per_sm * 148 / (ms * 1e-3) / 1e12);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {1, 4, 8, 16, 32}) {
        run<1, 0>("m8n8k4", w, 256);  run<4, 0>("m8n8k4", w, 256);  run<8, 0>("m8n8k4", w, 256);
        run<1, 2>("m16n8k4", w, 512); run<4, 2>("m16n8k4", w, 512);
        run<1, 1>("m16n8k8", w, 1024); run<4, 1>("m16n8k8", w, 1024);
        run<1, 3>("dfma", w, 32);     run<4, 3>("dfma", w, 32);     run<8, 3>("dfma", w, 32);
    }
    return 0;
}
