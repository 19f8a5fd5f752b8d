// Development microbenchmark: issue rate of tcgen05.mma kind::tf32 (M = 128, N = n, K = 8, both operands in shared memory,
// 128-byte-swizzled K-major) from one elected lane, alone and with concurrent shared-memory traffic from the other
// warps (the A-stage stores / weight TMA of the real kernels).  Prints cycles per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I sdf_nmpc_b200/csrc tools/mma_rate.cu -o tools/bin/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace sdfnmpc;
using namespace sdfnmpc::tc;

constexpr int THREADS = 9 * 32;          // 8 traffic warps + 1 issuer warp
constexpr int SMEM = 2 * 65536 + 2 * 32768 + 1024 + 1024;

__global__ void __launch_bounds__(THREADS, 1) mma_rate_kernel(int n, int n_mma, int mode, const unsigned char* gsrc, long long* out) {
    extern __shared__ __align__(1024) unsigned char raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    unsigned char* wst = smem;                       // 2 x 64 KB
    unsigned char* ast = smem + 2 * 65536;           // 2 x 32 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * 65536 + 2 * 32768);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 4);
    volatile int* stop = reinterpret_cast<volatile int*>(bars + 6);
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t done = smem_u32(bars), tbar = smem_u32(bars + 1);
    for (int i = tid; i < (2 * 65536 + 2 * 32768) / 4; i += THREADS) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 1023);
    if (tid == 0) { mbar_init(done, 1); mbar_init(tbar, 1); *stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *slot, 0);
    if (warp == 8) {
        const uint32_t idesc = make_idesc(n);
        const uint64_t ad = make_desc(smem_u32(ast)), al = make_desc(smem_u32(ast) + 16384);
        const uint64_t bd = make_desc(smem_u32(wst)), bl = make_desc(smem_u32(wst) + n * 128);
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int i = 0; i < n_mma / 12; ++i) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    tc_mma_tf32(tmem, ad + 2 * ks, bd + 2 * ks, idesc, 1u);
                    tc_mma_tf32(tmem, al + 2 * ks, bd + 2 * ks, idesc, 1u);
                    tc_mma_tf32(tmem, ad + 2 * ks, bl + 2 * ks, idesc, 1u);
                }
            }
            tc_commit(done);
            t1 = clock64();
        }
        __syncwarp();
        mbar_wait(done, 0);
        const long long t2 = clock64();
        if (tid == 8 * 32) { *stop = 1; }
        if (elect_one() && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    } else {
        // traffic generators on the second A / W stages (never read by the MMAs)
        unsigned char* dst = ast + 32768;
        const int row = (warp & 3) * 32 + (tid & 31), ch = warp >> 2;
        float hi[16], lo[16];
        for (int i = 0; i < 16; ++i) { hi[i] = 1.f + i; lo[i] = 0.5f * i; }
        uint32_t it = 0;
        while (!*stop) {
            if (mode & 1) {              // A-stage stores + proxy fence, as the producers do (32 KB per 256 threads)
                store_a16(dst, row, ch, hi, lo);
                fence_async_smem();
            }
            if ((mode & 2) && tid == 0) {   // 64 KB bulk copies into the second W stage
                mbar_expect_tx(tbar, 65536u);
                bulk_g2s(smem_u32(wst + 65536), gsrc, 65536u, tbar);
                mbar_wait(tbar, it & 1);
            }
            if (mode & 4) __nanosleep(64);
            ++it;
            hi[it & 15] += 1.f;
        }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    unsigned char* g; long long* out; long long h[2];
    cudaMalloc(&g, 1 << 20); cudaMemset(g, 0, 1 << 20); cudaMalloc(&out, 16);
    cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    const int n_mma = 12 * 400;
    for (int n : {256, 128, 64})
        for (int mode : {0, 1, 2, 3, 5}) {
            for (int rep = 0; rep < 2; ++rep) mma_rate_kernel<<<148, THREADS, SMEM>>>(n, n_mma, mode, g, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            printf("N=%3d mode=%d (1=A stores, 2=W TMA, 4=sleep)  issue %.1f cyc/MMA  complete %.1f cyc/MMA\n", n, mode,
                   (double)h[0] / n_mma, (double)h[1] / n_mma);
        }
    return 0;
}
