// Common macros for the sdf-nmpc B200 kernels.
//
// Every algorithmic header (dual.cuh, model_att.cuh, qp_core.cuh) compiles both
// for sm_100a and for the host (g++ -x c++ -DSDFNMPC_HOST_EMUL): on the host a
// "warp" is a single lane, which lets the CPU-only test-suite exercise the exact
// source of the device algorithms (tests/test_host_emulation.py).  The host
// build is a development aid: it is never linked into libsdfnmpc_b200.so.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>

#if defined(__CUDACC__)
#define SDF_HD __host__ __device__ __forceinline__
#define SDF_D __device__ __forceinline__
#else
#define SDF_HD inline
#define SDF_D inline
#endif

#if defined(__CUDA_ARCH__)
#define SDF_WARP_LANES 32
#else
#define SDF_WARP_LANES 1
#endif

namespace sdfnmpc {

// One cooperating group of lanes working on one NMPC instance (a warp on the GPU).
struct Lanes {
    int lane;
    SDF_HD static constexpr int count() { return SDF_WARP_LANES; }
    SDF_HD void sync() const {
#if defined(__CUDA_ARCH__)
        __syncwarp();
#endif
    }
    SDF_HD double min(double v) const {
#if defined(__CUDA_ARCH__)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
#endif
        return v;
    }
    SDF_HD double max(double v) const {
#if defined(__CUDA_ARCH__)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
#endif
        return v;
    }
    SDF_HD double sum(double v) const {
#if defined(__CUDA_ARCH__)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
        return v;
    }
    SDF_HD int any(int p) const {
#if defined(__CUDA_ARCH__)
        return __any_sync(0xffffffffu, p);
#else
        return p;
#endif
    }
};

// Asynchronous global -> shared staging of `n` doubles (n even, both pointers 16-byte aligned), split over the lanes.
// Device: cp.async (LDGSTS, .cg = served from L2, so data this group wrote with plain stores is never read from a stale
// L1 line) - the copy proceeds while the group computes on the other buffer.  Host: plain copy.
SDF_HD void stage_async(const Lanes& L, double* dst_smem, const double* src, int n) {
#if defined(__CUDA_ARCH__)
    const unsigned d0 = (unsigned)__cvta_generic_to_shared(dst_smem);
    for (int c = L.lane; c < n / 2; c += 32)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + c * 16), "l"(src + 2 * c) : "memory");
#else
    for (int c = 0; c < n; ++c) dst_smem[c] = src[c];
    (void)L;
#endif
}
SDF_HD void stage_commit() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
SDF_HD void stage_wait_all() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}

}  // namespace sdfnmpc

#if defined(__CUDACC__)
#define SDF_CUDA_CHECK(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            fprintf(stderr, "[sdfnmpc_b200] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e), \
                    __FILE__, __LINE__, cudaGetErrorString(_e));                               \
            return (int)_e;                                                                    \
        }                                                                                      \
    } while (0)
#endif
