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

// Asynchronous global -> shared staging with the bulk-copy engine (TMA, cp.async.bulk / UBLKCP): ONE instruction issued
// by lane 0 moves a whole contiguous array (bytes a multiple of 16, both addresses 16-byte aligned) and signals an
// mbarrier owned by the lane group; the other lanes spend no issue slots on the copy.  `StagePipe` is a two-slot pipeline:
// slot s is filled with begin(s, total bytes) + copy(...) and consumed after wait(s).  The copy engine reads global
// memory through the async proxy, so publish() orders the group's earlier generic-proxy stores before it with a proxy
// fence (the caller has synchronised the lanes).  Host emulation: plain copies.
struct StagePipe {
    unsigned bar[2];      // shared-window addresses of the two mbarriers
    unsigned parity;      // bit s = phase parity slot s waits for next

    SDF_HD void init(const Lanes& L, unsigned long long* bars) {
#if defined(__CUDA_ARCH__)
        bar[0] = (unsigned)__cvta_generic_to_shared(bars);
        bar[1] = bar[0] + 8;
        parity = 0;
        if (L.lane == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar[0]) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar[1]) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        L.sync();
#else
        (void)L; (void)bars; parity = 0;
#endif
    }
    // Make the group's earlier generic-proxy global stores visible to the copy engine (call after a lane barrier, once
    // per sweep: a sweep only stages data written before it started).
    SDF_HD void publish(const Lanes& L) const {
#if defined(__CUDA_ARCH__) && defined(SDFNMPC_STAGE_BULK)
        if (L.lane == 0) asm volatile("fence.proxy.async;" ::: "memory");
#else
        (void)L;
#endif
    }
    SDF_HD void begin(const Lanes& L, int slot, unsigned bytes) const {
#if defined(__CUDA_ARCH__) && !defined(SDFNMPC_STAGE_BULK)
        (void)L; (void)slot; (void)bytes;
#elif defined(__CUDA_ARCH__)
        if (L.lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar[slot]), "r"(bytes) : "memory");
#else
        (void)L; (void)slot; (void)bytes;
#endif
    }
    // Compile-time-sized copy: per lane one base address pair, then one LDGSTS per 16-byte chunk with the chunk offset as an
    // immediate (the generic loop below spent ~7 instructions per chunk on 64-bit address arithmetic and predicates: 11 % of
    // the QP kernel's instructions).
    template <int NDOUBLES>
    SDF_HD void copy(const Lanes& L, int slot, double* dst_smem, const double* src) const {
        static_assert(NDOUBLES % 2 == 0, "16-byte units");
#if defined(__CUDA_ARCH__) && !defined(SDFNMPC_STAGE_BULK) && !defined(SDFNMPC_STAGE_GENERIC_COPY)
        const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem) + (unsigned)L.lane * 16u;
        const char* g = reinterpret_cast<const char*>(src) + L.lane * 16;
        chunk<NDOUBLES / 2, 0>(d, g, L.lane);
        (void)slot;
#else
        copy(L, slot, dst_smem, src, NDOUBLES);
#endif
    }
#if defined(__CUDA_ARCH__)
    template <int NCH, int IT>
    __device__ __forceinline__ static void chunk(unsigned d, const char* g, int lane) {
        if constexpr (IT * 32 < NCH) {
            if (IT * 32 + 31 < NCH || IT * 32 + lane < NCH)
                asm volatile("cp.async.cg.shared.global [%0+%2], [%1+%2], 16;" ::"r"(d), "l"(g), "n"(IT * 512) : "memory");
            chunk<NCH, IT + 1>(d, g, lane);
        }
    }
#endif
    SDF_HD void copy(const Lanes& L, int slot, double* dst_smem, const double* src, int n) const {
#if defined(__CUDA_ARCH__) && !defined(SDFNMPC_STAGE_BULK)
        // development variant: per-lane 16-byte cp.async (LDGSTS) instead of the bulk engine
        const unsigned d0 = (unsigned)__cvta_generic_to_shared(dst_smem);
        for (int c = L.lane; c < n / 2; c += 32)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + c * 16), "l"(src + 2 * c) : "memory");
        (void)slot;
#elif defined(__CUDA_ARCH__)
        if (L.lane == 0)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             (unsigned)__cvta_generic_to_shared(dst_smem)),
                         "l"(src), "r"(n * 8), "r"(bar[slot])
                         : "memory");
#else
        for (int c = 0; c < n; ++c) dst_smem[c] = src[c];
        (void)L; (void)slot;
#endif
    }
    // closes the batch of copies of one stage: the group is committed right behind its copies, so the later wait is a plain
    // scoreboard test (`cp.async.wait_all` commits at wait time).  Measured: QP 5.36 -> 5.34 ms, i.e. within noise.
    SDF_HD void commit(const Lanes& L) const {
#if defined(__CUDA_ARCH__) && !defined(SDFNMPC_STAGE_BULK)
        asm volatile("cp.async.commit_group;" ::: "memory");
#endif
        (void)L;
    }
    SDF_HD void wait(const Lanes& L, int slot) {
#if defined(__CUDA_ARCH__) && !defined(SDFNMPC_STAGE_BULK)
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        L.sync();
        (void)slot;
#elif defined(__CUDA_ARCH__)
        (void)L;
        const unsigned ph = (parity >> slot) & 1u;
        unsigned done = 0;
        while (!done)
            asm volatile(
                "{\n\t.reg .pred P1;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
                "selp.b32 %0, 1, 0, P1;\n\t}"
                : "=r"(done)
                : "r"(bar[slot]), "r"(ph)
                : "memory");
        parity ^= 1u << slot;
#else
        (void)L; (void)slot;
#endif
    }
};

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
