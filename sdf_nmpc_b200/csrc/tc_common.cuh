// Building blocks shared by the tcgen05 kernels (SDF network: sdf_tc_kernel.cuh, encoder convolutions: conv_tc.cuh):
// mbarrier / bulk-copy / tcgen05 wrappers, UMMA shared-memory and instruction descriptors for K-major tf32 operands in the
// 128-byte-swizzled layout, and the producer-side store of a 128 x 32 A chunk as tf32 hi | lo parts (3-MMA split
// D += A_hi B_hi + A_lo B_hi + A_hi B_lo = fp32-grade products).  Device-inline only: safe to include
// from several translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace sdfnmpc {

constexpr int TC_TM = 128;                 // rows (points / pixels) per tile = UMMA M
constexpr int TC_KC = 32;                  // K depth of one chunk (one 128-byte swizzle row of tf32)
constexpr int TC_ASTAGE_BYTES = TC_TM * TC_KC * 4 * 2;   // 32 KB: hi | lo

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
// producers: back off between probes so that spinning warps do not steal issue slots from the MMA issuer
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity) {
#if defined(TC_EXP_NO_BACKOFF)
    mbar_wait(bar, parity);
    return;
#endif
    uint32_t done = 0;
    while (true) {
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, P1;\n\t"
            "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(32);
    }
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// One lane of the converged warp (elect.sync).  Unlike a lane-id comparison, the compiler knows that exactly one lane is
// active inside `if (elect_one())`, so tcgen05 operands are broadcast once instead of through a per-lane waterfall loop
// (ELECT + 7 R2UR + branch around every MMA: measured ~150 cycles per MMA issue).
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred = 0, laneid = 0;
    asm volatile(
        "{\n\t"
        ".reg .b32 %%rx;\n\t"
        ".reg .pred %%px;\n\t"
        "elect.sync %%rx|%%px, %2;\n\t"
        "@%%px mov.s32 %1, 1;\n\t"
        "mov.s32 %0, %%rx;\n\t"
        "}"
        : "+r"(laneid), "+r"(pred)
        : "r"(0xFFFFFFFFu));
    return pred;
}
// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// 32 lanes x 8 consecutive 32-bit columns -> 8 registers per thread
__device__ __forceinline__ void tc_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// 32 lanes x 16 consecutive 32-bit columns -> 16 registers per thread
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// K-major, 128-byte swizzle (the canonical UMMA layout ((8,n),2):((8,SBO),1) o Swizzle<3,4,3>): a row is the 128 bytes
// of one 32-feature chunk, 8 rows form a 1024-byte atom whose 16-byte units are XOR-ed with the row index; SBO (between
// 8-row groups) = 1024 B, LBO unused (1); descriptor version 1 (Blackwell), layout type 2 = SWIZZLE_128B.
// The unswizzled INTERLEAVE layout measured 433 cycles per 128x256x8 MMA instead of 128 (operand fetch bound).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int n) {
    // c_format F32 (1) @4, a/b format TF32 (2) @7/@10, K-major both, N>>3 @17, M>>4 @24
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_TM >> 4) << 24);
}

// The tf32 hi / lo split is done by split_a16 BEFORE the producer waits for its stage to be free, so that only the
// eight 16-byte stores remain between "stage free" and "stage full".
// TC_SPLIT_MODE: 0 = hi and lo rounded to nearest (two cvt.rna.tf32 per element: 16 % of the SDF kernel's instructions, on a
// throttled pipe - the top line of the r2g source-level profile); 1 = hi by truncation (one LOP3; x - hi stays exact), lo
// rounded; 2 (default) = hi by truncation, lo left as the fp32 difference - the tensor core ignores the low 13 mantissa bits
// of a tf32 operand, so no cvt is needed at all.  Representation error |x - hi - lo_tf32| <= 2^-20 |x| instead of 2^-22.
// Measured (profiles/r2m_split_ab.txt; modes 0 / 1 / 2): SDF 256.6 / 262.7 / 265.9 Mpts/s with value error 1.62 / 1.66 /
// 1.76e-6 and gradient error 6.0 / 6.6 / 6.9e-6 against the float64 oracle (tolerances 1e-5 / 1e-4); encoder 32.1 / 32.6 /
// 32.8 k frames/s with latent error 1.5 / 1.7 / 2.0e-6 (tolerance 1e-4).  Weights are pre-split on the host with rounding.
#ifndef TC_SPLIT_MODE
#define TC_SPLIT_MODE 2
#endif
__device__ __forceinline__ void split1(float x, float& hi, float& lo) {
#if TC_SPLIT_MODE == 0
    hi = tf32_rna(x); lo = tf32_rna(x - hi);
#elif TC_SPLIT_MODE == 1
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); lo = tf32_rna(x - hi);
#else
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); lo = x - hi;
#endif
}
__device__ __forceinline__ void split_a16(const float* v, float* hi, float* lo) {
#pragma unroll
    for (int e = 0; e < 16; ++e) split1(v[e], hi[e], lo[e]);
}
__device__ __forceinline__ void store_a16(unsigned char* a_stage, int row, int ch, const float* hi, const float* lo) {
    unsigned char* base = a_stage + (row >> 3) * 1024 + (row & 7) * 128;
    const int sw = row & 7;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int unit = ((ch * 4 + q) ^ sw) * 16;
        *reinterpret_cast<float4*>(base + unit) = make_float4(hi[q * 4], hi[q * 4 + 1], hi[q * 4 + 2], hi[q * 4 + 3]);
        *reinterpret_cast<float4*>(base + TC_ASTAGE_BYTES / 2 + unit) = make_float4(lo[q * 4], lo[q * 4 + 1], lo[q * 4 + 2], lo[q * 4 + 3]);
    }
}

// 8-feature variants (thread <-> row, quarter `cg` of the 32-feature chunk): two 16-byte units of hi and of lo
__device__ __forceinline__ void split_a8(const float* v, float* hi, float* lo) {
#pragma unroll
    for (int e = 0; e < 8; ++e) split1(v[e], hi[e], lo[e]);
}
__device__ __forceinline__ void store_a8(unsigned char* a_stage, int row, int cg, const float* hi, const float* lo) {
    unsigned char* base = a_stage + (row >> 3) * 1024 + (row & 7) * 128;
    const int sw = row & 7;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int unit = ((cg * 2 + q) ^ sw) * 16;
        *reinterpret_cast<float4*>(base + unit) = make_float4(hi[q * 4], hi[q * 4 + 1], hi[q * 4 + 2], hi[q * 4 + 3]);
        *reinterpret_cast<float4*>(base + TC_ASTAGE_BYTES / 2 + unit) = make_float4(lo[q * 4], lo[q * 4 + 1], lo[q * 4 + 2], lo[q * 4 + 3]);
    }
}

}  // namespace tc

}  // namespace sdfnmpc
