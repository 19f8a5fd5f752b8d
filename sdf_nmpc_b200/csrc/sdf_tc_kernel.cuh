// Fused neural-SDF forward + input-gradient on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same operator as sdf_eval_kernel (sdf_kernels.cuh) - NeuralDF.forward (network/neural_df.py:91-103) and the
// reverse sweep L4CasADi's jacrev performs (gen_model.py:38-39) - but the eight hidden-layer GEMMs of a 128-point
// tile run as tcgen05.mma kind::tf32 with fp32 accumulators in TMEM.  fp32-equivalent accuracy comes from the
// 2-way TF32 split: x = hi + lo, D += A_hi B_hi + A_lo B_hi + A_hi B_lo.
//
// Structure of one CTA (160 threads, one CTA per SM, persistent over tiles of 128 points):
//   * warps 0-3 ("producers", thread t <-> point t <-> TMEM lane t): build the A operand of the next GEMM in
//     32-feature chunks directly in the UMMA canonical K-major (no-swizzle) shared-memory layout, as hi and lo
//     TF32 parts: the embedding for layer 1 / the skip connection, act(D + bias) read back from TMEM with
//     tcgen05.ld for the hidden layers, D * act' for the reverse sweep.  act' values and the partial embedding
//     gradient are parked in an L2-resident per-CTA scratch (TMEM and smem are full).
//   * warp 4, one elected thread: streams the matching 32-column weight chunk (pre-split hi|lo, pre-laid-out in the
//     canonical layout on the host) with ONE bulk TMA copy (cp.async.bulk, UBLKCP) per chunk and issues the
//     tcgen05.mma instructions; tcgen05.commit releases the 2-stage smem ring and publishes finished accumulators.
//   A chunk "program" (SdfTcOp[]) built on the host drives both roles in lock-step.
// TMEM column plan (512 columns, all allocated): forward D1 @0, D2 @256, D3 @0, D4 @128; reverse D_g3 @0,
// D_g2 @128, D_ge3 @384 (drained to scratch), D_g1 @0 (rows 0..127 of W2^T) + @384 (rest), D_ge1 @128.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "sdf_net.h"

namespace sdfnmpc {

constexpr int TC_TM = 128;                 // points per tile = UMMA M
constexpr int TC_KC = 32;                  // features per chunk
constexpr int TC_THREADS = 160;
constexpr int TC_WSTAGE_BYTES = 256 * TC_KC * 4 * 2;   // 64 KB: up to 256 weight rows, hi | lo
constexpr int TC_ASTAGE_BYTES = TC_TM * TC_KC * 4 * 2;   // 32 KB: hi | lo
constexpr int TC_SMEM_BYTES = 2 * TC_WSTAGE_BYTES + 2 * TC_ASTAGE_BYTES + 128 + 1024;

enum { TC_PROD_EMBED = 0, TC_PROD_ACT = 1, TC_PROD_MUL = 2, TC_PROD_G4 = 3 };

struct SdfTcOp {
    int prod;            // producer kind
    int src_col;         // TMEM column of the 32 source accumulator columns (prod != EMBED)
    int feat0;           // first feature of the chunk within its layer
    int layer;           // 1,2,3: hidden layer whose bias / act' scratch is used; 4 for G4
    int wait_d;          // producers wait for the previous accumulator before this chunk
    int pre_drain_ge3;   // producers first drain D_ge3 (columns 384..) to scratch
    uint32_t w_off;      // byte offset of the weight chunk in the packed blob
    uint32_t w_bytes;    // hi + lo bytes
    int n_total;         // rows of the weight chunk
    int n_piece;         // 1 or 2 MMAs along N
    int d_col[2], n_rows[2], b_row0[2];
    int first, last;     // first chunk of an accumulator (overwrite) / last (publish accumulator)
};

struct SdfTcParams {
    SdfNetDev net;
    const SdfTcOp* ops;
    int n_ops;
    const unsigned char* wblob;     // packed weight chunks
    float* scratch;                 // per CTA: c1 | c2 | c3 | ge3, each [(feat/4)][128][4]
    size_t scratch_floats_per_cta;
    int sc_off_c1, sc_off_c2, sc_off_c3, sc_off_ge3;    // float offsets inside the CTA's scratch
    const float* b1f;
    const float* b3f;
    const float* pts;
    const int* pt2inst;
    int pts_per_inst;
    int n;
    float* sdf_out;
    float* grad_out;
};

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
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// K-major, no swizzle: LBO (K direction, between the two 16-byte core-matrix columns) = 128 B,
// SBO (M/N direction, between 8-row groups) = 1024 B; descriptor version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((128u >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((1024u >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int n) {
    // c_format F32 (1) @4, a/b format TF32 (2) @7/@10, K-major both, N>>3 @17, M>>4 @24
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_TM >> 4) << 24);
}

__device__ __forceinline__ void act_pair(int act, float w0, float a, float& h, float& dh) {
    if (act == SDF_ACT_SIN) {
        float s, c;
        sincosf(w0 * a, &s, &c);
        h = s; dh = w0 * c;
    } else if (act == SDF_ACT_RELU) {
        h = fmaxf(a, 0.f); dh = a > 0.f ? 1.f : 0.f;
    } else {
        h = a > 20.f ? a : log1pf(expf(a)); dh = 1.f / (1.f + expf(-a));
    }
}

// write 32 features of this thread's row into the A stage (hi | lo), canonical K-major layout
__device__ __forceinline__ void store_a_chunk(unsigned char* a_stage, int row, const float* v) {
    unsigned char* base = a_stage + (row >> 3) * 1024 + (row & 7) * 16;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { hi[e] = tf32_rna(v[q * 4 + e]); lo[e] = tf32_rna(v[q * 4 + e] - hi[e]); }
        *reinterpret_cast<float4*>(base + q * 128) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(base + TC_ASTAGE_BYTES / 2 + q * 128) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
}

}  // namespace tc

__global__ void __launch_bounds__(TC_THREADS, 1) sdf_eval_tc_kernel(const SdfTcParams P) {
    using namespace tc;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* wst[2] = {smem, smem + TC_WSTAGE_BYTES};
    unsigned char* ast[2] = {smem + 2 * TC_WSTAGE_BYTES, smem + 2 * TC_WSTAGE_BYTES + TC_ASTAGE_BYTES};
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * TC_WSTAGE_BYTES + 2 * TC_ASTAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const uint32_t wfull[2] = {smem_u32(bars + 0), smem_u32(bars + 1)};
    const uint32_t afull[2] = {smem_u32(bars + 2), smem_u32(bars + 3)};
    const uint32_t sfree[2] = {smem_u32(bars + 4), smem_u32(bars + 5)};
    const uint32_t dfull = smem_u32(bars + 6);

    const int tid = threadIdx.x, warp = tid >> 5;
    const SdfNetDev& net = P.net;

    if (tid == 0) {
        mbar_init(wfull[0], 1); mbar_init(wfull[1], 1);
        mbar_init(afull[0], TC_TM); mbar_init(afull[1], TC_TM);
        mbar_init(sfree[0], 1); mbar_init(sfree[1], 1);
        mbar_init(dfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_tiles = (P.n + TC_TM - 1) / TC_TM;
    const int n_ops = P.n_ops;
    uint32_t chunk = 0;           // global chunk counter (stage = chunk & 1, phase = (chunk >> 1) & 1)
    uint32_t dphase = 0;          // parity of the accumulator-complete barrier

    if (warp < 4) {
        // ============================== producers: thread <-> point ==============================
        const int row = tid;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        float* scr = P.scratch + (size_t)blockIdx.x * P.scratch_floats_per_cta;
        const int DF = net.D * net.F;
        const float HALF_PI = 1.57079632679489661923f;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int gi = tile * TC_TM + row;
            const int ic = gi < P.n ? gi : P.n - 1;
            const float sx = P.pts[(size_t)ic * 3], sy = P.pts[(size_t)ic * 3 + 1], sz = P.pts[(size_t)ic * 3 + 2];
            const int inst = P.pt2inst ? P.pt2inst[ic] : ic / P.pts_per_inst;
            const float* b1 = P.b1f + (size_t)inst * net.H1;
            const float* b3 = P.b3f + (size_t)inst * net.H3;
            float d_acc = net.b5;
            for (int oi = 0; oi < n_ops; ++oi, ++chunk) {
                const SdfTcOp& op = P.ops[oi];
                const int s = chunk & 1;
                if (op.wait_d) { mbar_wait(dfull, dphase); dphase ^= 1; tc_fence_after(); }
                if (op.pre_drain_ge3) {
                    // park the skip-connection part of the embedding gradient: D_ge3 (columns 384..) -> scratch
                    for (int f0 = 0; f0 < net.Ep; f0 += 32) {
                        float v[32];
                        tc_ld32(lane_addr + 384 + f0, v);
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            *reinterpret_cast<float4*>(scr + P.sc_off_ge3 + ((size_t)((f0 >> 2) + q) * TC_TM + row) * 4) =
                                make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                    }
                    tc_fence_before();
                }
                float v[32];
                if (op.prod == TC_PROD_EMBED) {
#pragma unroll 4
                    for (int i = 0; i < 32; ++i) {
                        const int j = op.feat0 + i;
                        float e = 0.f;
                        if (j < 3) {
                            e = j == 0 ? sx : (j == 1 ? sy : sz);
                        } else if (j < net.E) {
                            const int q = (j - 3) % DF, d = q / net.F, f = q % net.F;
                            const float proj = fmaf(sz, net.dirs[2 * net.D + d], fmaf(sy, net.dirs[net.D + d], sx * net.dirs[d]));
                            float xb = proj * (float)(1 << f);
                            if (j - 3 >= DF) xb = xb + HALF_PI;
                            e = sinf(xb);
                        }
                        v[i] = e;
                    }
                } else {
                    tc_ld32(lane_addr + op.src_col, v);
                    if (op.prod == TC_PROD_ACT) {
                        const float* bias = op.layer == 1 ? b1 : (op.layer == 2 ? net.b2 : b3);
                        float* cs = scr + (op.layer == 1 ? P.sc_off_c1 : (op.layer == 2 ? P.sc_off_c2 : P.sc_off_c3));
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 bq = __ldg(reinterpret_cast<const float4*>(bias + op.feat0) + q);
                            const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
                            float c[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) act_pair(net.act, net.w0, v[q * 4 + e] + bb[e], v[q * 4 + e], c[e]);
                            *reinterpret_cast<float4*>(cs + ((size_t)((op.feat0 >> 2) + q) * TC_TM + row) * 4) = make_float4(c[0], c[1], c[2], c[3]);
                        }
                    } else if (op.prod == TC_PROD_MUL) {
                        const float* cs = scr + (op.layer == 1 ? P.sc_off_c1 : (op.layer == 2 ? P.sc_off_c2 : P.sc_off_c3));
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 c = *reinterpret_cast<const float4*>(cs + ((size_t)((op.feat0 >> 2) + q) * TC_TM + row) * 4);
                            v[q * 4] *= c.x; v[q * 4 + 1] *= c.y; v[q * 4 + 2] *= c.z; v[q * 4 + 3] *= c.w;
                        }
                    } else {   // TC_PROD_G4: output layer + seed of the reverse sweep
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float w5 = __ldg(net.W5 + op.feat0 + i);
                            float h, c;
                            act_pair(net.act, net.w0, v[i] + __ldg(net.b4 + op.feat0 + i), h, c);
                            d_acc = fmaf(w5, h, d_acc);
                            v[i] = w5 * c;
                        }
                    }
                    tc_fence_before();
                }
                mbar_wait(sfree[s], ((chunk >> 1) & 1) ^ 1);
                store_a_chunk(ast[s], row, v);
                fence_async_smem();
                mbar_arrive(afull[s]);
            }
            // ---- final: g_e = D_ge1 (+ parked skip part) ; chain rule through the embedding ----
            mbar_wait(dfull, dphase); dphase ^= 1; tc_fence_after();
            float gx = 0.f, gy = 0.f, gz = 0.f;
            for (int f0 = 0; f0 < net.Ep; f0 += 32) {
                float v[32];
                tc_ld32(lane_addr + 128 + f0, v);
                if (net.res_state) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 c = *reinterpret_cast<const float4*>(scr + P.sc_off_ge3 + ((size_t)((f0 >> 2) + q) * TC_TM + row) * 4);
                        v[q * 4] += c.x; v[q * 4 + 1] += c.y; v[q * 4 + 2] += c.z; v[q * 4 + 3] += c.w;
                    }
                }
#pragma unroll 4
                for (int i = 0; i < 32; ++i) {
                    const int j = f0 + i;
                    if (j < 3) {
                        if (j == 0) gx += v[i]; else if (j == 1) gy += v[i]; else gz += v[i];
                    } else if (j < net.E) {
                        const int q = (j - 3) % DF, d = q / net.F, f = q % net.F;
                        const float dx = net.dirs[d], dy = net.dirs[net.D + d], dz = net.dirs[2 * net.D + d];
                        const float proj = fmaf(sz, dz, fmaf(sy, dy, sx * dx));
                        const float sc = (float)(1 << f);
                        float xb = proj * sc;
                        if (j - 3 >= DF) xb = xb + HALF_PI;
                        const float t = v[i] * cosf(xb) * sc;
                        gx = fmaf(t, dx, gx); gy = fmaf(t, dy, gy); gz = fmaf(t, dz, gz);
                    }
                }
            }
            tc_fence_before();
            if (gi < P.n) {
                if (P.sdf_out) P.sdf_out[gi] = d_acc;
                if (P.grad_out) { P.grad_out[(size_t)gi * 3] = gx; P.grad_out[(size_t)gi * 3 + 1] = gy; P.grad_out[(size_t)gi * 3 + 2] = gz; }
            }
        }
    } else if (tid == 128) {
        // ============================== weight streamer + MMA issuer (one thread) ==============================
        const uint64_t n_chunks_total = (uint64_t)((n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * n_ops;
        if (n_chunks_total > 0) {
            const SdfTcOp& o0 = P.ops[0];
            mbar_expect_tx(wfull[0], o0.w_bytes);
            bulk_g2s(smem_u32(wst[0]), P.wblob + o0.w_off, o0.w_bytes, wfull[0]);
        }
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int oi = 0; oi < n_ops; ++oi, ++chunk) {
                const SdfTcOp& op = P.ops[oi];
                const int s = chunk & 1;
                // prefetch the next chunk's weights into the other stage once its previous MMAs have retired
                if ((uint64_t)chunk + 1 < n_chunks_total) {
                    const SdfTcOp& nx = P.ops[oi + 1 < n_ops ? oi + 1 : 0];
                    const uint32_t c1 = chunk + 1;
                    mbar_wait(sfree[c1 & 1], ((c1 >> 1) & 1) ^ 1);
                    mbar_expect_tx(wfull[c1 & 1], nx.w_bytes);
                    bulk_g2s(smem_u32(wst[c1 & 1]), P.wblob + nx.w_off, nx.w_bytes, wfull[c1 & 1]);
                }
                mbar_wait(wfull[s], (chunk >> 1) & 1);
                mbar_wait(afull[s], (chunk >> 1) & 1);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(ast[s]), a_lo = a_hi + TC_ASTAGE_BYTES / 2;
                const uint32_t w_hi = smem_u32(wst[s]), w_lo = w_hi + (uint32_t)op.n_total * 128u;
#pragma unroll 1
                for (int ks = 0; ks < TC_KC / 8; ++ks) {
#pragma unroll 1
                    for (int sp = 0; sp < 3; ++sp) {
                        const uint32_t a_base = (sp == 1 ? a_lo : a_hi) + ks * 256;
                        const uint32_t w_base = (sp == 2 ? w_lo : w_hi) + ks * 256;
                        for (int pc = 0; pc < op.n_piece; ++pc) {
                            const uint32_t acc = (op.first && ks == 0 && sp == 0) ? 0u : 1u;
                            tc_mma_tf32(tmem_base + (uint32_t)op.d_col[pc], make_desc(a_base),
                                        make_desc(w_base + (uint32_t)(op.b_row0[pc] >> 3) * 1024u), make_idesc(op.n_rows[pc]), acc);
                        }
                    }
                }
                tc_commit(sfree[s]);
                if (op.last) tc_commit(dfull);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace sdfnmpc
