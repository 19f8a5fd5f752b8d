// Fused neural-SDF forward + input-gradient on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same operator as sdf_eval_kernel (sdf_kernels.cuh) - NeuralDF.forward (network/neural_df.py:91-103) and the
// reverse sweep L4CasADi's jacrev performs (gen_model.py:38-39) - but the eight hidden-layer GEMMs of a 128-point
// tile run as tcgen05.mma kind::tf32 with fp32 accumulators in TMEM.  fp32-equivalent accuracy comes from the
// 2-way TF32 split: x = hi + lo, D += A_hi B_hi + A_lo B_hi + A_hi B_lo.
//
// Structure of one CTA (576 threads = 18 warps, one CTA per SM, persistent over tiles of 128 points):
//   * warps 0-15 ("producers"; all build every chunk: thread <-> (point = TMEM lane, 8 of the chunk's 32 features)):
//     build the A operand of the next GEMM in 32-feature chunks in the 128-byte-swizzled K-major UMMA layout, as hi and lo
//     TF32 parts: the embedding for layer 1 / the skip connection, act(D + bias) read back from TMEM with tcgen05.ld for
//     the hidden layers, D * act' for the reverse sweep.  act' values and the partial embedding gradient are parked in an
//     L2-resident per-CTA scratch (TMEM and smem are full).  Chunk c uses A stage (c & 1); the leading embedding chunks of
//     the NEXT tile are built around the current tile's epilogue.
//   * warp 16 (MMA issuer): waits for operands and issues the tcgen05.mma from one elected lane on warp-uniform operands;
//     tcgen05.commit releases the 2-stage smem ring and publishes finished accumulators.
//   * warp 17 (weight streamer): ONE bulk copy (cp.async.bulk, UBLKCP) per 32-column weight chunk (pre-split hi|lo,
//     pre-laid-out on the host) into the stage whose previous MMAs have retired.
//   A chunk "program" (SdfTcOp[]) built on the host drives all roles in lock-step.
// TMEM column plan (512 columns, all allocated): forward D1 @0, D2 @256, D3 @0, D4 @128; reverse D_g3 @0,
// D_g2 @128, D_ge3 @384 (drained to scratch), D_g1 @0 (rows 0..127 of W2^T) + @384 (rest), D_ge1 @256.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "sdf_net.h"
#include "tc_common.cuh"

namespace sdfnmpc {

constexpr int TC_PROD_WARPS = 16;            // producer warps: 4 TMEM lane quarters x 4 column groups of 8 features
constexpr int TC_MAX_ACC = 8;                // accumulator-complete barriers per tile
constexpr int TC_PROD_THREADS = TC_PROD_WARPS * 32;
constexpr int TC_THREADS = TC_PROD_THREADS + 64;          // + MMA issuer warp + weight-streaming warp
constexpr int TC_WSTAGE_BYTES = 256 * TC_KC * 4 * 2;   // 64 KB: up to 256 weight rows, hi | lo
// Timing-only ablations for development (results are wrong): -DTC_EXP_NO_STORE (producers skip the A-stage stores),
// -DTC_EXP_ONE_MMA (hi*hi only), -DTC_EXP_NO_TMA (16-byte weight copies).  Measured at 1e6 points: base 151 Mpts/s, no
// store 191, one MMA 181, no TMA 159, all three 278 - no single limiter; the producer skeleton alone costs 136 k cycles
// per tile against 57 k of MMA floor.
constexpr int TC_MAX_OPS = 80;
constexpr int TC_AUX_BYTES = 128 /*barriers*/ + 128 * 16 /*embedding table*/ + 4 * 128 * 4 * 4 /*row partials*/ + TC_MAX_OPS * 80 /*op program*/;
constexpr int TC_SMEM_BYTES = 2 * TC_WSTAGE_BYTES + 2 * TC_ASTAGE_BYTES + TC_AUX_BYTES + 1024;

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
    int d_need;          // accumulator completion (index within the tile) the producers need before reading TMEM, -1: none
    int d_commit;        // accumulator completion this chunk publishes (last == 1), else -1
};

struct SdfTcParams {
    SdfNetDev net;
    const SdfTcOp* ops;             // op program in global memory (producers copy it to shared memory)
    SdfTcOp opsv[TC_MAX_OPS];       // the same program by value: the MMA-issuing warp reads it from the constant bank, so
                                    // descriptors and TMEM columns are warp-uniform values (uniform registers, no R2UR)
    int n_ops;
    const unsigned char* wblob;     // packed weight chunks
    float* scratch;                 // per CTA: c1 | c2 | c3 | ge3, each [(feat/4)][128][4]
    size_t scratch_floats_per_cta;
    int sc_off_c1, sc_off_c2, sc_off_c3, sc_off_ge3;    // float offsets inside the CTA's scratch
    int sc_off_emb;                 // embedding cache [tile parity][value | cos][Ep / 4][128][4]
    const float* b1f;
    const float* b3f;
    const float* pts;
    const int* pt2inst;
    int pts_per_inst;
    int n;
    float* sdf_out;
    float* grad_out;
    int n_acc;                      // accumulator completions per tile
    long long* trace;               // optional [n_ops][8] clock64 samples of CTA 0's second tile (development aid)
};

namespace tc {

// sin and cos to ~1 ulp for |x| < 8192: 3-term Cody-Waite reduction by pi/2 + the cephes sinf/cosf minimax
// polynomials on [-pi/4, pi/4] (22 instructions instead of the generic sincosf; no MUFU approximation involved)
__device__ __forceinline__ void fast_sincos(float x, float& s, float& c) {
    const float q = rintf(x * 0.636619772367581343f);
    const int n = (int)q;
    // pi/2 = C1 + C2 + C3 with C1 = float(pi/2); each fmaf rounds the exact x - q*C once (verified in numpy: max abs
    // error 7e-8 for |x| <= 1000, the same as a correctly rounded float32 libm)
    float r = fmaf(q, -1.5707963705062866f, x);
    r = fmaf(q, 4.371138828673793e-08f, r);
    r = fmaf(q, 1.7763568394002505e-15f, r);
    const float r2 = r * r;
    float ps = fmaf(fmaf(fmaf(-1.9515295891e-4f, r2, 8.3321608736e-3f), r2, -1.6666654611e-1f), r2 * r, r);
    float pc = fmaf(fmaf(fmaf(fmaf(2.443315711809948e-5f, r2, -1.388731625493765e-3f), r2, 4.166664568298827e-2f), r2, -0.5f), r2, 1.0f);
    const bool sw = n & 1;
    float ss = sw ? pc : ps, cc = sw ? ps : pc;
    s = (n & 2) ? -ss : ss;
    c = ((n + 1) & 2) ? -cc : cc;
}
// SIREN activations: the special-function unit (MUFU.SIN / MUFU.COS after the hardware range reduction) instead of the
// polynomial pair.  Arguments w0*a of a SIREN stay within a few periods (|w0 a| <= 6.1 on the synthetic networks), where
// the unit's absolute error is ~2^-21 + |x| 2^-24; propagated through the network this moves the SDF value by < 1e-6 and
// its gradient by ~1e-6 (relative, max-norm; tolerances 1e-5 / 1e-4, tests/test_gpu_sdf.py).  The embedding keeps the
// ~1-ulp fast_sincos: its arguments reach 2^(F-1) * |s| ~ 100 rad.  -DSDFNMPC_SDF_POLY_SINCOS selects the polynomial pair.
__device__ __forceinline__ void act_sincos(float x, float& s, float& c) {
#if defined(SDFNMPC_SDF_POLY_SINCOS)
    fast_sincos(x, s, c);
#else
    s = __sinf(x);
    c = __cosf(x);
#endif
}

__device__ __forceinline__ void act_pair(int act, float w0, float a, float& h, float& dh) {
    if (act == SDF_ACT_SIN) {
        float s, c;
        act_sincos(w0 * a, s, c);
        h = s; dh = w0 * c;
    } else if (act == SDF_ACT_RELU) {
        h = fmaxf(a, 0.f); dh = a > 0.f ? 1.f : 0.f;
    } else {
        h = a > 20.f ? a : log1pf(expf(a)); dh = 1.f / (1.f + expf(-a));
    }
}

// write this thread's 16 features (column half `ch` of the 32-feature chunk) of its row into the A stage (hi | lo),
// 128-byte-swizzled K-major layout: 8-row atoms of 1024 B, row = 128 B, 16-byte unit index XOR row-in-atom
}  // namespace tc

__global__ void __launch_bounds__(TC_THREADS, 1) sdf_eval_tc_kernel(const SdfTcParams P) {
    using namespace tc;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    // stage addresses and barriers are computed from the stage index (a dynamically indexed local array costs a
    // local-memory load per use - measured in the convolution kernel's issuer)
    unsigned char* wst0 = smem;                                   // 2 x TC_WSTAGE_BYTES
    unsigned char* ast0 = smem + 2 * TC_WSTAGE_BYTES;             // 2 x TC_ASTAGE_BYTES
    unsigned char* aux = smem + 2 * TC_WSTAGE_BYTES + 2 * TC_ASTAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(aux);                   // 0-1 wfull, 2-3 afull, 4-5 sfree, 6-13 dfull ring
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);
    float4* etab = reinterpret_cast<float4*>(aux + 128);                 // [Ep] (dir * 2^f, phase)  / identity rows
    float* part = reinterpret_cast<float*>(aux + 128 + 128 * 16);        // [4 column groups][128 rows][4]
    SdfTcOp* ops_s = reinterpret_cast<SdfTcOp*>(aux + 128 + 128 * 16 + 4 * 128 * 4 * 4);   // op program
    const uint32_t wfull0 = smem_u32(bars + 0), afull0 = smem_u32(bars + 2), sfree0 = smem_u32(bars + 4);   // + 8 * stage
    const uint32_t dfull0 = smem_u32(bars + 6);                          // dfull[q] = dfull0 + 8 q

    // warp index through a lane-0 shuffle: the compiler then knows it is warp-uniform, which keeps the role branches and
    // everything computed inside the issuer role on the uniform datapath
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const SdfNetDev& net = P.net;
    const int DF = net.D * net.F;
    const float HALF_PI = 1.57079632679489661923f;

    if (tid == 0) {
        mbar_init(wfull0, 1); mbar_init(wfull0 + 8, 1);
        mbar_init(afull0, TC_PROD_THREADS); mbar_init(afull0 + 8, TC_PROD_THREADS);
        mbar_init(sfree0, 1); mbar_init(sfree0 + 8, 1);
        for (int q = 0; q < TC_MAX_ACC; ++q) mbar_init(dfull0 + 8 * q, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // embedding table: feature j >= 3 is sin(sx*t.x + sy*t.y + sz*t.z + t.w) with t = (dir_d * 2^f, phase); scaling the
    // direction by the power of two is exact, so the argument is bit-identical to (s . dir) * 2^f of the reference
    for (int j = tid; j < net.Ep; j += TC_THREADS) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j >= 3 && j < net.E) {
            const int q = (j - 3) % DF, d = q / net.F, f = q % net.F;
            const float sc = (float)(1 << f);
            t = make_float4(net.dirs[d] * sc, net.dirs[net.D + d] * sc, net.dirs[2 * net.D + d] * sc, (j - 3 >= DF) ? HALF_PI : 0.f);
        }
        etab[j] = t;
    }
    for (int i = tid; i < P.n_ops * (int)(sizeof(SdfTcOp) / 4); i += TC_THREADS)
        reinterpret_cast<int*>(ops_s)[i] = reinterpret_cast<const int*>(P.ops)[i];
    if (warp == TC_PROD_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    const int n_tiles = (P.n + TC_TM - 1) / TC_TM;
    const int n_ops = P.n_ops;
    uint32_t chunk = 0;           // global chunk counter (stage = chunk & 1, phase = (chunk >> 1) & 1)

    if (warp < TC_PROD_WARPS) {
        // ===== producers: all 16 warps build EVERY chunk together.  thread <-> (point = TMEM lane, 8 of the chunk's 32
        // features): warp w reads TMEM lane quarter w % 4 and owns feature quarter w / 4.  Chunks alternate between the two
        // A stages, so building chunk c + 1 overlaps the MMAs of chunk c; with half the work per thread the first chunk
        // after a layer boundary is ready in half the time (the drain at a boundary is what a tile loses most cycles to).
        const int lq = warp & 3, cg = warp >> 2;
        const int row = lq * 32 + (tid & 31);
        const uint32_t lane_addr = tmem_base + ((uint32_t)(lq * 32) << 16);
        float* scr = P.scratch + (size_t)blockIdx.x * P.scratch_floats_per_cta;
        uint32_t tile_it = 0;
        // Software pipelining across tiles: the first two chunks of a tile (embedding chunks: no TMEM input) are built
        // for the NEXT tile before the current tile's result is drained, so its MMAs start during the epilogue (the
        // accumulator D_ge1 sits at columns 256.., clear of the next D1).
        // (Two chunks fit the two A stages before the epilogue; the remaining leading embedding chunks follow in the middle of
        // the epilogue, once the TMEM reads are done and the first two stages have been consumed.)
        int n_lead = 0;                                        // leading embedding chunks of a tile
        while (n_lead < n_ops && n_lead < 4 && ops_s[n_lead].prod == TC_PROD_EMBED) ++n_lead;
        const int n_pre = (n_ops % 2 == 0 && n_lead >= 2) ? 2 : 0;
        const int n_pre2 = n_pre ? n_lead : 0;                 // total pre-built chunks once the epilogue is through
        int pre_done = 0;
        auto publish = [&](uint32_t chunk_id, const float* v) {
            float vhi[8], vlo[8];
            split_a8(v, vhi, vlo);
            const uint32_t s = chunk_id & 1;
            mbar_wait_backoff(sfree0 + 8 * s, ((chunk_id >> 1) & 1) ^ 1);
#if !defined(TC_EXP_NO_STORE)
            store_a8(ast0 + s * TC_ASTAGE_BYTES, row, cg, vhi, vlo);
#endif
            fence_async_smem();
            mbar_arrive(afull0 + 8 * s);
        };
        // The embedding of a point is needed three times per tile - input of layer 1, skip connection into layer 3, and its
        // cosines in the chain rule of the epilogue.  fast_sincos is ~30 instructions per feature with the argument and the
        // select, which made the three evaluations 39 % of the producers' instructions; now the first one parks values and
        // cosines in the CTA's L2 scratch (by tile parity: the next tile's leading chunks are built before this tile's epilogue)
        // and the other two read them back.  -DTC_NO_EMB_CACHE restores the recomputation.
        auto emb_slot = [&](uint32_t par, int what, int f8) {
            return reinterpret_cast<float4*>(scr + P.sc_off_emb + (size_t)(par * 2 + what) * net.Ep * TC_TM + ((size_t)(f8 >> 2) * TC_TM + row) * 4);
        };
        auto embed8 = [&](int f8, float px, float py, float pz, float* v, int layer, uint32_t par) {
#if !defined(TC_NO_EMB_CACHE)
            if (layer == 3) {
                const float4* src = emb_slot(par, 0, f8);
                const float4 a = src[0], b = src[TC_TM];
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
                return;
            }
#endif
            float cv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int j = f8 + i;
                const float4 t = etab[j];
                const float arg = fmaf(pz, t.z, fmaf(py, t.y, px * t.x)) + t.w;
                float sn, cs;
                fast_sincos(arg, sn, cs);
                v[i] = j < 3 ? (j == 0 ? px : (j == 1 ? py : pz)) : (j < net.E ? sn : 0.f);
                cv[i] = cs;
            }
#if !defined(TC_NO_EMB_CACHE)
            float4* dv = emb_slot(par, 0, f8);
            float4* dc = emb_slot(par, 1, f8);
            dv[0] = make_float4(v[0], v[1], v[2], v[3]); dv[TC_TM] = make_float4(v[4], v[5], v[6], v[7]);
            dc[0] = make_float4(cv[0], cv[1], cv[2], cv[3]); dc[TC_TM] = make_float4(cv[4], cv[5], cv[6], cv[7]);
#endif
        };
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_it) {
            const int gi = tile * TC_TM + row;
            const int ic = gi < P.n ? gi : P.n - 1;
            const float sx = __ldg(P.pts + (size_t)ic * 3), sy = __ldg(P.pts + (size_t)ic * 3 + 1), sz = __ldg(P.pts + (size_t)ic * 3 + 2);
            const int inst = P.pt2inst ? __ldg(P.pt2inst + ic) : ic / P.pts_per_inst;
            const float* b1 = P.b1f + (size_t)inst * net.H1;
            const float* b3 = P.b3f + (size_t)inst * net.H3;
            const uint32_t dpar = tile_it & 1;                  // every dfull[q] completes exactly once per tile
            float d_acc = 0.f;
            // clock64 trace of CTA 0's second tile: compiled in only with -DSDFNMPC_TRACE_BUILD (the three tests per chunk
            // were 4 % of the kernel's instructions in the r2g source-level profile)
#if defined(SDFNMPC_TRACE_BUILD)
            const bool tr = P.trace && blockIdx.x == 0 && tile == (int)gridDim.x && tid == 0;
#else
            constexpr bool tr = false;
#endif
            for (int oi = 0; oi < n_ops; ++oi, ++chunk) {
                if (oi < pre_done) continue;                       // already built during the previous tile's epilogue
                const SdfTcOp op = ops_s[oi];
                const int f8 = op.feat0 + cg * 8;                  // this thread's first feature within the layer
                if (tr) P.trace[oi * 8 + 0] = clock64();
                if (op.d_need >= 0) { mbar_wait_backoff(dfull0 + 8 * op.d_need, dpar); tc_fence_after(); }
                if (op.pre_drain_ge3) {
                    // park the skip-connection part of the embedding gradient: D_ge3 (columns 384..) -> scratch
                    for (int f0 = 0; f0 < net.Ep; f0 += 32) {
                        float v[8];
                        tc_ld8(lane_addr + 384 + f0 + cg * 8, v);
                        float4* dst = reinterpret_cast<float4*>(scr + P.sc_off_ge3 + ((size_t)((f0 + cg * 8) >> 2) * TC_TM + row) * 4);
#pragma unroll
                        for (int q = 0; q < 2; ++q) dst[q * TC_TM] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                    }
                    tc_fence_before();
                }
                float v[8];
                if (op.prod == TC_PROD_EMBED) {
                    embed8(f8, sx, sy, sz, v, op.layer, tile_it & 1);
                } else {
                    if (op.prod == TC_PROD_ACT) {
                        const float* bias = (op.layer == 1 ? b1 : (op.layer == 2 ? net.b2 : b3)) + f8;
                        float bb[8];
#pragma unroll
                        for (int q = 0; q < 2; ++q) {          // issue the bias loads before waiting on TMEM
                            const float4 bq = __ldg(reinterpret_cast<const float4*>(bias) + q);
                            bb[q * 4] = bq.x; bb[q * 4 + 1] = bq.y; bb[q * 4 + 2] = bq.z; bb[q * 4 + 3] = bq.w;
                        }
                        tc_ld8(lane_addr + op.src_col + cg * 8, v);
                        float* cs = scr + (op.layer == 1 ? P.sc_off_c1 : (op.layer == 2 ? P.sc_off_c2 : P.sc_off_c3)) + ((size_t)(f8 >> 2) * TC_TM + row) * 4;
                        float c[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) act_pair(net.act, net.w0, v[e] + bb[e], v[e], c[e]);
#pragma unroll
                        for (int q = 0; q < 2; ++q) reinterpret_cast<float4*>(cs)[q * TC_TM] = make_float4(c[q * 4], c[q * 4 + 1], c[q * 4 + 2], c[q * 4 + 3]);
                    } else if (op.prod == TC_PROD_MUL) {
                        const float* cs = scr + (op.layer == 1 ? P.sc_off_c1 : (op.layer == 2 ? P.sc_off_c2 : P.sc_off_c3)) + ((size_t)(f8 >> 2) * TC_TM + row) * 4;
                        float4 c[2];
#pragma unroll
                        for (int q = 0; q < 2; ++q) c[q] = reinterpret_cast<const float4*>(cs)[q * TC_TM];
                        tc_ld8(lane_addr + op.src_col + cg * 8, v);
#pragma unroll
                        for (int q = 0; q < 2; ++q) { v[q * 4] *= c[q].x; v[q * 4 + 1] *= c[q].y; v[q * 4 + 2] *= c[q].z; v[q * 4 + 3] *= c[q].w; }
                    } else {   // TC_PROD_G4: output layer + seed of the reverse sweep
                        tc_ld8(lane_addr + op.src_col + cg * 8, v);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float w5 = __ldg(net.W5 + f8 + i);
                            float h, c;
                            act_pair(net.act, net.w0, v[i] + __ldg(net.b4 + f8 + i), h, c);
                            d_acc = fmaf(w5, h, d_acc);
                            v[i] = w5 * c;
                        }
                    }
                    tc_fence_before();
                }
                if (tr) P.trace[oi * 8 + 1] = clock64();
                publish(chunk, v);
                if (tr) P.trace[oi * 8 + 3] = clock64();
            }
            pre_done = 0;
            const bool has_next = n_pre && tile + (int)gridDim.x < n_tiles;
            float nx = 0.f, ny = 0.f, nz = 0.f;
            if (has_next) {
                const int ngi = (tile + (int)gridDim.x) * TC_TM + row;
                const int nic = ngi < P.n ? ngi : P.n - 1;
                nx = __ldg(P.pts + (size_t)nic * 3); ny = __ldg(P.pts + (size_t)nic * 3 + 1); nz = __ldg(P.pts + (size_t)nic * 3 + 2);
                for (int q = 0; q < n_pre; ++q) {
                    float v[8];
                    embed8(ops_s[q].feat0 + cg * 8, nx, ny, nz, v, 1, (tile_it + 1) & 1);
                    publish(chunk + q, v);
                }
                pre_done = n_pre;
            }
            // ---- final: g_e = D_ge1 (+ parked skip part) ; chain rule through the embedding.  Every thread takes its 8-feature
            // quarter of every 32-feature block, so the four quarters cover a row exactly once. ----
            mbar_wait_backoff(dfull0 + 8 * (P.n_acc - 1), dpar); tc_fence_after();
            float gx = 0.f, gy = 0.f, gz = 0.f;
            for (int f0 = 0; f0 < net.Ep; f0 += 32) {
                float v[8];
                tc_ld8(lane_addr + 256 + f0 + cg * 8, v);
                const int f8 = f0 + cg * 8;
                if (net.res_state) {
                    const float* gs = scr + P.sc_off_ge3 + ((size_t)(f8 >> 2) * TC_TM + row) * 4;
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const float4 c = reinterpret_cast<const float4*>(gs)[q * TC_TM];
                        v[q * 4] += c.x; v[q * 4 + 1] += c.y; v[q * 4 + 2] += c.z; v[q * 4 + 3] += c.w;
                    }
                }
#if !defined(TC_NO_EMB_CACHE)
                float cv[8];
                {
                    const float4* src = emb_slot(tile_it & 1, 1, f8);
                    const float4 a = src[0], b = src[TC_TM];
                    cv[0] = a.x; cv[1] = a.y; cv[2] = a.z; cv[3] = a.w; cv[4] = b.x; cv[5] = b.y; cv[6] = b.z; cv[7] = b.w;
                }
#endif
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int j = f8 + i;
                    const float4 t = etab[j];
#if !defined(TC_NO_EMB_CACHE)
                    const float cs = cv[i];
#else
                    const float arg = fmaf(sz, t.z, fmaf(sy, t.y, sx * t.x)) + t.w;
                    float sn, cs;
                    fast_sincos(arg, sn, cs);
#endif
                    const float w = (j >= 3 && j < net.E) ? v[i] * cs : 0.f;        // d/darg sin(arg) = cos(arg)
                    gx = fmaf(w, t.x, gx); gy = fmaf(w, t.y, gy); gz = fmaf(w, t.z, gz);
                    if (j == 0) gx += v[i];
                    if (j == 1) gy += v[i];
                    if (j == 2) gz += v[i];
                }
            }
            tc_fence_before();
            if (has_next) {
                for (int q = n_pre; q < n_pre2; ++q) {
                    float v[8];
                    embed8(ops_s[q].feat0 + cg * 8, nx, ny, nz, v, 1, (tile_it + 1) & 1);
                    publish(chunk + q, v);
                }
                pre_done = n_pre2;
            }
            // combine the four feature-quarter partials of each row
            reinterpret_cast<float4*>(part)[cg * TC_TM + row] = make_float4(gx, gy, gz, d_acc);
            asm volatile("bar.sync 1, %0;" ::"n"(TC_PROD_THREADS) : "memory");
            if (cg == 0 && gi < P.n) {
                const float4 p0 = reinterpret_cast<const float4*>(part)[row], p1 = reinterpret_cast<const float4*>(part)[TC_TM + row];
                const float4 p2 = reinterpret_cast<const float4*>(part)[2 * TC_TM + row], p3 = reinterpret_cast<const float4*>(part)[3 * TC_TM + row];
                if (P.sdf_out) P.sdf_out[gi] = ((p0.w + p1.w) + (p2.w + p3.w)) + net.b5;
                if (P.grad_out) {
                    P.grad_out[(size_t)gi * 3] = (p0.x + p1.x) + (p2.x + p3.x);
                    P.grad_out[(size_t)gi * 3 + 1] = (p0.y + p1.y) + (p2.y + p3.y);
                    P.grad_out[(size_t)gi * 3 + 2] = (p0.z + p1.z) + (p2.z + p3.z);
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(TC_PROD_THREADS) : "memory");
        }
    } else if (warp == TC_PROD_WARPS + 1) {
        // ============================== weight streamer ==============================
        // One bulk copy per chunk into the stage its MMAs will read, as soon as the stage's previous MMAs have retired.
        // (Issued from the MMA warp these copies cost that thread 500-800 cycles each - clock64 trace - during which no MMA
        // was queued.)
        const uint32_t leader = elect_one();
        const uint64_t n_chunks_total = (uint64_t)((n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * n_ops;
        int oi = 0;
        for (uint64_t c = 0; c < n_chunks_total; ++c) {
            const SdfTcOp& op = P.opsv[oi];
            const uint32_t q = (uint32_t)(c & 1);
            mbar_wait(sfree0 + 8 * q, (uint32_t)(((c >> 1) & 1) ^ 1));
            if (leader) {
#if defined(TC_EXP_NO_TMA)
                mbar_expect_tx(wfull0 + 8 * q, 16);
                bulk_g2s(smem_u32(wst0 + q * TC_WSTAGE_BYTES), P.wblob + op.w_off, 16, wfull0 + 8 * q);
#else
                mbar_expect_tx(wfull0 + 8 * q, op.w_bytes);
                bulk_g2s(smem_u32(wst0 + q * TC_WSTAGE_BYTES), P.wblob + op.w_off, op.w_bytes, wfull0 + 8 * q);
#endif
            }
            __syncwarp();
            if (++oi == n_ops) oi = 0;
        }
    } else {
        // ============================== MMA issuer ==============================
        // The whole warp runs this loop converged, on warp-uniform values (op program from the constant bank), so that
        // descriptors live in uniform registers; one elected lane executes the MMA / commit instructions.
        const uint32_t leader = elect_one();
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#if defined(SDFNMPC_TRACE_BUILD)
            const bool tr = P.trace && blockIdx.x == 0 && tile == (int)gridDim.x;
#else
            constexpr bool tr = false;
#endif
            for (int oi = 0; oi < n_ops; ++oi, ++chunk) {
                const SdfTcOp& op = P.opsv[oi];
                const uint32_t s = chunk & 1;
#if !defined(TC_TRACE_PROD)
                if (tr && leader) { P.trace[oi * 8 + 4] = clock64(); P.trace[oi * 8 + 5] = clock64(); }
#endif
                mbar_wait(wfull0 + 8 * s, (chunk >> 1) & 1);
#if !defined(TC_TRACE_PROD)
                if (tr && leader) P.trace[oi * 8 + 6] = clock64();
#endif
                mbar_wait(afull0 + 8 * s, (chunk >> 1) & 1);
#if !defined(TC_TRACE_PROD)
                if (tr && leader) P.trace[oi * 8 + 7] = clock64();
#endif
                tc_fence_after();
                // descriptors are built once per chunk; per MMA only the start-address field advances (+32 B = +2 units
                // per K step of 8 tf32)
                const uint32_t a_hi = smem_u32(ast0) + s * TC_ASTAGE_BYTES, w_hi = smem_u32(wst0) + s * TC_WSTAGE_BYTES;
                const uint64_t ad_hi = make_desc(a_hi), ad_lo = make_desc(a_hi + TC_ASTAGE_BYTES / 2);
                const uint64_t wd_hi = make_desc(w_hi), wd_lo = make_desc(w_hi + (uint32_t)op.n_total * 128u);
                const uint32_t first_acc = op.first ? 0u : 1u;
                if (elect_one()) {
#pragma unroll 1
                    for (int pc = 0; pc < op.n_piece; ++pc) {
                        const uint32_t dcol = tmem_base + (uint32_t)op.d_col[pc];
                        const uint32_t idesc = make_idesc(op.n_rows[pc]);
                        const uint64_t boff = (uint64_t)((op.b_row0[pc] >> 3) * (1024 >> 4));
                        const uint64_t bh = wd_hi + boff, bl = wd_lo + boff;
#pragma unroll
                        for (int ks = 0; ks < TC_KC / 8; ++ks) {
                            tc_mma_tf32(dcol, ad_hi + 2 * ks, bh + 2 * ks, idesc, ks == 0 ? first_acc : 1u);
#if !defined(TC_EXP_ONE_MMA)
                            tc_mma_tf32(dcol, ad_lo + 2 * ks, bh + 2 * ks, idesc, 1u);
                            tc_mma_tf32(dcol, ad_hi + 2 * ks, bl + 2 * ks, idesc, 1u);
#endif
                        }
                    }
                    tc_commit(sfree0 + 8 * s);
                    if (op.d_commit >= 0) tc_commit(dfull0 + 8 * op.d_commit);
                }
                __syncwarp();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_PROD_WARPS) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace sdfnmpc
