// Encoder convolutions on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Implicit GEMM of one Conv2d(+folded BatchNorm)+residual+activation layer of the depth encoder
// (sdf_nmpc/network/resnet.py:32-56, network/vae.py:19-38):   D[pixel][cout] = sum_k A[pixel][k] * W[cout][k],
// k = (ci, ky, kx), M = 128 output pixels per tile, N = NT output channels per tile (<= 256), K in chunks of 32.
// fp32-grade arithmetic from the tf32 hi / lo split (3 MMAs per product, fp32 accumulation in TMEM), as in the SDF kernel.
//
// One persistent CTA per SM, 17 warps:
//   * 16 producer warps in two ping-pong groups: thread <-> (pixel row, 16 of the chunk's 32 k values); they gather the
//     im2col patch values from the NCHW activation (lanes = consecutive pixels: coalesced for every k), split them into
//     tf32 hi | lo and store the A chunk in the 128-byte-swizzled K-major UMMA layout; after the K loop the same warps
//     drain the accumulator (tcgen05.ld) through bias + residual + activation to the NCHW output (again pixel-coalesced).
//   * 1 weight-streaming warp: the weights (pre-split hi | lo and pre-laid-out on the host) travel as sub-chunks of <= 128
//     output channels x 32 k (32 KB, one cp.async.bulk each) through a 4-stage ring.
//   * 1 MMA warp: waits for operands and issues the tcgen05.mma from one elected lane on uniform operands; it does
//     nothing else.  (With the bulk copies issued from the MMA warp - 500-800 cycles per copy in that thread, measured with
//     the clock64 trace - and a 2 x 64 KB ring that waited for "previous chunk retired" before every issue, the tensor pipe
//     idled as long as it worked: 4 k cycles per N = 256 chunk against 1.5 k of MMA.  tools/mma_rate.cu shows the pipe
//     itself sustains exactly 128 / 64 cycles per N = 256 / 128 MMA next to store and TMA traffic.)
// K is ordered (ci, ky, kx): the 9 taps of a channel sit in the same chunk and reuse the same L1 lines.  The (ky, kx, ci) order
// (one bounds test per chunk, no table) was measured slower, 3.25 vs 2.94 ms per 64 frames: no reuse inside a chunk.
// Measured alternatives (64 frames, tensor-core layers 2.1 ms of the 2.94 ms total): NHWC activations with a (ky, kx, ci)
// K order (4 x LDG.128 per thread and chunk instead of 16 table-driven loads) 3.05 ms - the gather is not the limiter;
// a 3-deep A ring 3.30 ms; timing-only ablations: no A-stage stores -29 % of the tensor-core time, 1 of 3 MMAs -22 %,
// no weight TMA 0 %, all three -57 %.  As in the SDF kernel the chunk period (~4 k cycles at N = 256 against a 1.5 k MMA
// floor) is set by shared-memory traffic that the operand fetch, the A stores and their proxy fences share.
// The accumulator is double-buffered in TMEM (2 x 256 columns), so the MMAs of the next work item start while the
// previous item's accumulator is still being drained by the slower producer group.
#pragma once
#include "tc_common.cuh"

namespace sdfnmpc {

constexpr int CV_PROD_WARPS = 16;
constexpr int CV_GROUP_THREADS = CV_PROD_WARPS * 32 / 2;
constexpr int CV_PROD_THREADS = CV_PROD_WARPS * 32;
constexpr int CV_THREADS = CV_PROD_THREADS + 64;         // + MMA issuer warp + weight-streaming warp
constexpr int CV_NSUB = 128;                             // output channels per weight sub-chunk
constexpr int CV_WSTAGES = 4;
constexpr int CV_WSTAGE_BYTES = CV_NSUB * TC_KC * 4 * 2; // 32 KB: hi | lo
constexpr int CV_SMEM_BYTES = CV_WSTAGES * CV_WSTAGE_BYTES + 2 * TC_ASTAGE_BYTES + 256 + 1024;

enum { CV_ACT_NONE = 0, CV_ACT_RELU = 1, CV_ACT_ELU = 2 };

struct ConvTcParams {
    int Cin, Hin, Win, Cout, Hout, Wout, stride, pad;
    long long npix;               // B * Hout * Wout
    int NT;                       // output channels per N tile: 32, 64, 128, 192 or 256 (Cout % NT == 0)
    int n_ntiles, n_kchunks;
    const float* in;              // [B][Cin][Hin][Win]
    const unsigned char* wblob;   // [n_ntiles][n_kchunks][NT / NS sub-chunks][hi: NS x 128 B swizzled | lo], NS = min(NT, 128)
    const int* ktab;              // [n_kchunks * 32] packed im2col entries: bits 0-23 element offset (ci*Hin + ky)*Win + kx, bits 24-28
                                  // the tap ky * ks + kx (second source: 0; k >= K: 31, never valid); four entries per 16-byte
                                  // broadcast load.  Whether a tap falls inside the image does not depend on the channel, so a
                                  // thread forms one validity bit per tap when it starts a tile and the gather tests that bit
                                  // (the four compares per element had been half of the kernel's instructions, r2n profile)
    int ks;                       // kernel size (ks * ks <= 30)
    // Second source (the ResBlock shortcut fused into the block's last convolution, resnet.py:40-54): k entries with source 2
    // read in2[b][ci][s2 * oy][s2 * ox] (offset ci * H2 * W2, always inside the image), i.e. the 1x1 stride-s2 convolution of
    // the block input accumulates into the same tile; its folded-BatchNorm shift is part of `bias`.
    const float* in2;             // [B][C2][H2][W2] or null
    int C2, H2, W2, s2;
    int kc2;                      // first chunk of the second source (= n_kchunks without one; Cin k^2 must be a multiple of 32)
    const float* bias;            // [Cout] (folded BatchNorm shift)
    const float* resid;           // [B][Cout][Hout][Wout] or null
    int act;
    float* out;                   // [B][Cout][Hout][Wout]
    // Split-K (the single-frame latency path): the K chunks of a tile are divided over `ksplit` work items, each of which
    // writes its raw accumulator to part[split][B][Cout][Hout][Wout]; conv_splitk_reduce_kernel adds the parts IN FIXED ORDER
    // and applies bias / residual / activation.  ksplit = 1, part = null: the epilogue below does it directly.
    int ksplit;
    float* part;
    // Sub-pixel output (conv_tc_kernel<*, true>, the decoder's stride-2 transposed convolutions, encoder_api.cu "decoder"): the
    // Cout channels are 4 parity classes x Cout / 4 real channels, class (py, px) of conv pixel (oy, ox) is output pixel
    // (2 oy + py, 2 ox + px) of a [B][Cout / 4][2 Hout][2 Wout] tensor (bias has Cout entries, resid / out / part that layout).
    long long* trace;             // development: clock64 samples of CTA 0's issuer, [sub-chunk][4] (null = off)
};

// SPLIT = false is the batch kernel (ksplit == 1 at compile time: the split bookkeeping costs it 6 % at 64 frames otherwise)
template <bool SPLIT, bool SHUF = false>
__global__ void __launch_bounds__(CV_THREADS, 1) conv_tc_kernel(const ConvTcParams P) {
    using namespace tc;
    extern __shared__ __align__(1024) unsigned char cv_smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)cv_smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* wst0 = smem;                                                               // CV_WSTAGES x 32 KB
    unsigned char* ast0 = smem + CV_WSTAGES * CV_WSTAGE_BYTES;                                // 2 x 32 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CV_WSTAGES * CV_WSTAGE_BYTES + 2 * TC_ASTAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    const uint32_t wfull0 = smem_u32(bars + 0), wfree0 = smem_u32(bars + 4);                  // + 8 * W stage
    // (stage addresses are computed, not looked up: a dynamically indexed local array costs a local-memory load per use)
    const uint32_t afull0 = smem_u32(bars + 8), sfree0 = smem_u32(bars + 10);                 // + 8 * A stage; sfree = consumed
    const uint32_t dfull = smem_u32(bars + 12);

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (tid == 0) {
        for (int q = 0; q < CV_WSTAGES; ++q) { mbar_init(wfull0 + 8 * q, 1); mbar_init(wfree0 + 8 * q, 1); }
        mbar_init(afull0, CV_GROUP_THREADS); mbar_init(afull0 + 8, CV_GROUP_THREADS);
        mbar_init(sfree0, 1); mbar_init(sfree0 + 8, 1);
        mbar_init(dfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == CV_PROD_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    const int n_ptiles = (int)((P.npix + TC_TM - 1) / TC_TM);
    const int KSP = SPLIT ? P.ksplit : 1;
    const int n_items = n_ptiles * P.n_ntiles * KSP;                 // item = (pixel tile * n_ntiles + n tile) * ksplit + k split
    const int nk = P.n_kchunks;
    const size_t out_total = (size_t)P.npix * P.Cout;
    const int hw = P.Hout * P.Wout;
    uint32_t chunk = 0;                                              // stage = chunk & 1, phase = (chunk >> 1) & 1

    if (warp < CV_PROD_WARPS) {
        const int g = warp >> 3, wq = warp & 7, lq = wq & 3, ch = wq >> 2;
        const int row = lq * 32 + (tid & 31);
        const uint32_t lane_addr = tmem_base + ((uint32_t)(lq * 32) << 16);
        uint32_t item_it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_it) {
            const int ksp = item % KSP, tile = item / KSP;
            const int pt = tile / P.n_ntiles, nt = tile - pt * P.n_ntiles;
            const int kc0 = (int)((long long)ksp * nk / KSP), kc1 = (int)((long long)(ksp + 1) * nk / KSP);
            const long long pix = (long long)pt * TC_TM + row;
            const bool pvalid = pix < P.npix;
            int pb = 0, oy = 0, ox = 0, r = 0;
            if (pvalid) {
                pb = (int)(pix / hw);
                r = (int)(pix - (long long)pb * hw);
                oy = r / P.Wout;
                ox = r - oy * P.Wout;
            }
            const int iy0 = oy * P.stride - P.pad, ix0 = ox * P.stride - P.pad;
            const float* inb = P.in + (size_t)pb * P.Cin * P.Hin * P.Win + (long long)iy0 * P.Win + ix0;
            uint32_t vmask = 0;                                          // bit tap: the tap's input pixel exists
            if (pvalid)
                for (int ky = 0, tp = 0; ky < P.ks; ++ky)
                    for (int kx = 0; kx < P.ks; ++kx, ++tp)
                        if (iy0 + ky >= 0 && iy0 + ky < P.Hin && ix0 + kx >= 0 && ix0 + kx < P.Win) vmask |= 1u << tp;
            for (int kc = kc0; kc < kc1; ++kc, ++chunk) {
                if ((int)(chunk & 1) != g) continue;
                const int4* kt = reinterpret_cast<const int4*>(P.ktab + kc * TC_KC + ch * 16);
                int tk[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {                            // same entries for the whole warp: broadcast loads
                    const int4 a = __ldg(kt + q);
                    tk[4 * q] = a.x; tk[4 * q + 1] = a.y; tk[4 * q + 2] = a.z; tk[4 * q + 3] = a.w;
                }
                float v[16];
                if (kc < P.kc2) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int t = tk[i];
                        v[i] = ((vmask >> ((t >> 24) & 31)) & 1u) ? __ldg(inb + (t & 0xFFFFFF)) : 0.f;
                    }
                } else {
                    // chunks of the second source (a whole chunk is one or the other: kc2 * 32 = Cin k^2): the loop above stays
                    // as it was - a per-element source select cost it the 16 loads in flight (18 % of the layer time, measured)
                    const float* inb2 = P.in2 + (size_t)pb * P.C2 * P.H2 * P.W2 + (long long)(oy * P.s2) * P.W2 + ox * P.s2;
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = (pvalid && ((tk[i] >> 24) & 31) == 0) ? __ldg(inb2 + (tk[i] & 0xFFFFFF)) : 0.f;
                }
                float vhi[16], vlo[16];
                split_a16(v, vhi, vlo);
                mbar_wait_backoff(sfree0 + 8 * g, ((chunk >> 1) & 1) ^ 1);
                store_a16(ast0 + g * TC_ASTAGE_BYTES, row, ch, vhi, vlo);
                fence_async_smem();
                mbar_arrive(afull0 + 8 * g);
            }
            // ---- epilogue: accumulator -> bias, residual, activation -> NCHW output.  The four (group, half) combinations
            // of a row take a quarter of the NT columns each. ----
            mbar_wait_backoff(dfull, item_it & 1);
            tc_fence_after();
            const uint32_t acc = lane_addr + (item_it & 1) * 256u;
            // (NT = 32, the decoder's last block: two of the four take 16 columns each, the others none)
            const int ncol = P.NT >= 64 ? P.NT >> 2 : 16, c_begin = (g * 2 + ch) * ncol;
            for (int c0 = c_begin; c0 < c_begin + ncol && c0 < P.NT; c0 += 16) {
                float v[16];
                tc_ld16(acc + c0, v);
                // output offset of channel co of this thread's pixel
                auto out_off = [&](int co) -> size_t {
                    if constexpr (SHUF) {
                        const int cr = P.Cout >> 2, cls = co / cr, c = co - cls * cr;
                        return (((size_t)pb * cr + c) * (2 * P.Hout) + 2 * oy + (cls >> 1)) * (size_t)(2 * P.Wout) + 2 * ox + (cls & 1);
                    } else {
                        return ((size_t)pb * P.Cout + co) * hw + r;
                    }
                };
                if (SPLIT && pvalid) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int co = nt * P.NT + c0 + e;
                        P.part[(size_t)ksp * out_total + out_off(co)] = v[e];
                    }
                } else if (!SPLIT && pvalid) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int co = nt * P.NT + c0 + e;
                        const size_t o = out_off(co);
                        float val = v[e] + __ldg(P.bias + co);
                        if (P.resid) val += __ldg(P.resid + o);
                        if (P.act == CV_ACT_RELU) val = fmaxf(val, 0.f);
                        else if (P.act == CV_ACT_ELU) val = val > 0.f ? val : expm1f(val);
                        P.out[o] = val;
                    }
                }
            }
            tc_fence_before();
        }
    } else {
        const uint32_t leader = elect_one();
        const int NS = P.NT < CV_NSUB ? P.NT : CV_NSUB, nsub = P.NT / NS;       // sub-chunks per chunk
        const uint32_t sub_bytes = (uint32_t)NS * TC_KC * 4 * 2;
        if (warp == CV_PROD_WARPS + 1) {
            // ============================== weight streamer ==============================
            if constexpr (!SPLIT) {
                // batch kernel: one flat loop over the sub-chunks of all of this CTA's items
                const int my_items = blockIdx.x < n_items ? (n_items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
                const uint64_t n_sub_total = (uint64_t)my_items * nk * nsub;
                int pf_item = blockIdx.x, pf_kc = 0, pf_h = 0;
                for (uint64_t pf = 0; pf < n_sub_total; ++pf) {
                    const uint32_t q = (uint32_t)(pf % CV_WSTAGES);
                    mbar_wait(wfree0 + 8 * q, (uint32_t)(((pf / CV_WSTAGES) & 1) ^ 1));      // previous user of the slot has retired
                    if (leader) {
                        const int pnt = pf_item % P.n_ntiles;
                        mbar_expect_tx(wfull0 + 8 * q, sub_bytes);
                        bulk_g2s(smem_u32(wst0 + q * CV_WSTAGE_BYTES), P.wblob + (((size_t)pnt * nk + pf_kc) * nsub + pf_h) * sub_bytes, sub_bytes,
                                 wfull0 + 8 * q);
                    }
                    __syncwarp();
                    if (++pf_h == nsub) { pf_h = 0; if (++pf_kc == nk) { pf_kc = 0; pf_item += gridDim.x; } }
                }
            } else {
                uint64_t pf = 0;
                for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                    const int ksp = item % KSP, pnt = (item / KSP) % P.n_ntiles;
                    const int kc0 = (int)((long long)ksp * nk / KSP), kc1 = (int)((long long)(ksp + 1) * nk / KSP);
                    for (int kc = kc0; kc < kc1; ++kc)
                        for (int h = 0; h < nsub; ++h, ++pf) {
                            const uint32_t q = (uint32_t)(pf % CV_WSTAGES);
                            mbar_wait(wfree0 + 8 * q, (uint32_t)(((pf / CV_WSTAGES) & 1) ^ 1));      // previous user of the slot has retired
                            if (leader) {
                                mbar_expect_tx(wfull0 + 8 * q, sub_bytes);
                                bulk_g2s(smem_u32(wst0 + q * CV_WSTAGE_BYTES), P.wblob + (((size_t)pnt * nk + kc) * nsub + h) * sub_bytes, sub_bytes,
                                         wfull0 + 8 * q);
                            }
                            __syncwarp();
                        }
                }
            }
        } else {
            // ============================== MMA issuer ==============================
            const uint32_t idesc = make_idesc(NS);
            uint32_t item_it = 0;
            uint64_t jj = 0;                                                         // sub-chunk counter (W stage = jj % 4)
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++item_it) {
                const uint32_t dcol = tmem_base + (item_it & 1) * 256u;
                const int ksp = item % KSP;
                const int kc0 = (int)((long long)ksp * nk / KSP), kc1 = (int)((long long)(ksp + 1) * nk / KSP);
                for (int kc = kc0; kc < kc1; ++kc, ++chunk) {
                    const int s = chunk & 1;
                    const uint32_t a_hi = smem_u32(ast0) + (uint32_t)s * TC_ASTAGE_BYTES;
                    const uint64_t ad_hi = make_desc(a_hi), ad_lo = make_desc(a_hi + TC_ASTAGE_BYTES / 2);
                    for (int h = 0; h < nsub; ++h, ++jj) {
                        const uint32_t q = (uint32_t)(jj % CV_WSTAGES);
#if defined(SDFNMPC_TRACE_BUILD)
                        const bool tr = P.trace && blockIdx.x == 0 && jj < 512 && leader;
#else
                        constexpr bool tr = false;
#endif
                        if (tr) P.trace[jj * 4 + 0] = clock64();
                        mbar_wait(wfull0 + 8 * q, (uint32_t)((jj / CV_WSTAGES) & 1));
                        if (tr) P.trace[jj * 4 + 1] = clock64();
                        if (h == 0) mbar_wait(afull0 + 8 * s, (chunk >> 1) & 1);
                        if (tr) P.trace[jj * 4 + 2] = clock64();
                        tc_fence_after();
                        const uint32_t w_hi = smem_u32(wst0 + q * CV_WSTAGE_BYTES);
                        const uint64_t wd_hi = make_desc(w_hi), wd_lo = make_desc(w_hi + (uint32_t)NS * 128u);
                        const uint32_t dc = dcol + (uint32_t)(h * NS);
                        if (elect_one()) {
#pragma unroll
                            for (int ks = 0; ks < TC_KC / 8; ++ks) {
                                tc_mma_tf32(dc, ad_hi + 2 * ks, wd_hi + 2 * ks, idesc, (kc == kc0 && ks == 0) ? 0u : 1u);
                                tc_mma_tf32(dc, ad_lo + 2 * ks, wd_hi + 2 * ks, idesc, 1u);
                                tc_mma_tf32(dc, ad_hi + 2 * ks, wd_lo + 2 * ks, idesc, 1u);
                            }
                            tc_commit(wfree0 + 8 * q);
                            if (h == nsub - 1) {
                                tc_commit(sfree0 + 8 * s);
                                if (kc == kc1 - 1) tc_commit(dfull);
                            }
                        }
                        __syncwarp();
                        if (tr) P.trace[jj * 4 + 3] = clock64();
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == CV_PROD_WARPS) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// Second half of a split-K convolution: out = act(sum_s part[s] + bias + resid), parts added in the fixed order s = 0, 1, ...
__global__ void conv_splitk_reduce_kernel(const float* __restrict__ part, int ksplit, long long total, int Cout, int hw,
                                          const float* __restrict__ bias, const float* __restrict__ resid, int act,
                                          float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float acc = part[i];
    for (int s = 1; s < ksplit; ++s) acc += part[(size_t)s * total + i];
    float val = acc + __ldg(bias + (int)((i / hw) % Cout));
    if (resid) val += __ldg(resid + i);
    if (act == CV_ACT_RELU) val = fmaxf(val, 0.f);
    else if (act == CV_ACT_ELU) val = val > 0.f ? val : expm1f(val);
    out[i] = val;
}

}  // namespace sdfnmpc
