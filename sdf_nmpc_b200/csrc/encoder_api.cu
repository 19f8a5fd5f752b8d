// Batched depth-image encoder (C ABI + kernels), fp32 implicit-GEMM convolutions on CUDA cores.
//
// Replaces VaeWrapper.set_img + encode (sdf_nmpc/vae.py:15-40): the preprocessing pipeline
// (utils/preprocessing.py: ToDevice :263-276, Reshape :83-96, ClipDistance :67-79, Depth2Range :5-30)
// is fused into the gather of the first convolution, BatchNorm (eval mode) is folded into the
// convolution weights, and the ResNet-style encoder of network/vae.py:19-38 + resnet.py:32-56 runs
// for a whole batch of frames: conv7x7/2+ELU, maxpool3/2, 4 ResBlocks, adaptive avg-pool (2,2), Linear.
//
// Convolution = implicit GEMM: M = Cout, N = B*OH*OW output pixels, K = Cin*KH*KW.  One CTA computes a
// 64 x 64 tile with a 16-deep shared-memory K pipeline; each of the 256 threads owns a 4x4 register tile.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "handles.h"
#include "conv_tc.cuh"

using namespace sdfnmpc;

namespace {

enum { ENC_ACT_NONE = 0, ENC_ACT_RELU = 1, ENC_ACT_ELU = 2 };

struct ConvShape {
    int Cin, Hin, Win, Cout, Hout, Wout, stride, pad;
    int shuffle = 0;      // sub-pixel output: Cout = 4 parity classes x Cout / 4 channels of a [2 Hout][2 Wout] map (the decoder)
};

constexpr int TCO = 64, TPX = 64, TK = 16;

// PRE: apply clip(clip(raw/dmax,0,1)*yz,0,1) to the (single-channel) input while gathering
template <int KS, bool PRE>
__global__ void __launch_bounds__(256)
conv_igemm_kernel(ConvShape s, int B, const float* __restrict__ in, const float* __restrict__ w /*[Cout][Cin*KS*KS]*/,
                  const float* __restrict__ bias, const float* __restrict__ resid, int act, float* __restrict__ out,
                  const float* __restrict__ yz, float dmax, int clip_dist, int depth2range) {
    __shared__ float As[TK][TCO + 4];
    __shared__ float Bs[TK][TPX + 4];
    const int K = s.Cin * KS * KS;
    const long long npix = (long long)B * s.Hout * s.Wout;
    const int co0 = blockIdx.y * TCO;
    const long long px0 = (long long)blockIdx.x * TPX;
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    // gather mapping: this thread always loads pixel (tid & 63), k rows (tid >> 6) + 4 j
    const int gp = tid & 63, gk = tid >> 6;
    const long long pix = px0 + gp;
    const bool pvalid = pix < npix;
    int pb = 0, oy = 0, ox = 0;
    if (pvalid) {
        pb = (int)(pix / (s.Hout * s.Wout));
        const int r = (int)(pix % (s.Hout * s.Wout));
        oy = r / s.Wout;
        ox = r % s.Wout;
    }
    const int iy0 = oy * s.stride - s.pad, ix0 = ox * s.stride - s.pad;
    const float* inb = in + (size_t)pb * s.Cin * s.Hin * s.Win;
    // weight mapping: k row = tid & 15 (contiguous in memory), co = (tid >> 4) + 16 j
    const int wk = tid & 15, wc = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = wc + 16 * j;
            float a = 0.f;
            if (k0 + wk < K && co0 + co < s.Cout) a = __ldg(w + (size_t)(co0 + co) * K + k0 + wk);
            As[wk][co] = a;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int kk = gk + 4 * j, k = k0 + kk;
            float b = 0.f;
            if (k < K && pvalid) {
                const int ci = k / (KS * KS), rem = k % (KS * KS), ky = rem / KS, kx = rem % KS;
                const int iy = iy0 + ky, ix = ix0 + kx;
                if (iy >= 0 && iy < s.Hin && ix >= 0 && ix < s.Win) {
                    b = __ldg(inb + ((size_t)ci * s.Hin + iy) * s.Win + ix);
                    if (PRE) {
                        if (clip_dist) b = fminf(fmaxf(b / dmax, 0.f), 1.f);
                        if (depth2range) b = fminf(fmaxf(b * __ldg(yz + iy * s.Win + ix), 0.f), 1.f);
                    }
                }
            }
            Bs[kk][gp] = b;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    const int hw = s.Hout * s.Wout;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int co = co0 + ty * 4 + i;
        if (co >= s.Cout) continue;
        const float bco = bias ? bias[co] : 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long p = px0 + tx * 4 + j;
            if (p >= npix) continue;
            const int b = (int)(p / hw), r = (int)(p % hw);
            size_t o = ((size_t)b * s.Cout + co) * hw + r;
            if (s.shuffle) {
                const int cr = s.Cout >> 2, cls = co / cr, c = co - cls * cr, oy = r / s.Wout, ox = r - oy * s.Wout;
                o = (((size_t)b * cr + c) * (2 * s.Hout) + 2 * oy + (cls >> 1)) * (size_t)(2 * s.Wout) + 2 * ox + (cls & 1);
            }
            float v = acc[i][j] + bco;
            if (resid) v += resid[o];
            if (act == ENC_ACT_RELU) v = fmaxf(v, 0.f);
            else if (act == ENC_ACT_ELU) v = v > 0.f ? v : expm1f(v);
            out[o] = v;
        }
    }
}

__global__ void maxpool3s2_kernel(int C, int Hin, int Win, int Hout, int Wout, long long total,
                                  const float* __restrict__ in, float* __restrict__ out) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int ox = (int)(gid % Wout), oy = (int)((gid / Wout) % Hout);
    const long long bc = gid / ((long long)Wout * Hout);
    const float* p = in + (size_t)bc * Hin * Win;
    float m = -INFINITY;
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * 2 - 1 + ky;
        if (iy < 0 || iy >= Hin) continue;
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = ox * 2 - 1 + kx;
            if (ix < 0 || ix >= Win) continue;
            m = fmaxf(m, p[iy * Win + ix]);
        }
    }
    out[gid] = m;
}


// ---------------------------------------------------------------------------------------------------------------------
// First layer fused: preprocessing -> Conv2d(1, 64, 7, stride 2, pad 3) -> +bias -> ELU -> MaxPool2d(3, 2, 1)
// (network/vae.py:20-22).  Direct convolution on the CUDA cores: with Cin = 1 the GEMM depth is only 49, too short for the
// tensor-core pipeline (measured), and the 530 MB conv0 activation never goes to HBM.  ELU is monotonic, so it (and the
// bias) is applied after the max.
// One CTA = 8 x 16 pooled pixels x 64 channels of one frame: conv region 17 x 36 (33 needed), input patch 39 x 77 in smem.
// A thread computes 4 adjacent conv pixels x 16 channels: per kernel row 13 patch values and 7 x 16 weights feed 448 FMAs.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int C0_PH = 8, C0_PW = 16, C0_CR = 2 * C0_PH + 1, C0_CC = 36, C0_IR = 2 * C0_CR + 5, C0_IC = 2 * C0_CC + 5;
constexpr int C0_THREADS = 320, C0_CH = 64, C0_CPR = 32;     // channels per round
// Patch layout in shared memory: column c of a row sits at (c % 8) * C0_S + c / 8 ("de-interleaved by 8"), rows C0_RS apart.
// A thread reads the 13 columns 8 cgp + i of its 4 conv pixels, so lanes with consecutive cgp read consecutive words; with the
// plain layout their stride was 8 words (3.2 wavefronts per load, simulated), with this one and C0_RS = 85 it is 1.9 - the
// rest comes from the ~4 patch rows a warp spans.
constexpr int C0_S = 10, C0_RS = 85;
static_assert(C0_S * 8 <= C0_RS && (C0_IC + 7) / 8 <= C0_S, "de-interleaved patch row");
constexpr int C0_PATCH_FLOATS = (C0_IR * C0_RS + 3) / 4 * 4;      // keeps the float4 regions behind it 16-byte aligned
constexpr int C0_SMEM_FLOATS = C0_PATCH_FLOATS + 49 * C0_CH + C0_CPR * C0_CR * C0_CC;

__global__ void __launch_bounds__(C0_THREADS, 2)
conv0_pool_kernel(int Hin, int Win, int Hc, int Wc, int Hp, int Wp, const float* __restrict__ img, const float* __restrict__ w /*[64][49]*/,
                  const float* __restrict__ bias, const float* __restrict__ yz, float dmax, int clip_dist, int depth2range,
                  float* __restrict__ out /*[B][64][Hp][Wp]*/) {
    extern __shared__ __align__(16) float c0sm[];
    float* patch = c0sm;                               // [C0_IR][C0_IC]
    float* wsm = patch + C0_PATCH_FLOATS;              // [49][64]
    float* csm = wsm + 49 * C0_CH;                     // [32][C0_CR][C0_CC]
    const int b = blockIdx.z, py0 = blockIdx.y * C0_PH, px0 = blockIdx.x * C0_PW;
    const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;    // first conv row / col of the region
    const int iy0 = 2 * cy0 - 3, ix0 = 2 * cx0 - 3;    // first input row / col of the patch
    const float* im = img + (size_t)b * Hin * Win;
    // patch staging in two passes - all of a thread's loads first, then the preprocessing and the stores: a rolled loop had one
    // load in flight per thread (long_scoreboard 1.1 and barrier 1.05 per issue in the r2g capture)
    {
        constexpr int PER = (C0_IR * C0_IC + C0_THREADS - 1) / C0_THREADS;
        float raw[PER], yzv[PER];
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const int i = threadIdx.x + q * C0_THREADS;
            const int r = i / C0_IC, c = i - r * C0_IC, iy = iy0 + r, ix = ix0 + c;
            const bool ok = i < C0_IR * C0_IC && iy >= 0 && iy < Hin && ix >= 0 && ix < Win;
            raw[q] = ok ? __ldg(im + (size_t)iy * Win + ix) : 0.f;
            yzv[q] = (ok && depth2range) ? __ldg(yz + iy * Win + ix) : 1.f;
        }
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const int i = threadIdx.x + q * C0_THREADS;
            const int r = i / C0_IC, c = i - r * C0_IC, iy = iy0 + r, ix = ix0 + c;
            float v = 0.f;
            if (iy >= 0 && iy < Hin && ix >= 0 && ix < Win) {
                v = raw[q];
                if (clip_dist) v = fminf(fmaxf(v / dmax, 0.f), 1.f);
                if (depth2range) v = fminf(fmaxf(v * yzv[q], 0.f), 1.f);
            }
            if (i < C0_IR * C0_IC) patch[r * C0_RS + (c & 7) * C0_S + (c >> 3)] = v;
        }
    }
    for (int i = threadIdx.x; i < 49 * C0_CH; i += C0_THREADS) {
        const int tap = i / C0_CH, c = i - tap * C0_CH;
        wsm[i] = __ldg(w + c * 49 + tap);
    }
    __syncthreads();
    constexpr int TASKS = C0_CR * (C0_CC / 4);          // (conv row, group of 4 conv columns) = 153
    const int t = threadIdx.x;
    const bool active = t < 2 * TASKS;
    const int cgl = active ? t / TASKS : 0, rem = active ? t - cgl * TASKS : 0, cr = rem / (C0_CC / 4), cgp = rem - cr * (C0_CC / 4);
    for (int round = 0; round < 2; ++round) {
        if (active) {
            const int cbase = round * C0_CPR + cgl * 16;
            float acc[4][16];
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int c = 0; c < 16; ++c) acc[p][c] = 0.f;
#pragma unroll 1
            for (int ky = 0; ky < 7; ++ky) {
                float seg[13];
                const float* pr = patch + (2 * cr + ky) * C0_RS + cgp;
#pragma unroll
                for (int i = 0; i < 13; ++i) seg[i] = pr[(i & 7) * C0_S + (i >> 3)];
#pragma unroll
                for (int kx = 0; kx < 7; ++kx) {
                    const float4* wp = reinterpret_cast<const float4*>(wsm + (ky * 7 + kx) * C0_CH + cbase);
                    const float4 w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
                    const float wv[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const float x = seg[2 * p + kx];
#pragma unroll
                        for (int c = 0; c < 16; ++c) acc[p][c] = fmaf(x, wv[c], acc[p][c]);
                    }
                }
            }
            // conv pixels outside the image are the max-pool's padding: stored as -inf, so that the pooling pass below reads its
            // nine values unconditionally (with the bounds tests per value that pass was 42 % of the kernel's instructions)
            const int cy = cy0 + cr, cxb = cx0 + 4 * cgp;
            const bool rowok = cy >= 0 && cy < Hc;
            bool ok[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) ok[p] = rowok && cxb + p >= 0 && cxb + p < Wc;
#pragma unroll
            for (int c = 0; c < 16; ++c)
                *reinterpret_cast<float4*>(csm + ((cgl * 16 + c) * C0_CR + cr) * C0_CC + 4 * cgp) =
                    make_float4(ok[0] ? acc[0][c] : -INFINITY, ok[1] ? acc[1][c] : -INFINITY, ok[2] ? acc[2][c] : -INFINITY, ok[3] ? acc[3][c] : -INFINITY);
        }
        __syncthreads();
        // (unrolling this loop, or batching the weight loads above like the patch loads, measured slower: 613 vs 589 us)
        for (int i = threadIdx.x; i < C0_CPR * C0_PH * C0_PW; i += C0_THREADS) {
            const int c = i / (C0_PH * C0_PW), q = i - c * (C0_PH * C0_PW), pyl = q / C0_PW, pxl = q - pyl * C0_PW;
            const int py = py0 + pyl, px = px0 + pxl;
            if (py >= Hp || px >= Wp) continue;
            const float* cw = csm + (c * C0_CR + 2 * pyl) * C0_CC + 2 * pxl;
            // (even offsets: the first two values of a row are one conflict-free 8-byte load)
            const float2 r0 = *reinterpret_cast<const float2*>(cw), r1 = *reinterpret_cast<const float2*>(cw + C0_CC),
                         r2 = *reinterpret_cast<const float2*>(cw + 2 * C0_CC);
            float m = fmaxf(fmaxf(r0.x, r0.y), cw[2]);
            m = fmaxf(m, fmaxf(fmaxf(r1.x, r1.y), cw[C0_CC + 2]));
            m = fmaxf(m, fmaxf(fmaxf(r2.x, r2.y), cw[2 * C0_CC + 2]));
            const int cglob = round * C0_CPR + c;
            const float v = m + __ldg(bias + cglob);
            out[(((size_t)b * C0_CH + cglob) * Hp + py) * Wp + px] = v > 0.f ? v : expm1f(v);
        }
        __syncthreads();
    }
}

// AdaptiveAvgPool2d((2,2)) + Flatten + Linear(4C -> L).  Grid (frames, PL_SPLIT): every CTA pools the frame's whole feature
// map (a warp per channel, lanes over the H*W contiguous pixels: coalesced) and computes L / PL_SPLIT of the outputs, so
// that a batch of 64 frames fills the GPU instead of 64 SMs.
constexpr int PL_SPLIT = 4;
__global__ void __launch_bounds__(256)
pool_linear_kernel(int C, int H, int W, int L, const float* __restrict__ feat, const float* __restrict__ mw,
                   const float* __restrict__ mb, float* __restrict__ latents) {
    extern __shared__ float f[];   // [4C], index c*4 + py*2 + px (the Flatten order of [C, 2, 2])
    const int b = blockIdx.x;
    const float* fb = feat + (size_t)b * C * H * W;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    // torch adaptive pooling bins: [floor(i*In/Out), ceil((i+1)*In/Out)) - with an odd size the middle row / column is in both
    const int ya1 = (H + 1) / 2, yb0 = H / 2, xa1 = (W + 1) / 2, xb0 = W / 2;
    const float inv[4] = {1.f / (float)(ya1 * xa1), 1.f / (float)(ya1 * (W - xb0)), 1.f / (float)((H - yb0) * xa1), 1.f / (float)((H - yb0) * (W - xb0))};
    for (int c = warp; c < C; c += nw) {
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        for (int p = lane; p < H * W; p += 32) {
            const int y = p / W, x = p - y * W;
            const float v = fb[(size_t)c * H * W + p];
            const bool ta = y < ya1, tb = y >= yb0, la = x < xa1, lb = x >= xb0;
            if (ta && la) s[0] += v;
            if (ta && lb) s[1] += v;
            if (tb && la) s[2] += v;
            if (tb && lb) s[3] += v;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
        }
        if (lane < 4) f[c * 4 + lane] = s[lane] * inv[lane];
    }
    __syncthreads();
    const int lper = (L + PL_SPLIT - 1) / PL_SPLIT, l0 = blockIdx.y * lper, l1 = min(L, l0 + lper);
    for (int l = l0 + warp; l < l1; l += nw) {
        const float* wr = mw + (size_t)l * 4 * C;
        float sum = 0.f;
        for (int j = lane; j < 4 * C; j += 32) sum = fmaf(wr[j], f[j], sum);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) latents[(size_t)b * L + l] = sum + mb[l];
    }
}

// yz_sqrt table of Depth2Range, computed with the reference's float32 operation order (preprocessing.py:20-27)
__global__ void yz_table_kernel(int H, int W, float tan_h, float tan_v, float* __restrict__ yz) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= H * W) return;
    const float u = (float)(idx % W), v = (float)(idx / W);
    const float a = tan_h * (1.f - 2.f * u / (float)W);
    const float b = tan_v * (1.f - 2.f * v / (float)H);
    yz[idx] = sqrtf(1.f + a * a + b * b);
}

struct Layer {
    ConvShape s;
    int ks;
    float *w = nullptr, *b = nullptr;
    // tensor-core form (conv_tc.cuh): weight chunks [n tile][k chunk][hi | lo] in the swizzled UMMA layout + im2col table
    unsigned char* wblob_tc = nullptr;
    int* ktab = nullptr;             // packed im2col entries (ConvTcParams::ktab)
    int NT = 0, n_ntiles = 0, n_kchunks = 0;
    // tensor-core form only: the block's 1x1 stride-s2 shortcut appended to K (ConvTcParams::in2), bias = both folded shifts
    float* b_fused = nullptr;
    int C2 = 0, H2 = 0, W2 = 0, s2 = 0;
};

inline float tf32_rna_host(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u + 0x1000u) & 0xFFFFE000u;       // round to nearest, ties away (cvt.rna.tf32.f32)
    float r;
    memcpy(&r, &u, 4);
    return r;
}

}  // namespace

struct sdfnmpc_encoder {
    sdfnmpc_encoder_desc desc;
    Layer conv0, c1[4], c2[4], sc[4];
    bool has_sc[4];
    float *mean_w = nullptr, *mean_b = nullptr, *yz = nullptr;
    float *buf[4] = {nullptr, nullptr, nullptr, nullptr};   // activation ping-pong
    size_t buf_floats = 0;
    int Hp, Wp;       // after max-pool
    int Hf, Wf;       // final feature map
    float *h_img = nullptr, *h_lat = nullptr;
    float* part = nullptr;      // split-K partial accumulators of the single-frame path
    size_t part_floats = 0;
};

static int out_dim(int in, int k, int s, int p) { return (in + 2 * p - k) / s + 1; }
constexpr int ENC_SPLIT_MAX_B = 4;      // batches of up to this many frames take the split-K latency path (conv_ksplit)
static int conv_ksplit(const Layer& L, int B, int sm_count);

static int upload(float** dst, const std::vector<float>& v) {
    SDF_CUDA_CHECK(cudaMalloc(dst, v.size() * sizeof(float)));
    SDF_CUDA_CHECK(cudaMemcpy(*dst, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}

// fold BN (weight, bias, mean, var) into conv weights [Cout][K] and a bias [Cout]
// (a sub-pixel layer has its BatchNorm / bias entries per REAL channel: index co % (Cout / 4))
// `sc` (optional): the shortcut layer of the same block, already made (folded weights sc->w [Cout][C2], shift sc->b); its K
// entries are appended to this layer's tensor-core operands.
static int make_layer(Layer& L, const float* w, const float* const bn[4], const float* conv_bias, float eps, const Layer* sc = nullptr,
                      const std::vector<float>* sc_w = nullptr, const std::vector<float>* sc_b = nullptr,
                      std::vector<float>* W_out = nullptr, std::vector<float>* B_out = nullptr) {
    const int K = L.s.Cin * L.ks * L.ks;
    std::vector<float> W((size_t)L.s.Cout * K), Bv(L.s.Cout, 0.f);
    const int cmod = L.s.shuffle ? L.s.Cout / 4 : L.s.Cout;
    for (int co = 0; co < L.s.Cout; ++co) {
        const int cb = co % cmod;
        float scale = 1.f, shift = conv_bias ? conv_bias[cb] : 0.f;
        if (bn && bn[0]) {
            const float inv = 1.0f / sqrtf(bn[3][cb] + eps);
            scale = bn[0][cb] * inv;
            shift = bn[1][cb] + (shift - bn[2][cb]) * scale;
        }
        for (int k = 0; k < K; ++k) W[(size_t)co * K + k] = w[(size_t)co * K + k] * scale;
        Bv[co] = shift;
    }
    if (int rc = upload(&L.w, W)) return rc;
    if (int rc = upload(&L.b, Bv)) return rc;
    if (W_out) *W_out = W;
    if (B_out) *B_out = Bv;
    // ---- tensor-core operands (layers with Cout a multiple of 64; conv0 with its fused preprocessing stays on the
    // CUDA-core kernel) ----
    const int NT = L.s.Cout <= 256 ? L.s.Cout : 256;
    // (conv0 - Cin = 1, K = 49: two chunks per tile - was measured on this kernel too: 2.0 ms vs 1.14 ms at B = 64; a
    // tile that short is all pipeline fill / epilogue, so it stays on the CUDA cores)
    if (L.s.Cin < 8 || (NT % 64 && NT != 32) || L.s.Cout % NT) return 0;
    if (sc && K % TC_KC) sc = nullptr;                           // a chunk must belong to one source
    const int K2 = sc ? sc->s.Cin : 0, KT = K + K2;              // second source: one k entry per shortcut input channel
    if (sc) {
        L.C2 = sc->s.Cin; L.H2 = sc->s.Hin; L.W2 = sc->s.Win; L.s2 = sc->s.stride;
        std::vector<float> bf(L.s.Cout);
        for (int co = 0; co < L.s.Cout; ++co) bf[co] = Bv[co] + (*sc_b)[co];
        if (int rc = upload(&L.b_fused, bf)) return rc;
    }
    auto wat = [&](int co, int kk) { return kk < K ? W[(size_t)co * K + kk] : (kk < KT ? (*sc_w)[(size_t)co * K2 + (kk - K)] : 0.f); };
    const int nk = (KT + TC_KC - 1) / TC_KC;
    L.NT = NT; L.n_ntiles = L.s.Cout / NT; L.n_kchunks = nk;
    // sub-chunks of NS = min(NT, 128) output channels: [n tile][k chunk][sub][hi | lo], each one bulk copy (conv_tc.cuh)
    const int NS = NT < CV_NSUB ? NT : CV_NSUB, nsub = NT / NS;
    std::vector<float> blob((size_t)L.n_ntiles * nk * NT * TC_KC * 2, 0.f);
    for (int nt = 0; nt < L.n_ntiles; ++nt)
        for (int kc = 0; kc < nk; ++kc)
            for (int hs = 0; hs < nsub; ++hs) {
                const size_t base = (((size_t)nt * nk + kc) * nsub + hs) * NS * TC_KC * 2;
                for (int n = 0; n < NS; ++n)
                    for (int k = 0; k < TC_KC; ++k) {
                        const int kk = kc * TC_KC + k;
                        const float v = wat(nt * NT + hs * NS + n, kk);
                        const float hi = tf32_rna_host(v), lo = tf32_rna_host(v - hi);
                        // 128-byte-swizzled K-major: 8-row atoms of 1024 B, row = 32 floats, 16-byte unit index XOR row-in-atom
                        const size_t idx = (size_t)(n >> 3) * 256 + (size_t)(n & 7) * 32 + (size_t)(((k >> 2) ^ (n & 7)) * 4) + (k & 3);
                        blob[base + idx] = hi;
                        blob[base + (size_t)NS * TC_KC + idx] = lo;
                    }
            }
    float* dblob = nullptr;
    if (int rc = upload(&dblob, blob)) return rc;
    L.wblob_tc = reinterpret_cast<unsigned char*>(dblob);
    // packed im2col table: 24 bits of element offset (a layer whose input does not fit stays on the CUDA-core kernel)
    if ((size_t)L.s.Cin * L.s.Hin * L.s.Win >= (1u << 24) || (size_t)K2 * L.H2 * L.W2 >= (1u << 24) || L.ks * L.ks > 30) {
        cudaFree(dblob); L.wblob_tc = nullptr; cudaFree(L.b_fused); L.b_fused = nullptr;
        return 0;
    }
    std::vector<int> kt((size_t)nk * TC_KC, 31 << 24);            // tap 31: never valid (k >= K)
    for (int k = 0; k < nk * TC_KC; ++k) {
        if (k < K) {
            const int ci = k / (L.ks * L.ks), rem = k % (L.ks * L.ks), ky = rem / L.ks, kx = rem % L.ks;
            kt[k] = ((ci * L.s.Hin + ky) * L.s.Win + kx) | (rem << 24);
        } else if (k < KT) {
            kt[k] = (k - K) * L.H2 * L.W2;                        // second source: tap 0
        }
    }
    SDF_CUDA_CHECK(cudaMalloc(&L.ktab, kt.size() * sizeof(int)));
    SDF_CUDA_CHECK(cudaMemcpy(L.ktab, kt.data(), kt.size() * sizeof(int), cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int sdfnmpc_encoder_create(sdfnmpc_encoder_t* out, const sdfnmpc_encoder_desc* d, const sdfnmpc_encoder_weights* w) {
    if (!out || !d || !w || d->height < 16 || d->width < 16 || d->max_batch < 1 || d->size_latent < 1) return SDFNMPC_E_INVALID;
    if (int rc = ensure_device()) return rc;
    auto* h = new sdfnmpc_encoder();
    h->desc = *d;
    const int H0 = d->height, W0 = d->width;
    h->conv0.s = {1, H0, W0, 64, out_dim(H0, 7, 2, 3), out_dim(W0, 7, 2, 3), 2, 3};
    h->conv0.ks = 7;
    int rc = make_layer(h->conv0, w->conv0_w, nullptr, w->conv0_b, w->bn_eps);
    h->Hp = out_dim(h->conv0.s.Hout, 3, 2, 1);
    h->Wp = out_dim(h->conv0.s.Wout, 3, 2, 1);
    int C = 64, Hc = h->Hp, Wc = h->Wp;
    size_t maxf = (size_t)64 * h->conv0.s.Hout * h->conv0.s.Wout;
    const int strides[4] = {2, 2, 2, 1};
    for (int i = 0; i < 4 && rc == 0; ++i) {
        const int st = strides[i], Co = C * st;
        const int Ho = out_dim(Hc, 3, st, 1), Wo = out_dim(Wc, 3, st, 1);
        h->c1[i].s = {C, Hc, Wc, Co, Ho, Wo, st, 1}; h->c1[i].ks = 3;
        h->c2[i].s = {Co, Ho, Wo, Co, Ho, Wo, 1, 1}; h->c2[i].ks = 3;
        rc = make_layer(h->c1[i], w->conv1_w[i], w->bn1[i], nullptr, w->bn_eps);
        h->has_sc[i] = st != 1;
        std::vector<float> scW, scB;
        if (h->has_sc[i]) {
            if (!w->sc_w[i]) { rc = SDFNMPC_E_INVALID; break; }
            h->sc[i].s = {C, Hc, Wc, Co, out_dim(Hc, 1, st, 0), out_dim(Wc, 1, st, 0), st, 0}; h->sc[i].ks = 1;
            if (!rc) rc = make_layer(h->sc[i], w->sc_w[i], w->sc_bn[i], nullptr, w->bn_eps, nullptr, nullptr, nullptr, &scW, &scB);
        }
        // the block's last convolution carries the shortcut as extra K entries on the tensor-core path (conv_tc.cuh: in2)
        if (!rc) rc = make_layer(h->c2[i], w->conv2_w[i], w->bn2[i], nullptr, w->bn_eps, h->has_sc[i] ? &h->sc[i] : nullptr, &scW, &scB);
        maxf = std::max(maxf, (size_t)Co * Ho * Wo);
        C = Co; Hc = Ho; Wc = Wo;
    }
    h->Hf = Hc; h->Wf = Wc;
    if (!rc && C * 4 != 2048) rc = SDFNMPC_E_INVALID;
    if (!rc) {
        std::vector<float> mw(w->mean_w, w->mean_w + (size_t)d->size_latent * 2048), mb(w->mean_b, w->mean_b + d->size_latent);
        rc = upload(&h->mean_w, mw);
        if (!rc) rc = upload(&h->mean_b, mb);
    }
    h->buf_floats = maxf * (size_t)d->max_batch;
    for (int i = 0; i < 3 && !rc; ++i) rc = (int)cudaMalloc(&h->buf[i], h->buf_floats * sizeof(float));
    if (!rc) rc = (int)cudaMalloc(&h->yz, (size_t)H0 * W0 * sizeof(float));
    if (!rc) {
        // split-K scratch of the small-batch path: the largest ksplit x output of any tensor-core layer at B = ENC_SPLIT_MAX_B
        int dev = 0, sms = 0;
        rc = (int)cudaGetDevice(&dev);
        if (!rc) rc = (int)cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const Layer* ls[] = {&h->c1[0], &h->c1[1], &h->c1[2], &h->c1[3], &h->c2[0], &h->c2[1], &h->c2[2], &h->c2[3], &h->sc[1], &h->sc[2], &h->sc[3]};
        for (const Layer* l : ls)
            if (l->wblob_tc && !rc) {
                const int ks = conv_ksplit(*l, 1, sms);
                const size_t nb = (size_t)std::min(d->max_batch, ENC_SPLIT_MAX_B);
                if (ks > 1) h->part_floats = std::max(h->part_floats, (size_t)ks * nb * l->s.Cout * l->s.Hout * l->s.Wout);
            }
        if (!rc && h->part_floats) rc = (int)cudaMalloc(&h->part, h->part_floats * sizeof(float));
    }
    if (!rc) {
        // torch.tan(torch.tensor(hfov, dtype=float32)): tangent evaluated on the float32-rounded angle
        yz_table_kernel<<<(H0 * W0 + 255) / 256, 256>>>(H0, W0, tanf((float)d->hfov), tanf((float)d->vfov), h->yz);
        count_launch();
        rc = (int)cudaDeviceSynchronize();
    }
    if (rc) { sdfnmpc_encoder_destroy(h); return rc; }
    *out = h;
    return 0;
}

extern "C" int sdfnmpc_encoder_destroy(sdfnmpc_encoder_t h) {
    if (!h) return 0;
    Layer* ls[] = {&h->conv0, &h->c1[0], &h->c1[1], &h->c1[2], &h->c1[3], &h->c2[0], &h->c2[1], &h->c2[2], &h->c2[3],
                   &h->sc[0], &h->sc[1], &h->sc[2], &h->sc[3]};
    for (Layer* l : ls) { cudaFree(l->w); cudaFree(l->b); cudaFree(l->wblob_tc); cudaFree(l->ktab); cudaFree(l->b_fused); }
    cudaFree(h->mean_w); cudaFree(h->mean_b); cudaFree(h->yz);
    for (float* b : h->buf) cudaFree(b);
    cudaFree(h->h_img); cudaFree(h->h_lat); cudaFree(h->part);
    delete h;
    return 0;
}

// SDFNMPC_ENC_IMPL=ffma forces the CUDA-core kernels (A/B runs and the cross-check in tests/test_gpu_encoder.py)
static bool enc_use_tc() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SDFNMPC_ENC_IMPL");
        v = (e && !strcmp(e, "ffma")) ? 0 : 1;
    }
    return v == 1;
}

// Split-K factor of a layer.  A single frame is the reference's operating point (vae.py:31-40): its deep layers have 2-16 tiles
// for 148 SMs and a K of up to 144 chunks, so the tile's K range is divided over the idle SMs (>= 2 chunks each).  The factor is
// a function of the LAYER only - the one that suits a single frame - and applies to batches of up to ENC_SPLIT_MAX_B frames, so
// that a row does not depend on which small batch it is encoded in.  Larger batches are never split (their rows must not depend
// on how a set of frames is divided into batches either: shard invariance of the rollouts); between the two regimes a row
// differs by the summation order only (measured 3e-5, the split being the one closer to the float64 oracle).
static int conv_ksplit(const Layer& L, int B, int sm_count) {
    if (B > ENC_SPLIT_MAX_B) return 1;
    const long long n_tiles = ((long long)L.s.Hout * L.s.Wout + TC_TM - 1) / TC_TM * L.n_ntiles;      // of ONE frame
    if (n_tiles * 2 > sm_count) return 1;
    const long long by_sm = sm_count / n_tiles, by_k = L.n_kchunks / 2;
    const long long k = by_sm < by_k ? by_sm : by_k;
    return k < 2 ? 1 : (int)k;
}

static int launch_conv_tc(const Layer& L, int B, const float* in, const float* resid, int act, float* out, cudaStream_t st,
                          float* part = nullptr, size_t part_floats = 0, const float* in2 = nullptr) {
    static int sm_count = 0;
    static bool attr_set = false;
    if (!sm_count) {
        int dev = 0;
        SDF_CUDA_CHECK(cudaGetDevice(&dev));
        SDF_CUDA_CHECK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    if (!attr_set) {
        SDF_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CV_SMEM_BYTES));
        SDF_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CV_SMEM_BYTES));
        SDF_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CV_SMEM_BYTES));
        SDF_CUDA_CHECK(cudaFuncSetAttribute(conv_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CV_SMEM_BYTES));
        attr_set = true;
    }
    ConvTcParams P;
    P.Cin = L.s.Cin; P.Hin = L.s.Hin; P.Win = L.s.Win; P.Cout = L.s.Cout; P.Hout = L.s.Hout; P.Wout = L.s.Wout;
    P.stride = L.s.stride; P.pad = L.s.pad; P.ks = L.ks;
    P.npix = (long long)B * L.s.Hout * L.s.Wout;
    P.NT = L.NT; P.n_ntiles = L.n_ntiles; P.n_kchunks = L.n_kchunks;
    P.in = in; P.wblob = L.wblob_tc; P.ktab = L.ktab; P.bias = L.b; P.resid = resid; P.act = act; P.out = out;
    P.in2 = nullptr; P.C2 = P.H2 = P.W2 = P.s2 = 0; P.kc2 = L.n_kchunks;
    if (L.b_fused) {                 // fused shortcut: the caller passes the block input instead of a residual tensor
        if (!in2 || resid) return SDFNMPC_E_INVALID;
        P.in2 = in2; P.C2 = L.C2; P.H2 = L.H2; P.W2 = L.W2; P.s2 = L.s2; P.bias = L.b_fused;
        P.kc2 = L.s.Cin * L.ks * L.ks / TC_KC;
    }
    P.ksplit = part ? conv_ksplit(L, B, sm_count) : 1;
    if ((size_t)P.ksplit * P.npix * P.Cout > part_floats) P.ksplit = 1;
    P.part = P.ksplit > 1 ? part : nullptr;
    const long long n_items = (P.npix + TC_TM - 1) / TC_TM * L.n_ntiles * P.ksplit;
    const int grid = (int)(n_items < sm_count ? n_items : sm_count);
    P.trace = nullptr;
    static long long* tbuf = nullptr;
    static const char* const tp = getenv("SDFNMPC_CV_TRACE");   // development: "<Cin>x<Cout>:<file>" traces that layer's issuer (read once)
    int tci = 0, tco = 0; char tfile[256] = {0};
    const bool want = tp && sscanf(tp, "%dx%d:%255s", &tci, &tco, tfile) == 3 && tci == L.s.Cin && tco == L.s.Cout && L.ks == 3;
    if (want) {
        if (!tbuf) cudaMalloc(&tbuf, 512 * 5 * sizeof(long long));
        cudaMemset(tbuf, 0, 512 * 5 * sizeof(long long));
        P.trace = tbuf;
    }
    if (L.s.shuffle) {
        if (P.ksplit > 1) conv_tc_kernel<true, true><<<grid, CV_THREADS, CV_SMEM_BYTES, st>>>(P);
        else conv_tc_kernel<false, true><<<grid, CV_THREADS, CV_SMEM_BYTES, st>>>(P);
    } else if (P.ksplit > 1) conv_tc_kernel<true><<<grid, CV_THREADS, CV_SMEM_BYTES, st>>>(P);
    else conv_tc_kernel<false><<<grid, CV_THREADS, CV_SMEM_BYTES, st>>>(P);
    if (want) {
        std::vector<long long> tr(512 * 5);
        cudaMemcpy(tr.data(), tbuf, tr.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        if (FILE* f = fopen(tfile, "w")) {
            for (int i = 0; i < 512; ++i) fprintf(f, "%lld %lld %lld %lld %lld\n", tr[i * 4] - tr[0], tr[i * 4 + 1] - tr[0], tr[i * 4 + 2] - tr[0], tr[i * 4 + 3] - tr[0], tr[2048 + i] - tr[0]);
            fclose(f);
        }
    }
    count_launch();
    SDF_CUDA_CHECK(cudaGetLastError());
    if (P.ksplit > 1) {
        const long long total = P.npix * P.Cout;
        const int cr = L.s.shuffle ? P.Cout / 4 : P.Cout, ohw = L.s.shuffle ? 4 * P.Hout * P.Wout : P.Hout * P.Wout;
        conv_splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(part, P.ksplit, total, cr, ohw, P.bias, resid, act, out);
        count_launch();
        SDF_CUDA_CHECK(cudaGetLastError());
    }
    return 0;
}

template <int KS, bool PRE>
static int launch_conv(const Layer& L, int B, const float* in, const float* resid, int act, float* out, cudaStream_t st,
                       const float* yz = nullptr, float dmax = 1.f, int clip = 0, int d2r = 0, float* part = nullptr, size_t part_floats = 0) {
    if (!PRE && L.wblob_tc && enc_use_tc()) return launch_conv_tc(L, B, in, resid, act, out, st, part, part_floats);
    const long long npix = (long long)B * L.s.Hout * L.s.Wout;
    dim3 grid((unsigned)((npix + TPX - 1) / TPX), (L.s.Cout + TCO - 1) / TCO);
    conv_igemm_kernel<KS, PRE><<<grid, 256, 0, st>>>(L.s, B, in, L.w, L.b, resid, act, out, yz, dmax, clip, d2r);
    count_launch();
    SDF_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int sdfnmpc_encoder_encode(sdfnmpc_encoder_t h, const float* imgs, int32_t B, float* latents, void* stream) {
    if (!h || !imgs || !latents || B < 1 || B > h->desc.max_batch) return SDFNMPC_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    float *a = h->buf[0], *b = h->buf[1], *c = h->buf[2];
    int rc = 0;
    if (enc_use_tc()) {
        // fused preprocessing + conv0 + ELU + max-pool (the CUDA-core reference path below keeps the separate kernels)
        static bool attr_set = false;
        if (!attr_set) {
            SDF_CUDA_CHECK(cudaFuncSetAttribute(conv0_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C0_SMEM_FLOATS * (int)sizeof(float)));
            attr_set = true;
        }
        dim3 grid((h->Wp + C0_PW - 1) / C0_PW, (h->Hp + C0_PH - 1) / C0_PH, B);
        conv0_pool_kernel<<<grid, C0_THREADS, C0_SMEM_FLOATS * sizeof(float), st>>>(
            h->desc.height, h->desc.width, h->conv0.s.Hout, h->conv0.s.Wout, h->Hp, h->Wp, imgs, h->conv0.w, h->conv0.b, h->yz,
            h->desc.dmax_scale, h->desc.is_normalized ? 0 : 1, h->desc.is_depth ? 1 : 0, b);
        count_launch();
        SDF_CUDA_CHECK(cudaGetLastError());
    } else {
        rc = launch_conv<7, true>(h->conv0, B, imgs, nullptr, ENC_ACT_ELU, a, st, h->yz, h->desc.dmax_scale,
                                  h->desc.is_normalized ? 0 : 1, h->desc.is_depth ? 1 : 0);
        if (rc) return rc;
        const long long total = (long long)B * 64 * h->Hp * h->Wp;
        maxpool3s2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(64, h->conv0.s.Hout, h->conv0.s.Wout, h->Hp, h->Wp, total, a, b);
        count_launch();
        SDF_CUDA_CHECK(cudaGetLastError());
    }
    float* x = b;          // block input
    float* t1 = a;         // conv1 output
    float* t2 = c;         // shortcut output / block output
    for (int i = 0; i < 4; ++i) {
        if ((rc = launch_conv<3, false>(h->c1[i], B, x, nullptr, ENC_ACT_RELU, t1, st, nullptr, 1.f, 0, 0, h->part, h->part_floats))) return rc;
        const float* resid = x;
        if (h->has_sc[i] && h->c2[i].b_fused && enc_use_tc()) {
            // tensor-core path: the 1x1 stride-2 shortcut of the block input rides in conv2's K range (no shortcut launch, no
            // residual tensor); the block input is an operand, so the result goes to the third buffer
            if ((rc = launch_conv_tc(h->c2[i], B, t1, nullptr, ENC_ACT_RELU, t2, st, h->part, h->part_floats, x))) return rc;
            float* tmp = x; x = t2; t2 = tmp;
        } else if (h->has_sc[i]) {
            if ((rc = launch_conv<1, false>(h->sc[i], B, x, nullptr, ENC_ACT_NONE, t2, st, nullptr, 1.f, 0, 0, h->part, h->part_floats))) return rc;
            resid = t2;
            // conv2 writes over the (now dead) block input
            if ((rc = launch_conv<3, false>(h->c2[i], B, t1, resid, ENC_ACT_RELU, x, st, nullptr, 1.f, 0, 0, h->part, h->part_floats))) return rc;
        } else {
            if ((rc = launch_conv<3, false>(h->c2[i], B, t1, resid, ENC_ACT_RELU, t2, st, nullptr, 1.f, 0, 0, h->part, h->part_floats))) return rc;
            float* tmp = x; x = t2; t2 = tmp;
        }
    }
    pool_linear_kernel<<<dim3(B, PL_SPLIT), 256, 2048 * sizeof(float), st>>>(512, h->Hf, h->Wf, h->desc.size_latent, x, h->mean_w, h->mean_b, latents);
    count_launch();
    SDF_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int sdfnmpc_encoder_encode_host(sdfnmpc_encoder_t h, const float* imgs, int32_t B, float* latents) {
    if (!h || !imgs || !latents || B < 1 || B > h->desc.max_batch) return SDFNMPC_E_INVALID;
    const size_t img_floats = (size_t)h->desc.height * h->desc.width;
    if (!h->h_img) SDF_CUDA_CHECK(cudaMalloc(&h->h_img, img_floats * h->desc.max_batch * sizeof(float)));
    if (!h->h_lat) SDF_CUDA_CHECK(cudaMalloc(&h->h_lat, (size_t)h->desc.size_latent * h->desc.max_batch * sizeof(float)));
    SDF_CUDA_CHECK(cudaMemcpyAsync(h->h_img, imgs, img_floats * B * sizeof(float), cudaMemcpyHostToDevice, 0));
    if (int rc = sdfnmpc_encoder_encode(h, h->h_img, B, h->h_lat, nullptr)) return rc;
    SDF_CUDA_CHECK(cudaMemcpyAsync(latents, h->h_lat, (size_t)h->desc.size_latent * B * sizeof(float), cudaMemcpyDeviceToHost, 0));
    SDF_CUDA_CHECK(cudaStreamSynchronize(0));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// VAE decoder (VaeWrapper.decode, sdf_nmpc/vae.py:42-45; network/vae.py:63-90, network/resnet.py:59-111): latent ->
// Linear + ELU -> [512, 8, 15] -> 4 x ResBlockDeconv (stride 2, output_padding 1) -> ConvTranspose2d(32, 1, 5, padding 2)
// -> bilinear Upsample(shape_imgs) -> Sigmoid.  BatchNorm in eval mode is folded, dropout is the identity.
//
// Every transposed convolution runs as an ordinary convolution on the tensor-core kernel of the encoder (conv_tc.cuh):
//   * stride 1, 3x3, padding 1: a 3x3 convolution with the kernel flipped and the channel axes exchanged;
//   * stride 2, 3x3, padding 1, output_padding 1 (oy = 2 iy - 1 + ky): output rows of parity 0 see (ky 1, iy = a), rows of
//     parity 1 see (ky 2, iy = a) and (ky 0, iy = a + 1) - a 2x2 convolution over the INPUT grid with 4 x Cout channels, one
//     group per output parity class, written through the kernel's sub-pixel epilogue (ConvTcParams: "Sub-pixel output").
//     K = 4 Cin taps for 9 useful ones per 4 outputs (1.8 x the useful products) instead of the 4 x of a convolution over the
//     zero-stuffed input;
//   * the 1x1 stride-2 shortcut (oy = 2 iy): the same form with a 1x1 window, class (0, 0) carries the weights, the other
//     classes only the folded BatchNorm shift.
// ---------------------------------------------------------------------------------------------------------------------
namespace {

// out[b][j] = ELU(w[j] . z[b] + bias[j]): a warp per output row, the batch looped inside (weights read once)
__global__ void __launch_bounds__(256)
dec_linear_elu_kernel(int L, int J, int B, const float* __restrict__ z, const float* __restrict__ w, const float* __restrict__ bias,
                      float* __restrict__ out) {
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j >= J) return;
    const float* wr = w + (size_t)j * L;
    for (int b = 0; b < B; ++b) {
        float s = 0.f;
        for (int k = lane; k < L; k += 32) s = fmaf(__ldg(wr + k), __ldg(z + (size_t)b * L + k), s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            const float v = s + bias[j];
            out[(size_t)b * J + j] = v > 0.f ? v : expm1f(v);
        }
    }
}

// ConvTranspose2d(Cin, 1, 5, stride 1, padding 2) + bias as a 5x5 convolution with the flipped kernel: 16 x 16 output pixels per
// CTA, 8 input channels of the 20 x 20 patch in shared memory at a time.
constexpr int DO_T = 16, DO_C = 8;
__global__ void __launch_bounds__(DO_T * DO_T)
dec_out_conv_kernel(int Cin, int H, int W, const float* __restrict__ in, const float* __restrict__ w /*[Cin][25], flipped*/, float bias,
                    float* __restrict__ out /*[B][H][W]*/) {
    __shared__ float patch[DO_C][DO_T + 4][DO_T + 4 + 1];
    __shared__ float wsm[DO_C][25];
    const int b = blockIdx.z, y0 = blockIdx.y * DO_T, x0 = blockIdx.x * DO_T, tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const float* inb = in + (size_t)b * Cin * H * W;
    float acc = 0.f;
    for (int c0 = 0; c0 < Cin; c0 += DO_C) {
        for (int i = threadIdx.x; i < DO_C * (DO_T + 4) * (DO_T + 4); i += DO_T * DO_T) {
            const int c = i / ((DO_T + 4) * (DO_T + 4)), q = i - c * (DO_T + 4) * (DO_T + 4), r = q / (DO_T + 4), cc = q - r * (DO_T + 4);
            const int iy = y0 + r - 2, ix = x0 + cc - 2;
            float v = 0.f;
            if (c0 + c < Cin && iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(inb + ((size_t)(c0 + c) * H + iy) * W + ix);
            patch[c][r][cc] = v;
        }
        if (threadIdx.x < DO_C * 25) {
            const int c = threadIdx.x / 25, t = threadIdx.x - c * 25;
            wsm[c][t] = c0 + c < Cin ? __ldg(w + (size_t)(c0 + c) * 25 + t) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < DO_C; ++c)
#pragma unroll
            for (int ky = 0; ky < 5; ++ky)
#pragma unroll
                for (int kx = 0; kx < 5; ++kx) acc = fmaf(patch[c][ty + ky][tx + kx], wsm[c][ky * 5 + kx], acc);
        __syncthreads();
    }
    const int y = y0 + ty, x = x0 + tx;
    if (y < H && x < W) out[((size_t)b * H + y) * W + x] = acc + bias;
}

// torch.nn.Upsample(size, mode='bilinear') (align_corners False: source = max(0, scale (dst + 0.5) - 0.5), scale = in / out in
// float32) + Sigmoid
__global__ void dec_upsample_sigmoid_kernel(int Hi, int Wi, int Ho, int Wo, long long total, const float* __restrict__ in,
                                            float* __restrict__ out) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= total) return;
    const int x = (int)(gid % Wo), y = (int)((gid / Wo) % Ho);
    const long long b = gid / ((long long)Wo * Ho);
    const float sh = (float)Hi / (float)Ho, sw = (float)Wi / (float)Wo;
    const float fy = fmaxf(sh * ((float)y + 0.5f) - 0.5f, 0.f), fx = fmaxf(sw * ((float)x + 0.5f) - 0.5f, 0.f);
    const int y0 = (int)fy, x0 = (int)fx, y1 = y0 + (y0 < Hi - 1 ? 1 : 0), x1 = x0 + (x0 < Wi - 1 ? 1 : 0);
    const float ly = fy - (float)y0, lx = fx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
    const float* p = in + (size_t)b * Hi * Wi;
    const float v = hy * (hx * p[y0 * Wi + x0] + lx * p[y0 * Wi + x1]) + ly * (hx * p[y1 * Wi + x0] + lx * p[y1 * Wi + x1]);
    out[gid] = 1.f / (1.f + expf(-v));
}

constexpr int DEC_C0 = 512, DEC_H0 = 8, DEC_W0 = 15;     // network/vae.py:73-75 (hard-coded in the reference)

}  // namespace

struct sdfnmpc_decoder {
    sdfnmpc_decoder_desc desc;
    Layer c1[4], c2[4], sc[4];
    float *lin_w = nullptr, *lin_b = nullptr, *out_w = nullptr;
    float out_b = 0.f;
    float* buf[3] = {nullptr, nullptr, nullptr};
    float *part = nullptr, *h_lat = nullptr, *h_img = nullptr;
    size_t part_floats = 0;
    int Hf = 0, Wf = 0, Cf = 0;
};

// ConvTranspose2d weight [Cin][Cout][3][3], stride 2, padding 1, output_padding 1 -> 2x2 convolution [4 Cout][Cin][2][2]
static std::vector<float> convt_s2_to_subpixel(const float* wt, int Cin, int Cout) {
    std::vector<float> w((size_t)4 * Cout * Cin * 4, 0.f);
    const int kmap[2][2] = {{1, -1}, {2, 0}};            // [output parity][window offset] -> kernel index (-1: no tap)
    for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px)
            for (int co = 0; co < Cout; ++co)
                for (int ci = 0; ci < Cin; ++ci)
                    for (int dy = 0; dy < 2; ++dy)
                        for (int dx = 0; dx < 2; ++dx) {
                            const int ky = kmap[py][dy], kx = kmap[px][dx];
                            if (ky < 0 || kx < 0) continue;
                            w[((((size_t)(py * 2 + px) * Cout + co) * Cin + ci) * 2 + dy) * 2 + dx] = wt[(((size_t)ci * Cout + co) * 3 + ky) * 3 + kx];
                        }
    return w;
}

// ConvTranspose2d weight [Cin][Cout][k][k], stride 1, padding (k - 1) / 2 -> convolution weight [Cout][Cin][k][k] (flipped)
static std::vector<float> convt_s1_to_conv(const float* wt, int Cin, int Cout, int k) {
    std::vector<float> w((size_t)Cout * Cin * k * k);
    for (int co = 0; co < Cout; ++co)
        for (int ci = 0; ci < Cin; ++ci)
            for (int ky = 0; ky < k; ++ky)
                for (int kx = 0; kx < k; ++kx)
                    w[(((size_t)co * Cin + ci) * k + ky) * k + kx] = wt[(((size_t)ci * Cout + co) * k + (k - 1 - ky)) * k + (k - 1 - kx)];
    return w;
}

extern "C" int sdfnmpc_decoder_destroy(sdfnmpc_decoder_t h) {
    if (!h) return 0;
    for (int i = 0; i < 4; ++i)
        for (Layer* l : {&h->c1[i], &h->c2[i], &h->sc[i]}) { cudaFree(l->w); cudaFree(l->b); cudaFree(l->wblob_tc); cudaFree(l->ktab); }
    cudaFree(h->lin_w); cudaFree(h->lin_b); cudaFree(h->out_w);
    for (float* b : h->buf) cudaFree(b);
    cudaFree(h->part); cudaFree(h->h_lat); cudaFree(h->h_img);
    delete h;
    return 0;
}

extern "C" int sdfnmpc_decoder_create(sdfnmpc_decoder_t* out, const sdfnmpc_decoder_desc* d, const sdfnmpc_decoder_weights* w) {
    if (!out || !d || !w || d->height < 1 || d->width < 1 || d->max_batch < 1 || d->size_latent < 1 || !w->lin_w || !w->lin_b || !w->out_w ||
        !w->out_b)
        return SDFNMPC_E_INVALID;
    if (int rc = ensure_device()) return rc;
    auto* h = new sdfnmpc_decoder();
    h->desc = *d;
    int rc = 0, C = DEC_C0, Hc = DEC_H0, Wc = DEC_W0;
    const size_t J = (size_t)DEC_C0 * DEC_H0 * DEC_W0;
    size_t maxf = J;
    rc = upload(&h->lin_w, std::vector<float>(w->lin_w, w->lin_w + J * d->size_latent));
    if (!rc) rc = upload(&h->lin_b, std::vector<float>(w->lin_b, w->lin_b + J));
    for (int i = 0; i < 4 && !rc; ++i) {
        if (!w->ct1_w[i] || !w->ct2_w[i] || !w->sc_w[i]) { rc = SDFNMPC_E_INVALID; break; }
        const int Co = C / 2;
        // conv form: 2x2 window over the input grid, "output" grid = input grid, 4 Co channels, sub-pixel epilogue
        h->c1[i].s = {C, Hc, Wc, 4 * Co, Hc, Wc, 1, 0, 1}; h->c1[i].ks = 2;
        h->sc[i].s = {C, Hc, Wc, 4 * Co, Hc, Wc, 1, 0, 1}; h->sc[i].ks = 1;
        h->c2[i].s = {Co, 2 * Hc, 2 * Wc, Co, 2 * Hc, 2 * Wc, 1, 1, 0}; h->c2[i].ks = 3;
        const std::vector<float> w1 = convt_s2_to_subpixel(w->ct1_w[i], C, Co);
        std::vector<float> wsc((size_t)4 * Co * C, 0.f);                      // class (0, 0) = rows [0, Co)
        for (int co = 0; co < Co; ++co)
            for (int ci = 0; ci < C; ++ci) wsc[(size_t)co * C + ci] = w->sc_w[i][(size_t)ci * Co + co];
        const std::vector<float> w2 = convt_s1_to_conv(w->ct2_w[i], Co, Co, 3);
        rc = make_layer(h->c1[i], w1.data(), w->bn1[i], nullptr, w->bn_eps);
        if (!rc) rc = make_layer(h->sc[i], wsc.data(), w->sc_bn[i], nullptr, w->bn_eps);
        if (!rc) rc = make_layer(h->c2[i], w2.data(), w->bn2[i], nullptr, w->bn_eps);
        C = Co; Hc *= 2; Wc *= 2;
        maxf = std::max(maxf, (size_t)C * Hc * Wc);
    }
    h->Cf = C; h->Hf = Hc; h->Wf = Wc;
    if (!rc) rc = upload(&h->out_w, convt_s1_to_conv(w->out_w, C, 1, 5));
    h->out_b = w->out_b[0];
    maxf = std::max(maxf, (size_t)d->height * d->width);
    for (int i = 0; i < 3 && !rc; ++i) rc = (int)cudaMalloc(&h->buf[i], maxf * (size_t)d->max_batch * sizeof(float));
    if (!rc) {
        int dev = 0, sms = 0;
        rc = (int)cudaGetDevice(&dev);
        if (!rc) rc = (int)cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const size_t nb = (size_t)std::min(d->max_batch, ENC_SPLIT_MAX_B);
        for (int i = 0; i < 4 && !rc; ++i)
            for (const Layer* l : {&h->c1[i], &h->c2[i], &h->sc[i]})
                if (l->wblob_tc) {
                    const int ks = conv_ksplit(*l, 1, sms);
                    if (ks > 1) h->part_floats = std::max(h->part_floats, (size_t)ks * nb * l->s.Cout * l->s.Hout * l->s.Wout);
                }
        if (!rc && h->part_floats) rc = (int)cudaMalloc(&h->part, h->part_floats * sizeof(float));
    }
    if (rc) { sdfnmpc_decoder_destroy(h); return rc; }
    *out = h;
    return 0;
}

extern "C" int sdfnmpc_decoder_decode(sdfnmpc_decoder_t h, const float* latents, int32_t B, float* imgs, void* stream) {
    if (!h || !latents || !imgs || B < 1 || B > h->desc.max_batch) return SDFNMPC_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    const int J = DEC_C0 * DEC_H0 * DEC_W0;
    float *x = h->buf[0], *t1 = h->buf[1], *t2 = h->buf[2];
    dec_linear_elu_kernel<<<(J + 7) / 8, 256, 0, st>>>(h->desc.size_latent, J, B, latents, h->lin_w, h->lin_b, x);
    count_launch();
    SDF_CUDA_CHECK(cudaGetLastError());
    int rc = 0;
    for (int i = 0; i < 4; ++i) {
        if ((rc = launch_conv<2, false>(h->c1[i], B, x, nullptr, ENC_ACT_RELU, t1, st, nullptr, 1.f, 0, 0, h->part, h->part_floats))) return rc;
        if ((rc = launch_conv<1, false>(h->sc[i], B, x, nullptr, ENC_ACT_NONE, t2, st, nullptr, 1.f, 0, 0, h->part, h->part_floats))) return rc;
        if ((rc = launch_conv<3, false>(h->c2[i], B, t1, t2, ENC_ACT_RELU, x, st, nullptr, 1.f, 0, 0, h->part, h->part_floats))) return rc;
    }
    dim3 grid((h->Wf + DO_T - 1) / DO_T, (h->Hf + DO_T - 1) / DO_T, B);
    dec_out_conv_kernel<<<grid, DO_T * DO_T, 0, st>>>(h->Cf, h->Hf, h->Wf, x, h->out_w, h->out_b, t1);
    count_launch();
    SDF_CUDA_CHECK(cudaGetLastError());
    const long long total = (long long)B * h->desc.height * h->desc.width;
    dec_upsample_sigmoid_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(h->Hf, h->Wf, h->desc.height, h->desc.width, total, t1, imgs);
    count_launch();
    SDF_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int sdfnmpc_decoder_decode_host(sdfnmpc_decoder_t h, const float* latents, int32_t B, float* imgs) {
    if (!h || !latents || !imgs || B < 1 || B > h->desc.max_batch) return SDFNMPC_E_INVALID;
    const size_t img_floats = (size_t)h->desc.height * h->desc.width;
    if (!h->h_img) SDF_CUDA_CHECK(cudaMalloc(&h->h_img, img_floats * h->desc.max_batch * sizeof(float)));
    if (!h->h_lat) SDF_CUDA_CHECK(cudaMalloc(&h->h_lat, (size_t)h->desc.size_latent * h->desc.max_batch * sizeof(float)));
    SDF_CUDA_CHECK(cudaMemcpyAsync(h->h_lat, latents, (size_t)h->desc.size_latent * B * sizeof(float), cudaMemcpyHostToDevice, 0));
    if (int rc = sdfnmpc_decoder_decode(h, h->h_lat, B, h->h_img, nullptr)) return rc;
    SDF_CUDA_CHECK(cudaMemcpyAsync(imgs, h->h_img, img_floats * B * sizeof(float), cudaMemcpyDeviceToHost, 0));
    SDF_CUDA_CHECK(cudaStreamSynchronize(0));
    return 0;
}
