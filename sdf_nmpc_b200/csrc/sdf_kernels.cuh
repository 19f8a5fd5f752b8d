// Fused neural-SDF forward + analytic input-gradient kernels (fp32, CUDA cores).
//
// Replaces, for a whole batch of query points, what the reference executes one point
// at a time on the CPU through L4CasADi/libtorch (sdf_nmpc/gen_model.py:38-39,60):
//   NeuralDF.forward        sdf_nmpc/network/neural_df.py:91-103
//   PositionEmbedding       sdf_nmpc/utils/embeddings.py:106-111
//   Sine                    sdf_nmpc/utils/activation.py:12-13
//   jacrev of the above     (columns 0..2 of jac_sdf_l4c; the latent columns are not needed
//                            because p is not a decision variable - SURVEY.md 8a-1)
// The latent enters only through per-instance folded biases b1' = W1z z + b1,
// b3' = W3z z + b3 (sdf_fold_latent_kernel), so the per-point work is the
// point-dependent 154 kMAC forward + the same reverse (arch B).
//
// Tiling: one CTA = 32 points x 256 threads.  Activations live in shared memory as
// [feature][point] (128-byte rows, conflict-free LDS.128 broadcast); each thread owns a
// 4(point) x NT(feature) register tile; weights stream from L2 through the read-only path
// with a register double buffer (each weight row is read exactly once per CTA).
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"
#include "sdf_net.h"

namespace sdfnmpc {

// ------------------------------------------------------------------------------------------
// b1f[i][n] = b1[n] + sum_l z[i][l] W1z_t[l][n]  (and b3f likewise when the residual carries z)
// ------------------------------------------------------------------------------------------
__global__ void sdf_fold_latent_kernel(SdfNetDev net, const float* __restrict__ z, int n_inst,
                                       float* __restrict__ b1f, float* __restrict__ b3f) {
    const int i = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const int ncol = net.H1 + net.H3;
    if (i >= n_inst || n >= ncol) return;
    const float* zi = z + (size_t)i * net.L;
    if (n < net.H1) {
        float s = net.b1[n];
        for (int l = 0; l < net.L; ++l) s = fmaf(zi[l], net.W1z_t[(size_t)l * net.H1 + n], s);
        b1f[(size_t)i * net.H1 + n] = s;
    } else {
        const int m = n - net.H1;
        float s = net.b3[m];
        if (net.res_latent)
            for (int l = 0; l < net.L; ++l) s = fmaf(zi[l], net.W3z_t[(size_t)l * net.H3 + m], s);
        b3f[(size_t)i * net.H3 + m] = s;
    }
}

namespace sdfdetail {

__device__ __forceinline__ void act_pair(int act, float w0, float a, float& h, float& dh) {
    if (act == SDF_ACT_SIN) {
        float s, c;
        sincosf(w0 * a, &s, &c);
        h = s;
        dh = w0 * c;
    } else if (act == SDF_ACT_RELU) {
        h = fmaxf(a, 0.f);
        dh = a > 0.f ? 1.f : 0.f;
    } else {
        h = a > 20.f ? a : log1pf(expf(a));
        dh = 1.f / (1.f + expf(-a));
    }
}

template <int NT> struct VecLoad;
template <> struct VecLoad<8> {
    static __device__ __forceinline__ void ld(float* d, const float* p) {
        float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
    }
};
template <> struct VecLoad<4> {
    static __device__ __forceinline__ void ld(float* d, const float* p) {
        float4 a = __ldg(reinterpret_cast<const float4*>(p));
        d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w;
    }
};
template <> struct VecLoad<2> {
    static __device__ __forceinline__ void ld(float* d, const float* p) {
        float2 a = __ldg(reinterpret_cast<const float2*>(p));
        d[0] = a.x; d[1] = a.y;
    }
};

// acc[mi][ni] += sum_k A[k][m0+mi] * Bt[k][n0+ni];  K multiple of 4; A in shared memory ([K][32]).
template <int NT>
__device__ __forceinline__ void gemm_acc(float (&acc)[4][NT], const float* __restrict__ A, int K,
                                         const float* __restrict__ Bt, int ldb, int n0, int m0) {
    constexpr int KU = 4;
    float bc[KU][NT], bn[KU][NT];
    const float* bp = Bt + n0;
#pragma unroll
    for (int u = 0; u < KU; ++u) VecLoad<NT>::ld(bc[u], bp + (size_t)u * ldb);
    for (int k = 0; k < K; k += KU) {
        const bool more = k + KU < K;
        if (more) {
#pragma unroll
            for (int u = 0; u < KU; ++u) VecLoad<NT>::ld(bn[u], bp + (size_t)(k + KU + u) * ldb);
        }
#pragma unroll
        for (int u = 0; u < KU; ++u) {
            const float4 a = *reinterpret_cast<const float4*>(A + (k + u) * SDF_TM + m0);
#pragma unroll
            for (int ni = 0; ni < NT; ++ni) {
                acc[0][ni] = fmaf(a.x, bc[u][ni], acc[0][ni]);
                acc[1][ni] = fmaf(a.y, bc[u][ni], acc[1][ni]);
                acc[2][ni] = fmaf(a.z, bc[u][ni], acc[2][ni]);
                acc[3][ni] = fmaf(a.w, bc[u][ni], acc[3][ni]);
            }
        }
        if (more) {
#pragma unroll
            for (int u = 0; u < KU; ++u)
#pragma unroll
                for (int ni = 0; ni < NT; ++ni) bc[u][ni] = bn[u][ni];
        }
    }
}

enum Epilogue {
    EPI_ACT = 0,     // a = acc + bias; out_h = act(a); out_c = act'(a)
    EPI_MUL = 1,     // out_c *= acc           (reverse sweep through an activation)
    EPI_STORE = 2,   // out_h = acc
    EPI_ADD = 3      // out_h += acc
};

struct LayerArgs {
    int N;                       // output features (multiple of 8)
    int nseg;                    // K-segments of the (concatenated) input
    const float* A[2];           // shared [K][32]
    const float* Bt[2];          // global [K][N]
    int K[2];
    int mode;
    float* out_h;                // shared [N][32]
    float* out_c;                // shared [N][32]
    const float* bias;           // global; bias[inst*bias_stride + n]
    int bias_stride;             // 0: one bias vector for all points
    const int* inst;             // shared [32]
    float w0;
    int act;
};

template <int NT>
__device__ __noinline__ void layer(const LayerArgs& L) {
    const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
    const int m0 = tx * 4;
    for (int nb = 0; nb < L.N; nb += 32 * NT) {
        const int n0 = nb + ty * NT;
        if (n0 >= L.N) continue;
        float acc[4][NT];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < NT; ++ni) acc[mi][ni] = 0.f;
        for (int s = 0; s < L.nseg; ++s) gemm_acc<NT>(acc, L.A[s], L.K[s], L.Bt[s], L.N, n0, m0);
        if (L.mode == EPI_ACT) {
            int io[4];
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) io[mi] = L.inst[m0 + mi] * L.bias_stride;
#pragma unroll
            for (int ni = 0; ni < NT; ++ni) {
                const int nn = n0 + ni;
                float hv[4], dv[4];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) act_pair(L.act, L.w0, acc[mi][ni] + __ldg(L.bias + io[mi] + nn), hv[mi], dv[mi]);
                *reinterpret_cast<float4*>(L.out_h + nn * SDF_TM + m0) = make_float4(hv[0], hv[1], hv[2], hv[3]);
                *reinterpret_cast<float4*>(L.out_c + nn * SDF_TM + m0) = make_float4(dv[0], dv[1], dv[2], dv[3]);
            }
        } else if (L.mode == EPI_MUL) {
#pragma unroll
            for (int ni = 0; ni < NT; ++ni) {
                float4* p = reinterpret_cast<float4*>(L.out_c + (n0 + ni) * SDF_TM + m0);
                const float4 c = *p;
                *p = make_float4(acc[0][ni] * c.x, acc[1][ni] * c.y, acc[2][ni] * c.z, acc[3][ni] * c.w);
            }
        } else if (L.mode == EPI_STORE) {
#pragma unroll
            for (int ni = 0; ni < NT; ++ni)
                *reinterpret_cast<float4*>(L.out_h + (n0 + ni) * SDF_TM + m0) = make_float4(acc[0][ni], acc[1][ni], acc[2][ni], acc[3][ni]);
        } else {
#pragma unroll
            for (int ni = 0; ni < NT; ++ni) {
                float4* p = reinterpret_cast<float4*>(L.out_h + (n0 + ni) * SDF_TM + m0);
                const float4 c = *p;
                *p = make_float4(acc[0][ni] + c.x, acc[1][ni] + c.y, acc[2][ni] + c.z, acc[3][ni] + c.w);
            }
        }
    }
}

__device__ __forceinline__ void layer_auto(const LayerArgs& L) {
    if (L.N > 128) layer<8>(L);
    else if (L.N > 64) layer<4>(L);
    else layer<2>(L);
}

}  // namespace sdfdetail

// ------------------------------------------------------------------------------------------
// pts [n][3] fp32;  inst = pt2inst ? pt2inst[i] : i / pts_per_inst;  b1f [n_inst][H1], b3f [n_inst][H3]
// out: sdf [n] (raw network output), grad [n][3] (d sdf / d point), either may be null
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SDF_THREADS, 1)
sdf_eval_kernel(SdfNetDev net, const float* __restrict__ pts, const int* __restrict__ pt2inst, int pts_per_inst,
                const float* __restrict__ b1f, const float* __restrict__ b3f, int n,
                float* __restrict__ sdf_out, float* __restrict__ grad_out) {
    using namespace sdfdetail;
    extern __shared__ __align__(16) float smem[];
    const int TM = SDF_TM;
    float* e_s = smem;                                  // [Ep][32]
    float* c1 = e_s + net.Ep * TM;                      // [H1][32]  act'(a1) -> g1
    float* c2 = c1 + net.H1 * TM;
    float* c3 = c2 + net.H2 * TM;
    float* c4 = c3 + net.H3 * TM;
    float* bx = c4 + net.H4 * TM;                       // h1 -> h3 -> g_e
    float* by = bx + sdf_max(sdf_max(net.H1, net.H3), net.Ep) * TM;   // h2 -> h4
    float* pt_s = by + sdf_max(net.H2, net.H4) * TM;    // [32][3]
    int* inst_s = reinterpret_cast<int*>(pt_s + TM * 3);

    const int tid = threadIdx.x;
    const int DF = net.D * net.F;
    const float HALF_PI = 1.57079632679489661923f;

    LayerArgs L;
    L.inst = inst_s;
    L.w0 = net.w0;
    L.act = net.act;

    for (int tile = blockIdx.x; (long long)tile * TM < n; tile += gridDim.x) {
        const int base = tile * TM;
        __syncthreads();   // previous tile fully consumed
        if (tid < TM) {
            const int i = base + tid;
            const int ic = i < n ? i : n - 1;
            pt_s[tid * 3 + 0] = pts[(size_t)ic * 3 + 0];
            pt_s[tid * 3 + 1] = pts[(size_t)ic * 3 + 1];
            pt_s[tid * 3 + 2] = pts[(size_t)ic * 3 + 2];
            inst_s[tid] = pt2inst ? pt2inst[ic] : ic / pts_per_inst;
        }
        __syncthreads();
        // ---- embedding  e = [s, sin(xb), sin(xb + pi/2)],  xb[d*F+f] = (s . dirs[:,d]) * 2^f ----
        for (int idx = tid; idx < net.Ep * TM; idx += SDF_THREADS) {
            const int j = idx / TM, m = idx % TM;
            const float sx = pt_s[m * 3], sy = pt_s[m * 3 + 1], sz = pt_s[m * 3 + 2];
            float v = 0.f;
            if (j < 3) {
                v = pt_s[m * 3 + j];
            } else if (j < net.E) {
                const int q = (j - 3) % DF, d = q / net.F, f = q % net.F;
                const float proj = fmaf(sz, net.dirs[2 * net.D + d], fmaf(sy, net.dirs[net.D + d], sx * net.dirs[d]));
                float xb = proj * (float)(1 << f);
                if (j - 3 >= DF) xb = xb + HALF_PI;       // the reference adds pi/2 in fp32 and takes the sine
                v = sinf(xb);
            }
            e_s[idx] = v;
        }
        __syncthreads();
        // ---- layer 1: a1 = W1e e + b1'(latent) ----
        L.mode = EPI_ACT;
        L.N = net.H1; L.nseg = 1; L.A[0] = e_s; L.K[0] = net.Ep; L.Bt[0] = net.W1e_t;
        L.out_h = bx; L.out_c = c1; L.bias = b1f; L.bias_stride = net.H1;
        layer_auto(L);
        __syncthreads();
        // ---- layer 2 ----
        L.N = net.H2; L.A[0] = bx; L.K[0] = net.H1; L.Bt[0] = net.W2_t;
        L.out_h = by; L.out_c = c2; L.bias = net.b2; L.bias_stride = 0;
        layer_auto(L);
        __syncthreads();
        // ---- layer 3: input [h2 ; e] (+ latent folded into b3') ----
        L.N = net.H3; L.nseg = net.res_state ? 2 : 1; L.A[0] = by; L.K[0] = net.H2; L.Bt[0] = net.W3h_t;
        L.A[1] = e_s; L.K[1] = net.Ep; L.Bt[1] = net.W3e_t;
        L.out_h = bx; L.out_c = c3; L.bias = b3f; L.bias_stride = net.H3;
        layer_auto(L);
        __syncthreads();
        // ---- layer 4 ----
        L.N = net.H4; L.nseg = 1; L.A[0] = bx; L.K[0] = net.H3; L.Bt[0] = net.W4_t;
        L.out_h = by; L.out_c = c4; L.bias = net.b4; L.bias_stride = 0;
        layer_auto(L);
        __syncthreads();
        // ---- output layer (8 threads per point) ----
        {
            const int m = tid >> 3, part = tid & 7;
            float s = 0.f;
            for (int nn = part; nn < net.H4; nn += 8) s = fmaf(net.W5[nn], by[nn * TM + m], s);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            if (part == 0 && sdf_out && base + m < n) sdf_out[base + m] = s + net.b5;
        }
        if (grad_out == nullptr) continue;
        // ---- seed of the reverse sweep: g4 = W5 .* act'(a4) ----
        for (int idx = tid; idx < net.H4 * TM; idx += SDF_THREADS) c4[idx] *= net.W5[idx / TM];
        __syncthreads();
        // ---- reverse layer 4:  g3 = (W4^T g4) .* act'(a3) ----
        L.mode = EPI_MUL;
        L.N = net.H3; L.A[0] = c4; L.K[0] = net.H4; L.Bt[0] = net.W4; L.out_c = c3;
        layer_auto(L);
        __syncthreads();
        // ---- reverse layer 3:  [g_h2 ; g_e] = W3^T g3 ;  g2 = g_h2 .* act'(a2) ----
        float* ge = bx;       // h3 is dead
        L.N = net.H2; L.A[0] = c3; L.K[0] = net.H3; L.Bt[0] = net.W3h; L.out_c = c2;
        layer_auto(L);
        if (net.res_state) {
            L.mode = EPI_STORE;
            L.N = net.Ep; L.Bt[0] = net.W3e; L.out_h = ge;
            layer_auto(L);
        } else {
            for (int idx = tid; idx < net.Ep * TM; idx += SDF_THREADS) ge[idx] = 0.f;
        }
        __syncthreads();
        // ---- reverse layer 2:  g1 = (W2^T g2) .* act'(a1) ----
        L.mode = EPI_MUL;
        L.N = net.H1; L.A[0] = c2; L.K[0] = net.H2; L.Bt[0] = net.W2; L.out_c = c1;
        layer_auto(L);
        __syncthreads();
        // ---- reverse layer 1:  g_e += W1e^T g1 ----
        L.mode = EPI_ADD;
        L.N = net.Ep; L.A[0] = c1; L.K[0] = net.H1; L.Bt[0] = net.W1e; L.out_h = ge;
        layer_auto(L);
        __syncthreads();
        // ---- through the embedding: grad = g_e[0:3] + dirs * sum_f 2^f (g_sin cos(xb) + g_cos cos(xb + pi/2)) ----
        {
            const int m = tid >> 3, part = tid & 7;
            const float sx = pt_s[m * 3], sy = pt_s[m * 3 + 1], sz = pt_s[m * 3 + 2];
            float gx = 0.f, gy = 0.f, gz = 0.f;
            for (int q = part; q < DF; q += 8) {
                const int d = q / net.F, f = q % net.F;
                const float dx = net.dirs[d], dy = net.dirs[net.D + d], dz = net.dirs[2 * net.D + d];
                const float proj = fmaf(sz, dz, fmaf(sy, dy, sx * dx));
                const float sc = (float)(1 << f);
                const float xb = proj * sc;
                const float gq = ge[(3 + q) * TM + m] * cosf(xb) + ge[(3 + DF + q) * TM + m] * cosf(xb + HALF_PI);
                const float t = gq * sc;
                gx = fmaf(t, dx, gx); gy = fmaf(t, dy, gy); gz = fmaf(t, dz, gz);
            }
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
                gx += __shfl_xor_sync(0xffffffffu, gx, o);
                gy += __shfl_xor_sync(0xffffffffu, gy, o);
                gz += __shfl_xor_sync(0xffffffffu, gz, o);
            }
            if (part == 0 && base + m < n) {
                grad_out[(size_t)(base + m) * 3 + 0] = gx + ge[0 * TM + m];
                grad_out[(size_t)(base + m) * 3 + 1] = gy + ge[1 * TM + m];
                grad_out[(size_t)(base + m) * 3 + 2] = gz + ge[2 * TM + m];
            }
        }
    }
}

}  // namespace sdfnmpc
