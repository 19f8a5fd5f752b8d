// Device-side description of a packed NeuralDF (no kernels here; see sdf_kernels.cuh).
#pragma once
#include <cstddef>
#include <cuda_runtime.h>

namespace sdfnmpc {

enum SdfAct { SDF_ACT_SIN = 0, SDF_ACT_RELU = 1, SDF_ACT_SOFTPLUS = 2 };

struct SdfNetDev {
    int D, F, E, Ep;               // directions, frequencies, embedding size, padded (multiple of 8)
    int H1, H2, H3, H4;            // hidden sizes, padded to multiples of 8 (zero weights in the padding)
    int L;                         // latent size
    int res_state, res_latent;     // residual connection carries the embedding / the latent
    int act;
    float w0, max_df, b5;
    const float* dirs;             // [3][D]
    // forward operands, "transposed": [K][N] with N contiguous
    const float *W1e_t, *W2_t, *W3h_t, *W3e_t, *W4_t, *W5;
    // reverse operands, original orientation: [K = out][N = in]
    const float *W4, *W3h, *W3e, *W2, *W1e;
    const float *b2, *b4;
    // latent folding
    const float *W1z_t, *W3z_t, *b1, *b3;   // [L][H1], [L][H3]
};

constexpr int SDF_TM = 32;          // points per CTA
constexpr int SDF_THREADS = 256;

__host__ __device__ inline int sdf_max(int a, int b) { return a > b ? a : b; }
__host__ __device__ inline size_t sdf_smem_bytes(const SdfNetDev& n) {
    size_t rows = (size_t)n.Ep + n.H1 + n.H2 + n.H3 + n.H4 + sdf_max(sdf_max(n.H1, n.H3), n.Ep) + sdf_max(n.H2, n.H4);
    return rows * SDF_TM * sizeof(float) + SDF_TM * (3 * sizeof(float) + sizeof(int)) + 64;
}

}  // namespace sdfnmpc
