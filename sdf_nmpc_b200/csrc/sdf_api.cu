// C ABI for the neural-SDF evaluator (see include/sdfnmpc_b200.h).
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include "handles.h"
#include "sdf_kernels.cuh"
#include "sdf_tc_kernel.cuh"

namespace sdfnmpc {
std::atomic<uint64_t> g_launches{0};

int ensure_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        fprintf(stderr, "[sdfnmpc_b200] no CUDA device available (%s); this library has no CPU path\n",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return SDFNMPC_E_NODEVICE;
    }
    return 0;
}
}  // namespace sdfnmpc

using namespace sdfnmpc;

extern "C" const char* sdfnmpc_version(void) { return "sdfnmpc_b200 0.1 (sm_100a)"; }
extern "C" uint64_t sdfnmpc_launch_count(void) { return g_launches.load(); }
extern "C" int32_t sdfnmpc_sizeof(int32_t which) {
    switch (which) {
        case 0: return (int32_t)sizeof(sdfnmpc_sdf_desc);
        case 1: return (int32_t)sizeof(sdfnmpc_solver_desc);
        case 2: return (int32_t)sizeof(sdfnmpc_encoder_desc);
        case 3: return (int32_t)sizeof(sdfnmpc_encoder_weights);
        case 4: return (int32_t)sizeof(sdfnmpc_decoder_desc);
        case 5: return (int32_t)sizeof(sdfnmpc_decoder_weights);
        default: return -1;
    }
}

static inline int pad8(int v) { return (v + 31) / 32 * 32; }   // every feature dimension is padded to the 32-wide tensor-core chunk


// ------------------------------------------------------------------------------------------
// Tensor-core path: chunk program + weight chunks pre-split (hi | lo TF32) and pre-laid-out in the UMMA canonical
// 128-byte-swizzled K-major layout, so that one bulk copy lands a ready-to-use B operand in shared memory.
// ------------------------------------------------------------------------------------------
namespace {
inline float tf32_rna_host(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u + 0x1000u) & 0xFFFFE000u;       // round to nearest, ties away (cvt.rna.tf32.f32)
    float r;
    memcpy(&r, &u, 4);
    return r;
}

// B[n][k] for n < N, k < K (row-major accessor); appends K/32 chunks and returns their byte offsets
template <class Acc>
void pack_chunks(std::vector<float>& blob, int N, int K, Acc at, std::vector<uint32_t>& offs) {
    for (int k0 = 0; k0 < K; k0 += 32) {
        offs.push_back((uint32_t)(blob.size() * sizeof(float)));
        const size_t base = blob.size();
        blob.resize(base + (size_t)N * 32 * 2, 0.f);
        for (int n = 0; n < N; ++n)
            for (int k = 0; k < 32; ++k) {
                const float v = at(n, k0 + k);
                const float hi = tf32_rna_host(v), lo = tf32_rna_host(v - hi);
                // 128-byte-swizzled K-major: 8-row atoms of 1024 B, row = 32 floats, 16-byte unit index XOR row-in-atom
                const size_t idx = (size_t)(n >> 3) * 256 + (size_t)(n & 7) * 32 + (size_t)(((k >> 2) ^ (n & 7)) * 4) + (k & 3);
                blob[base + idx] = hi;
                blob[base + (size_t)N * 32 + idx] = lo;
            }
    }
}

int build_tc(sdfnmpc_sdf* h, const float* W1, const float* W2, const float* W3, const float* W4) {
    const SdfNetDev& n = h->net;
    const sdfnmpc_sdf_desc& d = h->desc;
    if (n.H1 > 256 || n.H2 > 256 || n.H3 > 128 || n.H4 > 64 || n.Ep > 128 || n.H1 % 32 || n.H2 % 32 || n.H3 % 32 || n.H4 % 32) return 0;
    const int E = n.E, L = n.L, h1 = d.hidden[0], h2 = d.hidden[1], h3 = d.hidden[2], h4 = d.hidden[3];
    const int in1 = E + L, in3 = h2 + (d.res_state ? E : 0) + (d.res_latent ? L : 0);
    std::vector<float> blob;
    std::vector<uint32_t> oW1, oW2, oW3h, oW3e, oW4, oW4T, oW3hT, oW3eT, oW2T, oW1T;
    auto z = [](bool ok, float v) { return ok ? v : 0.f; };
    pack_chunks(blob, n.H1, n.Ep, [&](int r, int k) { return z(r < h1 && k < E, (r < h1 && k < E) ? W1[(size_t)r * in1 + k] : 0.f); }, oW1);
    pack_chunks(blob, n.H2, n.H1, [&](int r, int k) { return (r < h2 && k < h1) ? W2[(size_t)r * h1 + k] : 0.f; }, oW2);
    pack_chunks(blob, n.H3, n.H2, [&](int r, int k) { return (r < h3 && k < h2) ? W3[(size_t)r * in3 + k] : 0.f; }, oW3h);
    if (n.res_state) pack_chunks(blob, n.H3, n.Ep, [&](int r, int k) { return (r < h3 && k < E) ? W3[(size_t)r * in3 + h2 + k] : 0.f; }, oW3e);
    pack_chunks(blob, n.H4, n.H3, [&](int r, int k) { return (r < h4 && k < h3) ? W4[(size_t)r * h3 + k] : 0.f; }, oW4);
    // reverse operands: B[n = input feature][k = output feature]
    pack_chunks(blob, n.H3, n.H4, [&](int r, int k) { return (k < h4 && r < h3) ? W4[(size_t)k * h3 + r] : 0.f; }, oW4T);
    pack_chunks(blob, n.H2, n.H3, [&](int r, int k) { return (k < h3 && r < h2) ? W3[(size_t)k * in3 + r] : 0.f; }, oW3hT);
    if (n.res_state) pack_chunks(blob, n.Ep, n.H3, [&](int r, int k) { return (k < h3 && r < E) ? W3[(size_t)k * in3 + h2 + r] : 0.f; }, oW3eT);
    pack_chunks(blob, n.H1, n.H2, [&](int r, int k) { return (k < h2 && r < h1) ? W2[(size_t)k * h1 + r] : 0.f; }, oW2T);
    pack_chunks(blob, n.Ep, n.H1, [&](int r, int k) { return (k < h1 && r < E) ? W1[(size_t)k * in1 + r] : 0.f; }, oW1T);

    std::vector<SdfTcOp> ops;
    auto add = [&](int prod, int src_col, int feat0, int layer, int wait_d, uint32_t w_off, int n_total, int d_col, int first, int last) {
        SdfTcOp o{};
        o.prod = prod; o.src_col = src_col; o.feat0 = feat0; o.layer = layer; o.wait_d = wait_d; o.pre_drain_ge3 = 0;
        o.w_off = w_off; o.w_bytes = (uint32_t)n_total * 32 * 4 * 2; o.n_total = n_total; o.n_piece = 1;
        o.d_col[0] = d_col; o.n_rows[0] = n_total; o.b_row0[0] = 0; o.d_col[1] = 0; o.n_rows[1] = 0; o.b_row0[1] = 0;
        o.first = first; o.last = last;
        ops.push_back(o);
    };
    const int c1 = n.Ep / 32, cH1 = n.H1 / 32, cH2 = n.H2 / 32, cH3 = n.H3 / 32, cH4 = n.H4 / 32;
    // forward
    // (layer field of an embedding chunk: 1 = first evaluation of the tile - values and cosines go to the cache -, 3 = the skip
    // connection's second use - read back)
    for (int j = 0; j < c1; ++j) add(TC_PROD_EMBED, 0, 32 * j, 1, 0, oW1[j], n.H1, 0, j == 0, j == c1 - 1);
    for (int j = 0; j < cH1; ++j) add(TC_PROD_ACT, 0 + 32 * j, 32 * j, 1, j == 0, oW2[j], n.H2, 256, j == 0, j == cH1 - 1);
    // layer 3 input [h2; e]: the skip-connection (embedding) chunks go FIRST - they do not read TMEM, so the producers build
    // them while the last MMAs of layer 2 drain, instead of idling at the layer boundary
    if (n.res_state)
        for (int j = 0; j < c1; ++j) add(TC_PROD_EMBED, 0, 32 * j, 3, 0, oW3e[j], n.H3, 0, j == 0, 0);
    for (int j = 0; j < cH2; ++j) add(TC_PROD_ACT, 256 + 32 * j, 32 * j, 2, j == 0, oW3h[j], n.H3, 0, (!n.res_state) && j == 0, j == cH2 - 1);
    for (int j = 0; j < cH3; ++j) add(TC_PROD_ACT, 0 + 32 * j, 32 * j, 3, j == 0, oW4[j], n.H4, 128, j == 0, j == cH3 - 1);
    // reverse
    for (int j = 0; j < cH4; ++j) add(TC_PROD_G4, 128 + 32 * j, 32 * j, 4, j == 0, oW4T[j], n.H3, 0, j == 0, j == cH4 - 1);
    for (int j = 0; j < cH3; ++j) add(TC_PROD_MUL, 0 + 32 * j, 32 * j, 3, j == 0, oW3hT[j], n.H2, 128, j == 0, (!n.res_state) && j == cH3 - 1);
    if (n.res_state)
        for (int j = 0; j < cH3; ++j) add(TC_PROD_MUL, 0 + 32 * j, 32 * j, 3, 0, oW3eT[j], n.Ep, 384, j == 0, j == cH3 - 1);
    for (int j = 0; j < cH2; ++j) {
        add(TC_PROD_MUL, 128 + 32 * j, 32 * j, 2, j == 0, oW2T[j], n.H1, 0, j == 0, j == cH2 - 1);
        SdfTcOp& o = ops.back();
        if (j == 0 && n.res_state) o.pre_drain_ge3 = 1;
        if (n.H1 > 128) {       // D_g1 does not fit in one free column range: rows 0..127 @0, the rest @384
            o.n_piece = 2;
            o.n_rows[0] = 128; o.d_col[1] = 384; o.n_rows[1] = n.H1 - 128; o.b_row0[1] = 128;
        }
    }
    // D_ge1 @256 (D_g2 is dead by then): columns 0..255 stay free for the NEXT tile's D1, whose first chunks the producers
    // build before they drain this tile's result (sdf_tc_kernel.cuh)
    for (int j = 0; j < cH1; ++j) add(TC_PROD_MUL, j < 4 ? 32 * j : 384 + 32 * (j - 4), 32 * j, 1, j == 0, oW1T[j], n.Ep, 256, j == 0, j == cH1 - 1);

    // accumulator completions: d_commit = index of the accumulator an op publishes, d_need = index a TMEM-reading op needs
    {
        int n_acc = 0, cur_need = -1;
        for (SdfTcOp& o : ops) {
            if (o.wait_d) cur_need = n_acc - 1;
            o.d_need = o.prod == TC_PROD_EMBED ? -1 : cur_need;
            o.d_commit = o.last ? n_acc++ : -1;
        }
        if (n_acc > TC_MAX_ACC || (int)ops.size() > TC_MAX_OPS) return 0;   // program does not fit: CUDA-core kernel only
        h->tc_n_acc = n_acc;
    }
    h->tc_n_ops = (int)ops.size();
    SDF_CUDA_CHECK(cudaMalloc(&h->tc_ops, ops.size() * sizeof(SdfTcOp)));
    SDF_CUDA_CHECK(cudaMemcpy(h->tc_ops, ops.data(), ops.size() * sizeof(SdfTcOp), cudaMemcpyHostToDevice));
    h->tc_ops_host = malloc(ops.size() * sizeof(SdfTcOp));
    if (!h->tc_ops_host) return SDFNMPC_E_INVALID;
    memcpy(h->tc_ops_host, ops.data(), ops.size() * sizeof(SdfTcOp));
    SDF_CUDA_CHECK(cudaMalloc(&h->tc_wblob, blob.size() * sizeof(float)));
    SDF_CUDA_CHECK(cudaMemcpy(h->tc_wblob, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice));
    h->tc_sc_off[0] = 0;
    h->tc_sc_off[1] = h->tc_sc_off[0] + n.H1 * TC_TM;
    h->tc_sc_off[2] = h->tc_sc_off[1] + n.H2 * TC_TM;
    h->tc_sc_off[3] = h->tc_sc_off[2] + n.H3 * TC_TM;
    h->tc_sc_off_emb = h->tc_sc_off[3] + n.Ep * TC_TM;                      // embedding cache: [tile parity][value | cos][Ep][128]
    h->tc_scratch_per_cta = (size_t)h->tc_sc_off_emb + (size_t)4 * n.Ep * TC_TM;
    SDF_CUDA_CHECK(cudaMalloc(&h->tc_scratch, h->tc_scratch_per_cta * sizeof(float) * h->sm_count));
    SDF_CUDA_CHECK(cudaFuncSetAttribute(sdf_eval_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    const char* env = getenv("SDFNMPC_SDF_IMPL");
    h->tc_ok = !(env && strcmp(env, "ffma") == 0);
    if (env && strcmp(env, "tc") == 0) h->tc_min_points = 1;
    return 0;
}
}  // namespace

extern "C" int sdfnmpc_sdf_create(sdfnmpc_sdf_t* out, const sdfnmpc_sdf_desc* d, const float* dirs,
                                  const float* W1, const float* b1, const float* W2, const float* b2,
                                  const float* W3, const float* b3, const float* W4, const float* b4,
                                  const float* W5, const float* b5) {
    if (!out || !d || !dirs || !W1 || !W5) return SDFNMPC_E_INVALID;
    if (d->n_dirs < 1 || d->nb_freqs < 0 || d->nb_freqs > 20 || d->size_latent < 0 || d->act < 0 || d->act > 2) return SDFNMPC_E_INVALID;
    for (int i = 0; i < 4; ++i)
        if (d->hidden[i] < 1) return SDFNMPC_E_INVALID;
    if (int rc = ensure_device()) return rc;
    auto* h = new sdfnmpc_sdf();
    h->desc = *d;
    const int D = d->n_dirs, F = d->nb_freqs, L = d->size_latent;
    const int E = 3 + 2 * D * F, Ep = pad8(E);
    const int h1 = d->hidden[0], h2 = d->hidden[1], h3 = d->hidden[2], h4 = d->hidden[3];
    const int H1 = pad8(h1), H2 = pad8(h2), H3 = pad8(h3), H4 = pad8(h4);
    const int in1 = E + L;
    const int in3 = h2 + (d->res_state ? E : 0) + (d->res_latent ? L : 0);
    const int offz3 = h2 + (d->res_state ? E : 0);
    // ---- pack on the host ----
    std::vector<float> host;
    auto reserve = [&](size_t n) { size_t o = host.size(); host.resize(o + ((n + 3) / 4) * 4, 0.f); return o; };
    const size_t o_dirs = reserve((size_t)3 * D);
    const size_t o_W1e_t = reserve((size_t)Ep * H1), o_W2_t = reserve((size_t)H1 * H2), o_W3h_t = reserve((size_t)H2 * H3);
    const size_t o_W3e_t = reserve((size_t)Ep * H3), o_W4_t = reserve((size_t)H3 * H4), o_W5 = reserve(H4);
    const size_t o_W4 = reserve((size_t)H4 * H3), o_W3h = reserve((size_t)H3 * H2), o_W3e = reserve((size_t)H3 * Ep);
    const size_t o_W2 = reserve((size_t)H2 * H1), o_W1e = reserve((size_t)H1 * Ep);
    const size_t o_b2 = reserve(H2), o_b4 = reserve(H4), o_b1 = reserve(H1), o_b3 = reserve(H3);
    const size_t o_W1z_t = reserve((size_t)(L > 0 ? L : 1) * H1), o_W3z_t = reserve((size_t)(L > 0 ? L : 1) * H3);
    memcpy(&host[o_dirs], dirs, sizeof(float) * 3 * D);
    for (int n = 0; n < h1; ++n) {
        for (int k = 0; k < E; ++k) {
            host[o_W1e_t + (size_t)k * H1 + n] = W1[(size_t)n * in1 + k];
            host[o_W1e + (size_t)n * Ep + k] = W1[(size_t)n * in1 + k];
        }
        for (int l = 0; l < L; ++l) host[o_W1z_t + (size_t)l * H1 + n] = W1[(size_t)n * in1 + E + l];
        host[o_b1 + n] = b1 ? b1[n] : 0.f;
    }
    for (int n = 0; n < h2; ++n) {
        for (int k = 0; k < h1; ++k) {
            host[o_W2_t + (size_t)k * H2 + n] = W2[(size_t)n * h1 + k];
            host[o_W2 + (size_t)n * H1 + k] = W2[(size_t)n * h1 + k];
        }
        host[o_b2 + n] = b2 ? b2[n] : 0.f;
    }
    for (int n = 0; n < h3; ++n) {
        for (int k = 0; k < h2; ++k) {
            host[o_W3h_t + (size_t)k * H3 + n] = W3[(size_t)n * in3 + k];
            host[o_W3h + (size_t)n * H2 + k] = W3[(size_t)n * in3 + k];
        }
        if (d->res_state)
            for (int k = 0; k < E; ++k) {
                host[o_W3e_t + (size_t)k * H3 + n] = W3[(size_t)n * in3 + h2 + k];
                host[o_W3e + (size_t)n * Ep + k] = W3[(size_t)n * in3 + h2 + k];
            }
        if (d->res_latent)
            for (int l = 0; l < L; ++l) host[o_W3z_t + (size_t)l * H3 + n] = W3[(size_t)n * in3 + offz3 + l];
        host[o_b3 + n] = b3 ? b3[n] : 0.f;
    }
    for (int n = 0; n < h4; ++n) {
        for (int k = 0; k < h3; ++k) {
            host[o_W4_t + (size_t)k * H4 + n] = W4[(size_t)n * h3 + k];
            host[o_W4 + (size_t)n * H3 + k] = W4[(size_t)n * h3 + k];
        }
        host[o_b4 + n] = b4 ? b4[n] : 0.f;
        host[o_W5 + n] = W5[n];
    }
    h->blob_floats = host.size();
    cudaError_t e = cudaMalloc(&h->blob, host.size() * sizeof(float));
    if (e != cudaSuccess) { delete h; return (int)e; }
    e = cudaMemcpy(h->blob, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(h->blob); delete h; return (int)e; }
    SdfNetDev& n = h->net;
    n.D = D; n.F = F; n.E = E; n.Ep = Ep; n.H1 = H1; n.H2 = H2; n.H3 = H3; n.H4 = H4; n.L = L;
    n.res_state = d->res_state ? 1 : 0; n.res_latent = d->res_latent ? 1 : 0; n.act = d->act;
    n.w0 = d->w0; n.max_df = d->max_df; n.b5 = b5 ? b5[0] : 0.f;
    const float* B = h->blob;
    n.dirs = B + o_dirs; n.W1e_t = B + o_W1e_t; n.W2_t = B + o_W2_t; n.W3h_t = B + o_W3h_t; n.W3e_t = B + o_W3e_t;
    n.W4_t = B + o_W4_t; n.W5 = B + o_W5; n.W4 = B + o_W4; n.W3h = B + o_W3h; n.W3e = B + o_W3e; n.W2 = B + o_W2;
    n.W1e = B + o_W1e; n.b2 = B + o_b2; n.b4 = B + o_b4; n.b1 = B + o_b1; n.b3 = B + o_b3;
    n.W1z_t = B + o_W1z_t; n.W3z_t = B + o_W3z_t;
    h->smem = sdf_smem_bytes(n);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, dev);
    h->sm_count = prop.multiProcessorCount;
    if (h->smem > (size_t)prop.sharedMemPerBlockOptin) {
        fprintf(stderr, "[sdfnmpc_b200] network too wide for the shared-memory tiling (%zu B needed)\n", h->smem);
        cudaFree(h->blob); delete h; return SDFNMPC_E_INVALID;
    }
    e = cudaFuncSetAttribute(sdf_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem);
    if (e != cudaSuccess) { cudaFree(h->blob); delete h; return (int)e; }
    if (int rc = build_tc(h, W1, W2, W3, W4)) { sdfnmpc_sdf_destroy(h); return rc; }
    *out = h;
    return 0;
}

extern "C" int sdfnmpc_sdf_destroy(sdfnmpc_sdf_t h) {
    if (!h) return 0;
    cudaFree(h->blob); cudaFree(h->b1f); cudaFree(h->b3f);
    cudaFree(h->h_pts); cudaFree(h->h_out); cudaFree(h->h_lat); cudaFree(h->h_b1f); cudaFree(h->h_b3f);
    cudaFree(h->tc_ops); cudaFree(h->tc_wblob); cudaFree(h->tc_scratch);
    free(h->tc_ops_host);
    delete h;
    return 0;
}

extern "C" int sdfnmpc_sdf_set_latents(sdfnmpc_sdf_t h, const float* z, int32_t n_inst, void* stream) {
    if (!h || n_inst < 1 || (!z && h->net.L > 0)) return SDFNMPC_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    if (n_inst > h->cap_inst) {
        cudaFree(h->b1f); cudaFree(h->b3f);
        h->b1f = h->b3f = nullptr; h->cap_inst = 0;
        SDF_CUDA_CHECK(cudaMalloc(&h->b1f, (size_t)n_inst * h->net.H1 * sizeof(float)));
        SDF_CUDA_CHECK(cudaMalloc(&h->b3f, (size_t)n_inst * h->net.H3 * sizeof(float)));
        h->cap_inst = n_inst;
    }
    const int ncol = h->net.H1 + h->net.H3;
    dim3 grid((ncol + 127) / 128, n_inst);
    sdf_fold_latent_kernel<<<grid, 128, 0, st>>>(h->net, z, n_inst, h->b1f, h->b3f);
    count_launch();
    SDF_CUDA_CHECK(cudaGetLastError());
    h->n_inst = n_inst;
    return 0;
}

extern "C" int sdfnmpc_sdf_eval(sdfnmpc_sdf_t h, const float* pts, const int32_t* pt2inst, int32_t pts_per_inst,
                                int64_t n, float* sdf, float* grad, void* stream) {
    if (!h || n < 0 || n > 0x7fffffff || (!pt2inst && pts_per_inst < 1)) return SDFNMPC_E_INVALID;
    if (h->n_inst < 1) return SDFNMPC_E_STATE;
    if (n == 0) return 0;
    if (!pts) return SDFNMPC_E_INVALID;
    if (!pt2inst && (n + pts_per_inst - 1) / pts_per_inst > h->n_inst) return SDFNMPC_E_INVALID;
    if (n == 0) return 0;
    if (h->tc_ok && grad && sdf && n >= h->tc_min_points) {
        SdfTcParams P;
        P.net = h->net; P.ops = (const SdfTcOp*)h->tc_ops; P.n_ops = h->tc_n_ops; P.wblob = h->tc_wblob;
        memcpy(P.opsv, h->tc_ops_host, (size_t)h->tc_n_ops * sizeof(SdfTcOp));
        P.scratch = h->tc_scratch; P.scratch_floats_per_cta = h->tc_scratch_per_cta;
        P.sc_off_c1 = h->tc_sc_off[0]; P.sc_off_c2 = h->tc_sc_off[1]; P.sc_off_c3 = h->tc_sc_off[2]; P.sc_off_ge3 = h->tc_sc_off[3];
        P.sc_off_emb = h->tc_sc_off_emb;
        P.b1f = h->b1f; P.b3f = h->b3f; P.pts = pts; P.pt2inst = pt2inst; P.pts_per_inst = pts_per_inst; P.n = (int)n;
        P.sdf_out = sdf; P.grad_out = grad; P.n_acc = h->tc_n_acc;
        P.trace = nullptr;
        if (const char* tp = getenv("SDFNMPC_TC_TRACE")) {
            static long long* tbuf = nullptr;
            if (!tbuf) cudaMalloc(&tbuf, 8 * TC_MAX_OPS * sizeof(long long));
            P.trace = tbuf;
            (void)tp;
        }
        const int64_t tc_tiles = (n + TC_TM - 1) / TC_TM;
        const int tc_grid = (int)(tc_tiles < h->sm_count ? tc_tiles : h->sm_count);
        sdf_eval_tc_kernel<<<tc_grid, TC_THREADS, TC_SMEM_BYTES, (cudaStream_t)stream>>>(P);
        count_launch();
        SDF_CUDA_CHECK(cudaGetLastError());
        if (P.trace) {
            std::vector<long long> tr(8 * TC_MAX_OPS);
            cudaMemcpy(tr.data(), P.trace, tr.size() * sizeof(long long), cudaMemcpyDeviceToHost);
            FILE* f = fopen(getenv("SDFNMPC_TC_TRACE"), "w");
            if (f) {
                for (int i = 0; i < h->tc_n_ops; ++i) {
                    for (int k = 0; k < 8; ++k) fprintf(f, "%lld ", tr[i * 8 + k] - tr[0]);
                    fprintf(f, "\n");
                }
                fclose(f);
            }
        }
        return 0;
    }
    const int64_t tiles = (n + SDF_TM - 1) / SDF_TM;
    const int grid = (int)(tiles < (int64_t)h->sm_count * 8 ? tiles : (int64_t)h->sm_count * 8);
    sdf_eval_kernel<<<grid, SDF_THREADS, h->smem, (cudaStream_t)stream>>>(h->net, pts, pt2inst, pts_per_inst, h->b1f,
                                                                         h->b3f, (int)n, sdf, grad);
    count_launch();
    SDF_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int sdfnmpc_sdf_eval_host(sdfnmpc_sdf_t h, const float* pts, const float* latent, int64_t n, float* sdf, float* grad) {
    if (!h || !pts || n < 1 || (!latent && h->net.L > 0)) return SDFNMPC_E_INVALID;
    if ((size_t)n > h->h_cap) {
        cudaFree(h->h_pts); cudaFree(h->h_out);
        h->h_pts = h->h_out = nullptr; h->h_cap = 0;
        SDF_CUDA_CHECK(cudaMalloc(&h->h_pts, (size_t)n * 3 * sizeof(float)));
        SDF_CUDA_CHECK(cudaMalloc(&h->h_out, (size_t)n * 4 * sizeof(float)));
        h->h_cap = (size_t)n;
    }
    if (!h->h_lat) SDF_CUDA_CHECK(cudaMalloc(&h->h_lat, (size_t)(h->net.L > 0 ? h->net.L : 1) * sizeof(float)));
    SDF_CUDA_CHECK(cudaMemcpyAsync(h->h_pts, pts, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, 0));
    if (h->net.L > 0) SDF_CUDA_CHECK(cudaMemcpyAsync(h->h_lat, latent, (size_t)h->net.L * sizeof(float), cudaMemcpyHostToDevice, 0));
    // The host path folds its latent into its OWN bias buffers: a handle that is bound to the CasADi entry points and also
    // serves a batched solver keeps the solver's per-instance biases and instance count.
    if (!h->h_b1f) SDF_CUDA_CHECK(cudaMalloc(&h->h_b1f, (size_t)h->net.H1 * sizeof(float)));
    if (!h->h_b3f) SDF_CUDA_CHECK(cudaMalloc(&h->h_b3f, (size_t)h->net.H3 * sizeof(float)));
    float *sb1 = h->b1f, *sb3 = h->b3f;
    const int scap = h->cap_inst, sn = h->n_inst;
    h->b1f = h->h_b1f; h->b3f = h->h_b3f; h->cap_inst = 1; h->n_inst = 0;
    int rc_ = sdfnmpc_sdf_set_latents(h, h->h_lat, 1, nullptr);
    if (!rc_) rc_ = sdfnmpc_sdf_eval(h, h->h_pts, nullptr, (int32_t)n, n, h->h_out, grad ? h->h_out + n : nullptr, nullptr);
    h->b1f = sb1; h->b3f = sb3; h->cap_inst = scap; h->n_inst = sn;
    if (rc_) return rc_;
    if (sdf) SDF_CUDA_CHECK(cudaMemcpyAsync(sdf, h->h_out, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, 0));
    if (grad) SDF_CUDA_CHECK(cudaMemcpyAsync(grad, h->h_out + n, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, 0));
    SDF_CUDA_CHECK(cudaStreamSynchronize(0));
    return 0;
}

// ------------------------------------------------------------------------------------------
// CasADi external-function ABI: the symbols of L4CasADi's libsdf_l4c.so
// (sdf_nmpc/gen_model.py:38-39 -> name='sdf_l4c', with_jacobian=True; ocp.py:100-102).
// double in / double out, column-major dense, batch 1, model evaluated in fp32.
// ------------------------------------------------------------------------------------------
namespace {
std::mutex g_l4c_mu;
sdfnmpc_sdf_t g_l4c = nullptr;
bool g_l4c_owned = false;
std::vector<float> g_l4c_in;
casadi_int g_sp_in[3], g_sp_scalar[3] = {1, 1, 1}, g_sp_jac[3];

int l4c_ready() {
    if (g_l4c) return 0;
    const char* path = getenv("SDFNMPC_L4C_BLOB");
    if (!path) {
        fprintf(stderr, "[sdfnmpc_b200] sdf_l4c called before sdfnmpc_l4c_bind() and SDFNMPC_L4C_BLOB is not set\n");
        return SDFNMPC_E_STATE;
    }
    return sdfnmpc_l4c_load_blob(path);
}

int l4c_eval(const double* i0, double* val, double* jac) {
    std::lock_guard<std::mutex> lk(g_l4c_mu);
    if (int rc = l4c_ready()) return rc;
    const int L = g_l4c->net.L, nin = 3 + L;
    g_l4c_in.resize(nin);
    for (int i = 0; i < nin; ++i) g_l4c_in[i] = (float)i0[i];   // L4CasADi casts the CasADi doubles to float32
    float out[4] = {0, 0, 0, 0};
    int rc = sdfnmpc_sdf_eval_host(g_l4c, g_l4c_in.data(), g_l4c_in.data() + 3, 1, out, jac ? out + 1 : nullptr);
    if (rc) return rc;
    if (val) *val = (double)out[0];
    if (jac) {
        for (int i = 0; i < nin; ++i) jac[i] = 0.0;   // latent columns: p is not a decision variable
        jac[0] = out[1]; jac[1] = out[2]; jac[2] = out[3];
    }
    return 0;
}
}  // namespace

extern "C" int sdfnmpc_l4c_bind(sdfnmpc_sdf_t h) {
    std::lock_guard<std::mutex> lk(g_l4c_mu);
    if (g_l4c && g_l4c_owned) sdfnmpc_sdf_destroy(g_l4c);
    g_l4c = h;
    g_l4c_owned = false;
    return 0;
}

// Blob: int32 magic 'SDFB', int32 version, sdfnmpc_sdf_desc, then float arrays in the order of
// sdfnmpc_sdf_create (dirs, W1, b1, ..., W5, b5).  Written by sdf_nmpc_b200.weights.save_sdf_blob.
extern "C" int sdfnmpc_l4c_load_blob(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "[sdfnmpc_b200] cannot open %s\n", path); return SDFNMPC_E_INVALID; }
    int32_t hdr[2];
    sdfnmpc_sdf_desc d;
    if (fread(hdr, sizeof(hdr), 1, f) != 1 || hdr[0] != 0x42464453 || hdr[1] != 1 || fread(&d, sizeof(d), 1, f) != 1) { fclose(f); return SDFNMPC_E_INVALID; }
    // the descriptor comes from a file: range-check it before it sizes any allocation (sdfnmpc_sdf_create repeats the checks)
    auto bad = [](int v, int lo, int hi) { return v < lo || v > hi; };
    if (bad(d.n_dirs, 1, 64) || bad(d.nb_freqs, 0, 16) || bad(d.size_latent, 0, 4096) || bad(d.hidden[0], 1, 4096) ||
        bad(d.hidden[1], 1, 4096) || bad(d.hidden[2], 1, 4096) || bad(d.hidden[3], 1, 4096)) { fclose(f); return SDFNMPC_E_INVALID; }
    try {
    const int E = 3 + 2 * d.n_dirs * d.nb_freqs, L = d.size_latent;
    const int in3 = d.hidden[1] + (d.res_state ? E : 0) + (d.res_latent ? L : 0);
    const size_t sz[11] = {(size_t)3 * d.n_dirs, (size_t)d.hidden[0] * (E + L), (size_t)d.hidden[0],
                           (size_t)d.hidden[1] * d.hidden[0], (size_t)d.hidden[1], (size_t)d.hidden[2] * in3,
                           (size_t)d.hidden[2], (size_t)d.hidden[3] * d.hidden[2], (size_t)d.hidden[3],
                           (size_t)d.hidden[3], 1};
    std::vector<std::vector<float>> a(11);
    for (int i = 0; i < 11; ++i) {
        a[i].resize(sz[i]);
        if (fread(a[i].data(), sizeof(float), sz[i], f) != sz[i]) { fclose(f); return SDFNMPC_E_INVALID; }
    }
    fclose(f);
    f = nullptr;
    sdfnmpc_sdf_t h = nullptr;
    int rc = sdfnmpc_sdf_create(&h, &d, a[0].data(), a[1].data(), a[2].data(), a[3].data(), a[4].data(), a[5].data(),
                                a[6].data(), a[7].data(), a[8].data(), a[9].data(), a[10].data());
    if (rc) return rc;
    if (g_l4c && g_l4c_owned) sdfnmpc_sdf_destroy(g_l4c);
    g_l4c = h;
    g_l4c_owned = true;
    return 0;
    } catch (...) {      // nothing may unwind through an extern "C" boundary
        if (f) fclose(f);
        return SDFNMPC_E_INVALID;
    }
}

static int l4c_nin() { return g_l4c ? 3 + g_l4c->net.L : 131; }

#define SDFNMPC_CASADI_COMMON(F)                                                   \
    extern "C" int F##_alloc_mem(void) { return 0; }                               \
    extern "C" int F##_init_mem(int) { return 0; }                                 \
    extern "C" void F##_free_mem(int) {}                                           \
    extern "C" int F##_checkout(void) { return 0; }                                \
    extern "C" void F##_release(int) {}                                            \
    extern "C" void F##_incref(void) {}                                            \
    extern "C" void F##_decref(void) {}                                            \
    extern "C" double F##_default_in(casadi_int) { return 0.0; }

SDFNMPC_CASADI_COMMON(sdf_l4c)
SDFNMPC_CASADI_COMMON(jac_sdf_l4c)

extern "C" int sdf_l4c(const double** arg, double** res, casadi_int*, double*, int) {
    if (!arg || !arg[0]) return 1;
    double v = 0.0;
    if (l4c_eval(arg[0], &v, nullptr)) return 1;
    if (res && res[0]) res[0][0] = v;
    return 0;
}
extern "C" casadi_int sdf_l4c_n_in(void) { return 1; }
extern "C" casadi_int sdf_l4c_n_out(void) { return 1; }
extern "C" const char* sdf_l4c_name_in(casadi_int i) { return i == 0 ? "i0" : nullptr; }
extern "C" const char* sdf_l4c_name_out(casadi_int i) { return i == 0 ? "o0" : nullptr; }
extern "C" const casadi_int* sdf_l4c_sparsity_in(casadi_int i) {
    if (i != 0) return nullptr;
    g_sp_in[0] = l4c_nin(); g_sp_in[1] = 1; g_sp_in[2] = 1;   // dense: {nrow, ncol, 1}
    return g_sp_in;
}
extern "C" const casadi_int* sdf_l4c_sparsity_out(casadi_int i) { return i == 0 ? g_sp_scalar : nullptr; }
extern "C" int sdf_l4c_work(casadi_int* a, casadi_int* r, casadi_int* iw, casadi_int* w) {
    if (a) *a = 1; if (r) *r = 1; if (iw) *iw = 0; if (w) *w = 0;
    return 0;
}

extern "C" int jac_sdf_l4c(const double** arg, double** res, casadi_int*, double*, int) {
    if (!arg || !arg[0]) return 1;
    if (!res || !res[0]) return 0;
    return l4c_eval(arg[0], nullptr, res[0]) ? 1 : 0;
}
extern "C" casadi_int jac_sdf_l4c_n_in(void) { return 2; }
extern "C" casadi_int jac_sdf_l4c_n_out(void) { return 1; }
extern "C" const char* jac_sdf_l4c_name_in(casadi_int i) { return i == 0 ? "i0" : (i == 1 ? "out_o0" : nullptr); }
extern "C" const char* jac_sdf_l4c_name_out(casadi_int i) { return i == 0 ? "jac_o0_i0" : nullptr; }
extern "C" const casadi_int* jac_sdf_l4c_sparsity_in(casadi_int i) {
    if (i == 0) return sdf_l4c_sparsity_in(0);
    return i == 1 ? g_sp_scalar : nullptr;
}
extern "C" const casadi_int* jac_sdf_l4c_sparsity_out(casadi_int i) {
    if (i != 0) return nullptr;
    g_sp_jac[0] = 1; g_sp_jac[1] = l4c_nin(); g_sp_jac[2] = 1;
    return g_sp_jac;
}
extern "C" int jac_sdf_l4c_work(casadi_int* a, casadi_int* r, casadi_int* iw, casadi_int* w) {
    if (a) *a = 2; if (r) *r = 1; if (iw) *iw = 0; if (w) *w = 0;
    return 0;
}
