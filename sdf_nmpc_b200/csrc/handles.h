// Internal handle definitions shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <vector>
#include "../../include/sdfnmpc_b200.h"
#include "common.cuh"
#include "sdf_net.h"

namespace sdfnmpc {
extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
int ensure_device();   // 0 if a CUDA device is usable, SDFNMPC_E_NODEVICE otherwise (no CPU path)
}  // namespace sdfnmpc

struct sdfnmpc_sdf {
    sdfnmpc_sdf_desc desc;
    sdfnmpc::SdfNetDev net;          // device pointers into `blob`
    float* blob = nullptr;           // one allocation with every packed weight
    size_t blob_floats = 0;
    float* b1f = nullptr;            // [cap_inst][H1] folded biases
    float* b3f = nullptr;            // [cap_inst][H3]
    int cap_inst = 0, n_inst = 0;
    float *h_b1f = nullptr, *h_b3f = nullptr;   // folded biases of the host / CasADi path (its own: the batched state above is not touched)
    size_t smem = 0;
    int sm_count = 148;
    // tensor-core (tcgen05) path: chunk program, packed hi|lo weight chunks, per-CTA scratch
    bool tc_ok = false;
    void* tc_ops = nullptr;          // SdfTcOp[] (device)
    void* tc_ops_host = nullptr;     // the same program on the host (passed by value in the kernel parameters)
    int tc_n_ops = 0, tc_n_acc = 0;
    unsigned char* tc_wblob = nullptr;
    float* tc_scratch = nullptr;
    size_t tc_scratch_per_cta = 0;
    int tc_sc_off[4] = {0, 0, 0, 0};
    int tc_sc_off_emb = 0;
    int tc_min_points = 1;           // below this the CUDA-core kernel is used.  A mostly padded 128-point tile still takes 82 us
                                     // against 225 us of the CUDA-core kernel for the 21 query points of one instance (measured)
    // scratch for the host-convenience entry points
    float *h_pts = nullptr, *h_out = nullptr, *h_lat = nullptr;
    size_t h_cap = 0;
};
