// Host emulation of the device algorithms (development aid for the CPU-only test-suite).
// Compiles the SAME headers the CUDA kernels include, with one lane per "warp".
// Built by tests/test_host_emulation.py:  g++ -O2 -shared -fPIC -DSDFNMPC_HOST_EMUL -x c++ emul.cpp
#include <cstddef>
#include <cstring>
#include <vector>
#include "../../sdf_nmpc_b200/csrc/linearize_core.cuh"
#include "../../sdf_nmpc_b200/csrc/qp_core.cuh"
#include "../../sdf_nmpc_b200/csrc/term_lift.cuh"

using namespace sdfnmpc;

// lifted terminal rows (term_lift.cuh): state of the next emul_linearize / emul_qp calls; n_real < 0 = plain OCP
static int g_n_real = -1, g_stab = 0;
static const double *g_nlh = nullptr, *g_nuh = nullptr, *g_nzs = nullptr, *g_nZs = nullptr;
extern "C" void emul_set_lift(int n_real, int stab, const double* nlh, const double* nuh, const double* nzs, const double* nZs) {
    g_n_real = n_real; g_stab = stab; g_nlh = nlh; g_nuh = nuh; g_nzs = nzs; g_nZs = nZs;
}

static ModelConst model_const(int model, const double* mcv) {
    ModelConst mc;
    mc.g = mcv[0]; mc.gamma = mcv[1]; mc.roll = mcv[2]; mc.pitch = mcv[3]; mc.wz = mcv[4]; mc.mass = mcv[5];
    mc.cam_off[0] = mcv[6]; mc.cam_off[1] = mcv[7]; mc.cam_off[2] = mcv[8];
    mc.ax = mcv[9]; mc.ay = mcv[10]; mc.az = mcv[11]; mc.tau_roll = mcv[12]; mc.tau_pitch = mcv[13];
    if (model == MODEL_PROPS) {          // mcv continues: wp_max, Gf / m [12], J^-1 Gt [12], inertia [3]
        mc.wp_max = mcv[14];
        for (int i = 0; i < 12; ++i) { mc.Gf_m[i] = mcv[15 + i]; mc.JGt[i] = mcv[27 + i]; }
        for (int i = 0; i < 3; ++i) mc.inertia[i] = mcv[39 + i];
    }
    return mc;
}

extern "C" void emul_term_lift(int B, int N, int n_real, int n_lift, int degree, const double* coeffs, int n_coeffs, const double* mcv,
                               const double* pose, const double* flag, double max_df, const float* sdf_d, const float* sdf_g,
                               double* xbar, double* AB, double* bdyn, double* C, double* hval) {
    const ModelConst mc = model_const(MODEL_ATT, mcv);
    TermConst tc;
    tc.n_lift = n_lift; tc.degree = degree;
    for (int i = 0; i < TERM_MAX_COEFFS; ++i) tc.coeffs[i] = i < n_coeffs ? coeffs[i] : 0.0;
    for (int b = 0; b < B; ++b)
        term_lift_instance(mc, tc, N, n_real, pose + (size_t)b * 12, flag[b], max_df, sdf_d[(size_t)b * (N + 1) + n_real],
                           sdf_g + ((size_t)b * (N + 1) + n_real) * 3, xbar + (size_t)b * (N + 1) * ATT_NX,
                           AB + (size_t)b * N * ATT_NX * ATT_NW, bdyn + (size_t)b * N * ATT_NX, C + (size_t)b * (N + 1) * ATT_NH * ATT_NW,
                           hval + (size_t)b * (N + 1) * ATT_NH);
}

extern "C" void emul_linearize(int model, int nyN, int B, int N, const double* mcv, double* xbar, const double* ubar,
                               const double* yref, const double* yrefN, const double* qd, const double* pose,
                               const double* flag, const double* dt, double* AB, double* bdyn, double* J,
                               double* yres, double* C, double* hval, float* qpts, const double* x0) {
    const ModelConst mc = model_const(model, mcv);
    LinIn in{xbar, ubar, yref, yrefN, qd, pose, flag, dt, nyN, x0};
    in.n_real = g_n_real; in.stab = g_stab;
    LinOut out{AB, bdyn, J, yres, C, hval, qpts, nullptr};
    const int nw = model == MODEL_PROPS ? ModelDims<MODEL_PROPS>::NW : ATT_NW;
    for (int b = 0; b < B; ++b)
        for (int k = 0; k <= N; ++k)
            for (int d = 0; d < nw; ++d) {
                if (model == MODEL_ATT) linearize_dir<MODEL_ATT>(mc, N, in, out, b, k, d);
                else if (model == MODEL_ATT_TAU) linearize_dir<MODEL_ATT_TAU>(mc, N, in, out, b, k, d);
                else if (model == MODEL_PROPS) linearize_dir<MODEL_PROPS>(mc, N, in, out, b, k, d);
                else linearize_dir<MODEL_ACC>(mc, N, in, out, b, k, d);
            }
}

template <int NX, int NU, int NY, int NH>
static void emul_qp_t(int B, int N, int k0, int max_iter, int nyN, double sdf_cost_w, double lm, double tol, double tol_res,
                        double max_df, const double* flag, const double* AB, const double* bdyn, const double* J,
                        const double* yres, double* C, double* hval, const double* W, const double* WN,
                        const double* dt, const double* lbu, const double* ubu, const double* lh, const double* uh,
                        const double* zs, const double* Zs, const float* sdf_d, const float* sdf_g,
                        const double* pose, const double* x0, double* xbar, double* ubar, double* u0,
                        double* lam_hard, double* lam_soft, int* status, int* iters) {
    constexpr int NW = NX + NU;
    using Shape = QpShape<NX, NU, NY, NH>;
    using Prob = QpProb<NX, NU, NY, NH>;
    std::vector<double> work(Shape::work_doubles(N));
    std::vector<double> sm(Shape::smem_doubles(N));
    Lanes L{0};
    for (int b = 0; b < B; ++b) {
        Prob P;
        P.N = N; P.k0 = k0; P.max_iter = max_iter; P.nyN = nyN;
        P.n_real = g_n_real >= 0 ? g_n_real : N; P.nlh = g_nlh; P.nuh = g_nuh; P.nzs = g_nzs; P.nZs = g_nZs;
        const int NRu = P.n_real;               // the caller's W keeps the OCP's horizon
        P.lm = lm; P.tol = tol; P.tol_res = tol_res; P.max_df = max_df; P.flag = flag[b];
        P.sdf_cost_w = sdf_cost_w;
        P.AB = AB + (size_t)b * N * NX * NW;
        P.ABt = nullptr;
        P.bdyn = bdyn + (size_t)b * N * NX;
        P.J = J + (size_t)b * (N + 1) * NY * NW;
        P.yres = yres + (size_t)b * (N + 1) * NY;
        P.C = C + (size_t)b * (N + 1) * NH * NW;
        P.hval = hval + (size_t)b * (N + 1) * NH;
        P.W = W + (size_t)b * NRu * NY;
        P.WN = WN + (size_t)b * nyN;
        P.dt = dt; P.lbu = lbu; P.ubu = ubu; P.lh = lh; P.uh = uh; P.zs = zs; P.Zs = Zs;
        P.sdf_d = sdf_d + (size_t)b * (N + 1);
        P.sdf_g = sdf_g + (size_t)b * (N + 1) * 3;
        P.pose = pose + (size_t)b * 12;
        P.x0 = x0 + (size_t)b * NX;
        P.xbar = xbar + (size_t)b * (N + 1) * NX;
        P.ubar = ubar + (size_t)b * N * NU;
        P.work = work.data();
        P.u0 = u0 + (size_t)b * NU;
        P.lam_hard = lam_hard + (size_t)b * N * 2 * NU;
        P.lam_soft = lam_soft + (size_t)b * (N + 1) * 2 * NH;
        P.res_out = nullptr;
        P.trace = nullptr;
        P.debug = 0;
        P.status = status + b;
        P.iters = iters + b;
        qp_solve_instance<NX, NU, NY, NH>(L, P, sm.data());
    }
}

#define EMUL_QP_ARGS B, N, k0, max_iter, nyN, sdf_cost_w, lm, tol, tol_res, max_df, flag, AB, bdyn, J, yres, C, hval, W, WN, dt, lbu, ubu, lh, uh, \
                     zs, Zs, sdf_d, sdf_g, pose, x0, xbar, ubar, u0, lam_hard, lam_soft, status, iters
#define EMUL_QP_PARAMS int B, int N, int k0, int max_iter, int nyN, double sdf_cost_w, double lm, double tol, double tol_res,                   \
                       double max_df, const double* flag, const double* AB, const double* bdyn, const double* J, const double* yres, double* C, \
                       double* hval, const double* W, const double* WN, const double* dt, const double* lbu, const double* ubu,                  \
                       const double* lh, const double* uh, const double* zs, const double* Zs, const float* sdf_d, const float* sdf_g,          \
                       const double* pose, const double* x0, double* xbar, double* ubar, double* u0, double* lam_hard, double* lam_soft,        \
                       int* status, int* iters
extern "C" void emul_qp(EMUL_QP_PARAMS) { emul_qp_t<ATT_NX, ATT_NU, ATT_NY, ATT_NH>(EMUL_QP_ARGS); }
// model 'props': 13 states + one pad entry (model_att.cuh, ModelDims<MODEL_PROPS>)
extern "C" void emul_qp_props(EMUL_QP_PARAMS) {
    using D = ModelDims<MODEL_PROPS>;
    emul_qp_t<D::NX, D::NU, D::NY, D::NH>(EMUL_QP_ARGS);
}
