// Batched OCP-QP solver: structure-exploiting (Riccati) Mehrotra predictor-corrector
// interior-point method with analytically eliminated soft-constraint slacks, fp64,
// one cooperating group of lanes (a warp on the GPU) per NMPC instance.
//
// Role in the reference: the QP sub-problem acados hands to HPIPM at every SQP-RTI
// iteration (sdf_nmpc/ocp.py:109-120: SQP_RTI, FULL_CONDENSING_HPIPM, ROBUST, <=100
// iterations).  The reference condenses and runs a dense IPM; this solver keeps the
// stage structure (backward Riccati recursion + forward roll-out), which yields the
// same unique optimum of the strictly convex QP (Gauss-Newton + Levenberg-Marquardt,
// ocp.py:119-120).  QP semantics (cost/slack scaling by Ts, LM term, [u;x] ordering,
// soft rows lh <= h + s_l, h - s_u <= uh, no h rows at node 0): SURVEY.md 8a-5.
//
// Formulation.  Per node k the variable is w_k = [du_k; dx_k].  Every inequality is a
// one-sided row  g_i . w_k - d_i (+ s_i) = t_i >= 0  with multiplier lam_i; soft rows
// carry a slack s_i >= 0 (multiplier mu_i) with penalty z_i s_i + 1/2 Z_i s_i^2.  Slacks,
// t and the multipliers are eliminated from the Newton system row by row, which adds
// Theta~_i g_i g_i' to the stage Hessian and coef_i g_i to the stage right-hand side; the
// remaining equality-constrained QP in the STEP (delta form: right-hand sides are residuals,
// so steps keep full relative accuracy near the solution) is solved by the Riccati recursion.
//
// Work mapping (v2).  Row phases: lane k owns node k and keeps the node's (t, lam, s, mu) in
// registers for the whole solve - no index arithmetic, no scratch traffic.  Matrix phases
// (Riccati): lanes split the entries of the small dense products staged in shared memory.
// The stationarity / dynamics residuals of a Newton method on a QP contract exactly by
// (1 - alpha) per step, so they are kept as (initial residual) x (scalar phi) and verified
// once, explicitly, at the end (costates recovered by one backward sweep).
//
// Assumption encoded in HC0/HCN: the nonlinear rows h depend on the state only through the
// columns [HC0, HC0+HCN) of w (the position), which holds for every h of the reference
// (FoV rows and SDF row depend on W_p_B only: cost_const_helpers.py:60-65, gen_model.py:46-51).
#pragma once
#include "common.cuh"

namespace sdfnmpc {

template <int NX, int NU, int NY, int NH>
struct QpShape {
    static constexpr int NW = NX + NU;
    static constexpr int TRI = NW * (NW + 1) / 2;
    static constexpr int NCORR = NU + 6 + NU + 3;   // du, Dp(sym 3x3), gu, gp
    static constexpr int NPROD = 2 * NU + 4 * NH;   // second-order products per node
    SDF_HD static int kp(int N) { return N + 2; }   // padded node count of the [item][node] arrays
    // per-stage data staged asynchronously into shared memory: AB | H | rv0 | rd0 | PbM | L | K
    static constexpr int ST_AB = 0, ST_H = ST_AB + NX * NW, ST_RV = ST_H + NW * NW, ST_RD = ST_RV + NW, ST_PB = ST_RD + NX,
                         ST_L = ST_PB + NX, ST_K = ST_L + NU * NU, STAGE_DOUBLES = (ST_K + NU * NX + 1) / 2 * 2;
    // doubles of per-instance global scratch
    SDF_HD static size_t work_doubles(int N) {
        size_t n1 = (size_t)(N + 1);
        return n1 * NW * NW + n1 * NW      // H, g
               + n1 * NW + n1 * NX         // rv0, rdyn0
               + (size_t)N * (NU * NX + NU * NU + NX)   // K, Linv, PbM
               + n1 * NX;                  // costates (final verification)
    }
    // doubles of per-lane-group shared memory
    SDF_HD static int smem_doubles(int N) {
        return NX * NX + 2 * NX * (NW + 1) + NW * NW + NW + NU * NX + NU + 2 * NX   // Riccati workspace
               + 2 * NW * kp(N)                                                   // W (iterate), dW (step)
               + (NCORR + NPROD) * kp(N) + 8
               + 2 * STAGE_DOUBLES;                                               // double-buffered per-stage data
    }
};

template <int NX, int NU, int NY, int NH>
struct QpProb {            // all pointers are offset to ONE instance
    int N, k0, max_iter, nyN;
    double lm, tol, tol_res, max_df, flag;
    const double* AB;      // [N][NX][NW]
    const double* bdyn;    // [N][NX]
    const double* J;       // [N+1][NY][NW]
    const double* yres;    // [N+1][NY]
    double* C;             // [N+1][NH][NW]  (sdf row written here)
    double* hval;          // [N+1][NH]
    const double* W;       // [N][NY]
    const double* WN;      // [nyN]
    const double* dt;      // [N]
    const double *lbu, *ubu, *lh, *uh, *zs, *Zs;
    const float* sdf_d;    // [N+1]
    const float* sdf_g;    // [N+1][3]
    const double* pose;    // [12]
    const double* x0;      // [NX]
    double* xbar;          // [N+1][NX]   updated in place (full step)
    double* ubar;          // [N][NU]
    double* work;          // work_doubles(N)
    // outputs
    double* u0;            // [NU]
    double* lam_hard;      // [N][2NU] or null
    double* lam_soft;      // [N+1][2NH] or null
    double* res_out;       // [2] or null: verified KKT residual (inf-norm), inf-norm of the full step [du; dx]
    int* status;           // 0 ok, 2 max-iter, 4 NaN / factorisation failure (acados codes)
    int* iters;
};

namespace qpdetail {

template <int NW>
struct TriTable {
    unsigned char i[NW * (NW + 1) / 2], j[NW * (NW + 1) / 2];
    constexpr TriTable() : i(), j() {
        int e = 0;
        for (int a = 0; a < NW; ++a)
            for (int b = 0; b <= a; ++b) { i[e] = (unsigned char)a; j[e] = (unsigned char)b; ++e; }
    }
};

SDF_HD double rcp(double x) { return 1.0 / x; }
SDF_HD double step_to_boundary(double x, double dx) { return dx < 0.0 ? -x / dx : 1e30; }
SDF_HD int sym3(int a, int b) { return a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a; }

// Everything the elimination needs from one row; evaluated by ONE function so that the two places that use
// r_p / rho (right-hand side and step) see bit-identical values.
struct HardEval { double r_p, Th, coef; };
struct SoftEval { double r_p, r_s, Th, D, Tht, rho, coef; };

#if defined(__CUDA_ARCH__)
#define SDF_NOINLINE __device__ __noinline__
#else
#define SDF_NOINLINE inline
#endif

SDF_NOINLINE HardEval hard_eval(double t, double lam, double gw, double d, double corr) {
    HardEval r;
    const double rt = rcp(t);
    r.r_p = t - (gw - d);
    r.Th = lam * rt;
    const double kap = corr - lam * t;            // corr = 0 (affine) or tau - dlam_aff*dt_aff
    r.coef = -kap * rt - r.Th * r.r_p;
    return r;
}

SDF_NOINLINE SoftEval soft_eval(double t, double lam, double s, double mu, double gw, double d, double z, double Z,
                                double corr_t, double corr_s) {
    SoftEval r;
    const double rt = rcp(t), rs = rcp(s);
    r.r_p = t - (gw - d + s);
    r.r_s = z + Z * s - lam - mu;
    r.Th = lam * rt;
    const double zs = Z + mu * rs;
    r.D = zs + r.Th;
    const double rD = rcp(r.D);
    r.Tht = r.Th * zs * rD;
    const double kap = corr_t - lam * t, nu = corr_s - mu * s;
    r.rho = -r.r_s + kap * rt + nu * rs + r.Th * r.r_p;
    r.coef = -kap * rt + r.Th * (r.rho * rD - r.r_p);
    return r;
}

}  // namespace qpdetail

template <int NX, int NU, int NY, int NH, int NFIX = 0, int HC0 = NU, int HCN = 3>
SDF_HD void qp_solve_instance(const Lanes& L, const QpProb<NX, NU, NY, NH>& P, double* sm) {
    using namespace qpdetail;
    using S = QpShape<NX, NU, NY, NH>;
    static_assert(HCN == 3, "the position block is 3 wide");
    constexpr int NW = S::NW, NWB = NW + 1, TRI = S::TRI, NCORR = S::NCORR, NPROD = S::NPROD;
    constexpr TriTable<NW> tri{};
    const int N = NFIX > 0 ? NFIX : P.N;          // NFIX: horizon known at compile time (all strides become immediates)
    const int k0 = P.k0, KP = S::kp(N);
    const int NL = Lanes::count();
    const int lane = L.lane;
    const int ncomp = N * 2 * NU + 2 * (N + 1 - k0) * 2 * NH;

    // ---- global scratch ---------------------------------------------------------------------
    double* p = P.work;
    auto take = [&](size_t n) { double* r = p; p += n; return r; };
    const size_t n1 = (size_t)(N + 1);
    double* H = take(n1 * NW * NW);
    double* g = take(n1 * NW);
    double* rv0 = take(n1 * NW);       // stationarity residual at the initial point (scaled by phi afterwards)
    double* rd0 = take(n1 * NX);       // dynamics residual at the initial point: rd0[k+1] = AB_k w_k + b_k - w_x,k+1
    double* Kg = take((size_t)N * NU * NX);
    double* Lg = take((size_t)N * NU * NU);   // Cholesky factor of Ruu, diagonal stored INVERTED
    double* PbM = take((size_t)N * NX);
    double* pig = take(n1 * NX);

    // ---- shared workspace -----------------------------------------------------------------------
    double* Psm = sm;                    // [NX][NX]
    double* ABb = Psm + NX * NX;         // [NX][NW+1]
    double* PAB = ABb + NX * NWB;        // [NX][NW+1]
    double* Msm = PAB + NX * NWB;        // [NW][NW]
    double* rv = Msm + NW * NW;          // [NW]
    double* Ksm = rv + NW;               // [NU][NX]
    double* kfs = Ksm + NU * NX;         // [NU]
    double* pv = kfs + NU;               // [NX]
    double* vv = pv + NX;                // [NX]
    double* Wsm = vv + NX;               // [NW][KP]   iterate w
    double* dW = Wsm + NW * KP;          // [NW][KP]   step
    double* corr = dW + NW * KP;         // [NCORR][KP] du(NU) Dp(6) gu(NU) gp(3) per node
    double* prod = corr + NCORR * KP;    // [NPROD][KP] affine second-order products
    double* stg = prod + NPROD * KP + ((NPROD * KP) & 1);   // 2 x STAGE_DOUBLES (16-byte aligned)
    static_assert(NX % 2 == 0 && NW % 2 == 0 && NU % 2 == 0, "stage_async moves 16-byte units");

    auto scale = [&](int k) { return k < N ? P.dt[k] : 1.0; };
    auto wrow = [&](int k, int r) { return k < N ? P.W[k * NY + r] : (r < P.nyN ? P.WN[r] : 0.0); };

    // ---- prologue: sdf row of C / h, Gauss-Newton Hessian and gradient ----------------------
    for (int k = k0 + lane; k <= N; k += NL) {
        const double* R = P.pose + 3;
        double gx = P.sdf_g[k * 3 + 0], gy = P.sdf_g[k * 3 + 1], gz = P.sdf_g[k * 3 + 2];
        double* Cr = P.C + ((size_t)k * NH + (NH - 1)) * NW;
        for (int jj = 0; jj < 3; ++jj) Cr[HC0 + jj] = P.flag * (R[jj * 3 + 0] * gx + R[jj * 3 + 1] * gy + R[jj * 3 + 2] * gz);
        P.hval[k * NH + (NH - 1)] = P.flag * (double)P.sdf_d[k] + (1.0 - P.flag) * P.max_df;
    }
    for (int e = lane; e < (N + 1) * TRI; e += NL) {
        const int k = e / TRI, t = e - k * TRI, i = tri.i[t], j = tri.j[t];
        const double* Jk = P.J + (size_t)k * NY * NW;
        double s = 0.0;
        for (int r = 0; r < NY; ++r) s += Jk[r * NW + i] * wrow(k, r) * Jk[r * NW + j];
        if (i == j) s += P.lm;
        s *= scale(k);
        H[(size_t)k * NW * NW + i * NW + j] = s;
        H[(size_t)k * NW * NW + j * NW + i] = s;
    }
    for (int e = lane; e < (N + 1) * NW; e += NL) {
        const int k = e / NW, i = e - k * NW;
        const double* Jk = P.J + (size_t)k * NY * NW;
        const double* yr = P.yres + (size_t)k * NY;
        double s = 0.0;
        for (int r = 0; r < NY; ++r) s += Jk[r * NW + i] * wrow(k, r) * yr[r];
        g[e] = scale(k) * s;
        const double w0 = (k == 0 && i >= NU) ? P.x0[i - NU] - P.xbar[i - NU] : 0.0;
        Wsm[i * KP + k] = w0;
        dW[i * KP + k] = 0.0;
    }
    L.sync();

    // ---- per-lane node state (lane k <-> node k) ------------------------------------------------
    double tb[2 * NU], lb[2 * NU], ts[2 * NH], ls[2 * NH], ss[2 * NH], ms[2 * NH];
    double dtb[2 * NU], dlb[2 * NU], dts[2 * NH], dls[2 * NH], dss[2 * NH], dms[2 * NH];   // step of the node's rows

    // helpers reading the node's constants (L1-resident after the first touch)
    auto hard_d = [&](int kk, int a) {
        const int j = a < NU ? a : a - NU;
        const double ub = P.ubar[kk * NU + j];
        return a < NU ? P.lbu[j] - ub : ub - P.ubu[j];
    };
    auto soft_d = [&](int kk, int a) {
        const int r = a < NH ? a : a - NH;
        const double hv = P.hval[kk * NH + r];
        return a < NH ? P.lh[r] - hv : hv - P.uh[r];
    };
    auto c_pos = [&](int kk, int r, int c) { return P.C[((size_t)kk * NH + r) * NW + HC0 + c]; };

// entry loop over the lanes of the group, fully unrolled with predication so that the independent dot-product chains of
// one lane interleave (the dynamic-trip-count form serialised them: `wait` stalls) - constant bounds only
#define SDF_LANES_FOR(v, X)                                                                                       \
    _Pragma("unroll") for (int v##_it = 0; v##_it < ((X) + SDF_WARP_LANES - 1) / SDF_WARP_LANES; ++v##_it)        \
        if (const int v = lane + v##_it * SDF_WARP_LANES; v < (X))
#if defined(__CUDA_ARCH__)
#define SDF_NODE_LOOP(kk) for (int kk = lane; kk <= N; kk += 32)
#else
#define SDF_NODE_LOOP(kk) for (int kk = 0; kk <= N; ++kk)
#endif
#if !defined(__CUDA_ARCH__)
    // host emulation: one lane plays every node in turn, so the "registers" become per-node arrays
    double* TB = new double[(size_t)(N + 1) * (4 * NU + 8 * NH)];
    double* TD = new double[(size_t)(N + 1) * (4 * NU + 8 * NH)];
#define SDF_LOAD_DELTA(kk)                                                                         \
    {                                                                                              \
        double* q_ = TD + (size_t)(kk) * (4 * NU + 8 * NH);                                        \
        for (int a = 0; a < 2 * NU; ++a) { dtb[a] = q_[a]; dlb[a] = q_[2 * NU + a]; }              \
        for (int a = 0; a < 2 * NH; ++a) { dts[a] = q_[4 * NU + a]; dls[a] = q_[4 * NU + 2 * NH + a]; \
            dss[a] = q_[4 * NU + 4 * NH + a]; dms[a] = q_[4 * NU + 6 * NH + a]; }                  \
    }
#define SDF_STORE_DELTA(kk)                                                                        \
    {                                                                                              \
        double* q_ = TD + (size_t)(kk) * (4 * NU + 8 * NH);                                        \
        for (int a = 0; a < 2 * NU; ++a) { q_[a] = dtb[a]; q_[2 * NU + a] = dlb[a]; }              \
        for (int a = 0; a < 2 * NH; ++a) { q_[4 * NU + a] = dts[a]; q_[4 * NU + 2 * NH + a] = dls[a]; \
            q_[4 * NU + 4 * NH + a] = dss[a]; q_[4 * NU + 6 * NH + a] = dms[a]; }                  \
    }
#define SDF_LOAD_STATE(kk)                                                                         \
    {                                                                                              \
        double* q_ = TB + (size_t)(kk) * (4 * NU + 8 * NH);                                        \
        for (int a = 0; a < 2 * NU; ++a) { tb[a] = q_[a]; lb[a] = q_[2 * NU + a]; }                \
        for (int a = 0; a < 2 * NH; ++a) { ts[a] = q_[4 * NU + a]; ls[a] = q_[4 * NU + 2 * NH + a]; \
            ss[a] = q_[4 * NU + 4 * NH + a]; ms[a] = q_[4 * NU + 6 * NH + a]; }                    \
    }
#define SDF_STORE_STATE(kk)                                                                        \
    {                                                                                              \
        double* q_ = TB + (size_t)(kk) * (4 * NU + 8 * NH);                                        \
        for (int a = 0; a < 2 * NU; ++a) { q_[a] = tb[a]; q_[2 * NU + a] = lb[a]; }                \
        for (int a = 0; a < 2 * NH; ++a) { q_[4 * NU + a] = ts[a]; q_[4 * NU + 2 * NH + a] = ls[a]; \
            q_[4 * NU + 4 * NH + a] = ss[a]; q_[4 * NU + 6 * NH + a] = ms[a]; }                    \
    }
#else
#define SDF_LOAD_STATE(kk)
#define SDF_STORE_STATE(kk)
#define SDF_LOAD_DELTA(kk)
#define SDF_STORE_DELTA(kk)
#endif

    // ---- initial point and initial residuals ------------------------------------------------
    SDF_NODE_LOOP(kk) {
#pragma unroll
        for (int a = 0; a < 2 * NU; ++a) {
            double t = 1.0, l = 1.0;
            if (kk < N) {
                t = fmax(-hard_d(kk, a), 0.5);          // w_u = 0
                l = fmax(1.0 / t, 1e-3);
            }
            tb[a] = t; lb[a] = l;
        }
#pragma unroll
        for (int a = 0; a < 2 * NH; ++a) {
            const int r = a < NH ? a : a - NH;
            double t = 1.0, l = 1.0, s0 = 0.5;
            if (kk >= k0) {
                const double z = P.zs[r] * scale(kk), Z = P.Zs[r] * scale(kk);
                double cw = 0.0;
                for (int c = 0; c < 3; ++c) cw += c_pos(kk, r, c) * Wsm[(HC0 + c) * KP + kk];
                const double gw = a < NH ? cw : -cw;
                t = fmax(gw - soft_d(kk, a) + s0, 0.5);
                l = fmax(0.5 * (z + Z * s0), 1e-3);
            }
            ts[a] = t; ss[a] = s0; ls[a] = l; ms[a] = l;
        }
        SDF_STORE_STATE(kk);
        // rv0_k = H_k w_k + g_k - sum lam_i gvec_i   (w = 0 except dx_0; costates 0)
        for (int i = 0; i < NW; ++i) {
            double s = g[(size_t)kk * NW + i];
            if (kk == 0)
                for (int j = NU; j < NW; ++j) s += H[i * NW + j] * Wsm[j * KP];
            if (kk < N && i < NU) s -= lb[i] - lb[NU + i];
            if (kk >= k0 && i >= HC0 && i < HC0 + 3)
                for (int r = 0; r < NH; ++r) s -= (ls[r] - ls[NH + r]) * c_pos(kk, r, i - HC0);
            if ((kk == 0 && i >= NU) || (kk == N && i < NU)) s = 0.0;
            rv0[(size_t)kk * NW + i] = s;
        }
        // rd0_{k} (k >= 1) = AB_{k-1} w_{k-1} + b_{k-1} - w_x,k ; only w_x,0 is non-zero
        for (int i = 0; i < NX; ++i) {
            double s = 0.0;
            if (kk >= 1) {
                s = P.bdyn[(kk - 1) * NX + i];
                if (kk == 1)
                    for (int j = NU; j < NW; ++j) s += P.AB[(size_t)i * NW + j] * Wsm[j * KP];
            }
            rd0[(size_t)kk * NX + i] = s;
        }
    }
    L.sync();
    double res_v0 = 0.0;
    for (int e = lane; e < (N + 1) * NW; e += NL) res_v0 = fmax(res_v0, fabs(rv0[e]));
    for (int e = lane; e < (N + 1) * NX; e += NL) res_v0 = fmax(res_v0, fabs(rd0[e]));
    res_v0 = L.max(res_v0);

    int status = 2, it = 0;
    bool safe = false;              // conservative mode after a stall: more centering, no second-order term
    double cm_prev1 = 1e300, cm_prev2 = 1e300;
    double phi = 1.0;               // product of (1 - alpha): scale of the stationarity / dynamics residuals
    double res_track = 0.0;
    for (;; ++it) {
        double tau = 0.0, mubar = 0.0;
        bool stop = false;
        for (int pass = 0; pass < 2 && !stop; ++pass) {
            // =========== row elimination (lane = node): corrections of the stage Hessian / rhs ===========
            double res_rows = 0.0, cmax = 0.0, csum = 0.0;
            SDF_NODE_LOOP(kk) {
                SDF_LOAD_STATE(kk);
                double du[NU], gu[NU], Dp[6] = {0, 0, 0, 0, 0, 0}, gp[3] = {0, 0, 0};
#pragma unroll
                for (int j = 0; j < NU; ++j) {
                    double thn = 0.0, cn = 0.0;
                    if (kk < N) {
                        const double wu = Wsm[j * KP + kk];
#pragma unroll
                        for (int side = 0; side < 2; ++side) {
                            const int a = side * NU + j;
                            const double sgn = side == 0 ? 1.0 : -1.0;
                            const double cr = pass == 0 ? 0.0 : tau - prod[a * KP + kk];
                            const HardEval ev = hard_eval(tb[a], lb[a], sgn * wu, hard_d(kk, a), cr);
                            thn += ev.Th;
                            cn += sgn * ev.coef;
                            if (pass == 0) {
                                res_rows = fmax(res_rows, fabs(ev.r_p));
                                const double c = lb[a] * tb[a];
                                cmax = fmax(cmax, c); csum += c;
                            }
                        }
                    }
                    du[j] = thn; gu[j] = cn;
                }
                if (kk >= k0) {
#pragma unroll
                    for (int r = 0; r < NH; ++r) {
                        const double c0 = c_pos(kk, r, 0), c1 = c_pos(kk, r, 1), c2 = c_pos(kk, r, 2);
                        const double cw = c0 * Wsm[(HC0 + 0) * KP + kk] + c1 * Wsm[(HC0 + 1) * KP + kk] + c2 * Wsm[(HC0 + 2) * KP + kk];
                        const double z = P.zs[r] * scale(kk), Z = P.Zs[r] * scale(kk);
                        double thn = 0.0, cn = 0.0;
#pragma unroll
                        for (int side = 0; side < 2; ++side) {
                            const int a = side * NH + r;
                            const double sgn = side == 0 ? 1.0 : -1.0;
                            const double ct = pass == 0 ? 0.0 : tau - prod[(2 * NU + a) * KP + kk];
                            const double cs = pass == 0 ? 0.0 : tau - prod[(2 * NU + 2 * NH + a) * KP + kk];
                            const SoftEval ev = soft_eval(ts[a], ls[a], ss[a], ms[a], sgn * cw, soft_d(kk, a), z, Z, ct, cs);
                            thn += ev.Tht;
                            cn += sgn * ev.coef;
                            if (pass == 0) {
                                res_rows = fmax(res_rows, fmax(fabs(ev.r_p), fabs(ev.r_s)));
                                const double ca = ls[a] * ts[a], cb = ms[a] * ss[a];
                                cmax = fmax(cmax, fmax(ca, cb)); csum += ca + cb;
                            }
                        }
                        Dp[0] += thn * c0 * c0; Dp[1] += thn * c1 * c0; Dp[2] += thn * c1 * c1;
                        Dp[3] += thn * c2 * c0; Dp[4] += thn * c2 * c1; Dp[5] += thn * c2 * c2;
                        gp[0] += cn * c0; gp[1] += cn * c1; gp[2] += cn * c2;
                    }
                }
#pragma unroll
                for (int j = 0; j < NU; ++j) { corr[j * KP + kk] = du[j]; corr[(NU + 6 + j) * KP + kk] = gu[j]; }
#pragma unroll
                for (int c = 0; c < 6; ++c) corr[(NU + c) * KP + kk] = Dp[c];
#pragma unroll
                for (int c = 0; c < 3; ++c) corr[(2 * NU + 6 + c) * KP + kk] = gp[c];
            }
            if (pass == 0) {
                // ---- convergence test on the state BEFORE this iteration's step ----
                res_rows = L.max(res_rows);
                cmax = L.max(cmax);
                mubar = L.sum(csum) / ncomp;
                res_track = fmax(res_rows, phi * res_v0);
#if defined(SDFNMPC_QP_DEBUG) && !defined(__CUDA_ARCH__)
                printf("it %d res %.3e (rows %.3e) cmax %.3e phi %.3e\n", it, res_track, res_rows, cmax, phi);
#endif
                if (!(res_track == res_track) || !(mubar == mubar) || res_track > 1e300) { status = 4; stop = true; break; }
                if (res_track <= P.tol_res && cmax <= P.tol) { status = 0; stop = true; break; }
                if (it >= P.max_iter) { status = 2; stop = true; break; }
                // Mehrotra's corrector can cycle on poorly centred iterates; if the largest complementarity product
                // has not dropped by 30 % over two iterations switch (for good) to plain damped path following.
                if (it >= 2 && cmax > 0.7 * cm_prev2) safe = true;
                cm_prev2 = cm_prev1; cm_prev1 = cmax;
            }
            L.sync();
            // =========== backward Riccati sweep (pass 0: factorise + solve; pass 1: new right-hand side only) ===========
            if (pass == 0) {
                SDF_LANES_FOR(e, NX * NX) {
                    const int i = e / NX, j = e - i * NX;
                    double s = H[(size_t)N * NW * NW + (size_t)(NU + i) * NW + NU + j];
                    const int pi_ = i + NU - HC0, pj = j + NU - HC0;
                    if (pi_ >= 0 && pi_ < 3 && pj >= 0 && pj < 3) s += corr[(NU + sym3(pi_, pj)) * KP + N];
                    Psm[e] = s;
                }
            }
            SDF_LANES_FOR(e, NX) {
                double s = phi * rv0[(size_t)N * NW + NU + e];
                const int pe = e + NU - HC0;
                if (pe >= 0 && pe < 3) s += corr[(2 * NU + 6 + pe) * KP + N];
                pv[e] = s;
            }
            L.sync();
            // stage data of node kk is prefetched into the other half of `stg` while node kk+1 is processed
            auto prefetch = [&](int kk, double* dst) {
                stage_async(L, dst + S::ST_AB, P.AB + (size_t)kk * NX * NW, NX * NW);
                stage_async(L, dst + S::ST_RV, rv0 + (size_t)kk * NW, NW);
                if (pass == 0) {
                    stage_async(L, dst + S::ST_H, H + (size_t)kk * NW * NW, NW * NW);
                    stage_async(L, dst + S::ST_RD, rd0 + (size_t)(kk + 1) * NX, NX);
                } else {
                    stage_async(L, dst + S::ST_PB, PbM + (size_t)kk * NX, NX);
                    stage_async(L, dst + S::ST_L, Lg + (size_t)kk * NU * NU, NU * NU);
                    stage_async(L, dst + S::ST_K, Kg + (size_t)kk * NU * NX, NU * NX);
                }
                stage_commit();
            };
            prefetch(N - 1, stg);
            for (int kk = N - 1; kk >= 0; --kk) {
                const double* sb = stg + ((N - 1 - kk) & 1) * S::STAGE_DOUBLES;
                stage_wait_all();
                L.sync();
                if (kk > 0) prefetch(kk - 1, stg + ((N - kk) & 1) * S::STAGE_DOUBLES);
                const double* ABk = sb + S::ST_AB;
                if (pass == 0) {
                    SDF_LANES_FOR(e, NX * NW) { const int i = e / NW; ABb[i * NWB + (e - i * NW)] = ABk[e]; }
                    SDF_LANES_FOR(e, NX) ABb[e * NWB + NW] = phi * sb[S::ST_RD + e];
                    L.sync();
                    SDF_LANES_FOR(e, NX * NWB) {
                        const int i = e / NWB, j = e - i * NWB;
                        double s = 0.0;
#pragma unroll
                        for (int l = 0; l < NX; ++l) s += Psm[i * NX + l] * ABb[l * NWB + j];
                        if (j == NW) { PbM[kk * NX + i] = s; s += pv[i]; }
                        PAB[e] = s;
                    }
                    L.sync();
                    SDF_LANES_FOR(e, TRI + NW) {
                        if (e < TRI) {
                            const int i = tri.i[e], j = tri.j[e];
                            double s = sb[S::ST_H + i * NW + j];
#pragma unroll
                            for (int l = 0; l < NX; ++l) s += ABb[l * NWB + i] * PAB[l * NWB + j];
                            if (i == j && i < NU) s += corr[i * KP + kk];
                            const int pi_ = i - HC0, pj = j - HC0;
                            if (pi_ >= 0 && pi_ < 3 && pj >= 0 && pj < 3) s += corr[(NU + sym3(pi_, pj)) * KP + kk];
                            Msm[i * NW + j] = s;
                            Msm[j * NW + i] = s;
                        } else {
                            const int i = e - TRI;
                            double s = phi * sb[S::ST_RV + i];
#pragma unroll
                            for (int l = 0; l < NX; ++l) s += ABb[l * NWB + i] * PAB[l * NWB + NW];
                            if (i < NU) s += corr[(NU + 6 + i) * KP + kk];
                            const int pi_ = i - HC0;
                            if (pi_ >= 0 && pi_ < 3) s += corr[(2 * NU + 6 + pi_) * KP + kk];
                            rv[i] = s;
                        }
                    }
                    L.sync();
                    // Cholesky of Ruu (every lane, redundantly, in registers); diagonal kept inverted
                    double Lr[NU][NU];
                    bool ok = true;
#pragma unroll
                    for (int i = 0; i < NU; ++i)
#pragma unroll
                        for (int j = 0; j <= i; ++j) {
                            double s = Msm[i * NW + j];
#pragma unroll
                            for (int l = 0; l < j; ++l) s -= Lr[i][l] * Lr[j][l];
                            if (i == j) { ok = ok && (s > 0.0); Lr[i][i] = 1.0 / sqrt(s); }
                            else Lr[i][j] = s * Lr[j][j];
                        }
                    if (!ok) status = 4;
                    SDF_LANES_FOR(c, NX + 1) {
                        double x[NU];
#pragma unroll
                        for (int i = 0; i < NU; ++i) {
                            double s = c < NX ? Msm[i * NW + NU + c] : rv[i];
#pragma unroll
                            for (int l = 0; l < i; ++l) s -= Lr[i][l] * x[l];
                            x[i] = s * Lr[i][i];
                        }
#pragma unroll
                        for (int i = NU - 1; i >= 0; --i) {
                            double s = x[i];
#pragma unroll
                            for (int l = i + 1; l < NU; ++l) s -= Lr[l][i] * x[l];
                            x[i] = s * Lr[i][i];
                        }
#pragma unroll
                        for (int i = 0; i < NU; ++i) {
                            if (c < NX) { Ksm[i * NX + c] = -x[i]; Kg[(size_t)kk * NU * NX + i * NX + c] = -x[i]; }
                            else kfs[i] = -x[i];
                        }
                    }
                    if (lane == 0)
                        for (int i = 0; i < NU; ++i)
                            for (int j = 0; j < NU; ++j) Lg[(size_t)kk * NU * NU + i * NU + j] = j <= i ? Lr[i][j] : 0.0;
                    L.sync();
                    // P <- Qxx + Rux' K (lower triangle, mirrored: stays exactly symmetric), p <- r_x + Rux' kff
                    constexpr int TRX = NX * (NX + 1) / 2;
                    SDF_LANES_FOR(e, TRX + NX + NU) {
                        if (e < TRX) {
                            const int i = tri.i[e], j = tri.j[e];      // same row-major lower-triangle enumeration
                            double s = Msm[(NU + i) * NW + NU + j];
#pragma unroll
                            for (int u = 0; u < NU; ++u) s += Msm[(NU + i) * NW + u] * Ksm[u * NX + j];
                            Psm[i * NX + j] = s;
                            Psm[j * NX + i] = s;
                        } else if (e < TRX + NX) {
                            const int i = e - TRX;
                            double s = rv[NU + i];
#pragma unroll
                            for (int u = 0; u < NU; ++u) s += Msm[(NU + i) * NW + u] * kfs[u];
                            pv[i] = s;
                        } else {
                            const int j = e - TRX - NX;
                            dW[j * KP + kk] = kfs[j];        // feed-forward, consumed by the roll-out
                        }
                    }
                    L.sync();
                } else {
                    SDF_LANES_FOR(e, NX) vv[e] = sb[S::ST_PB + e] + pv[e];   // PbM already carries this iteration's phi
                    L.sync();
                    SDF_LANES_FOR(i, NW) {
                        double s = phi * sb[S::ST_RV + i];
#pragma unroll
                        for (int l = 0; l < NX; ++l) s += ABk[l * NW + i] * vv[l];
                        if (i < NU) s += corr[(NU + 6 + i) * KP + kk];
                        const int pi_ = i - HC0;
                        if (pi_ >= 0 && pi_ < 3) s += corr[(2 * NU + 6 + pi_) * KP + kk];
                        rv[i] = s;
                    }
                    L.sync();
                    const double* Lk = sb + S::ST_L;
                    double x[NU];
#pragma unroll
                    for (int i = 0; i < NU; ++i) {
                        double s = rv[i];
#pragma unroll
                        for (int l = 0; l < i; ++l) s -= Lk[i * NU + l] * x[l];
                        x[i] = s * Lk[i * NU + i];
                    }
#pragma unroll
                    for (int i = NU - 1; i >= 0; --i) {
                        double s = x[i];
#pragma unroll
                        for (int l = i + 1; l < NU; ++l) s -= Lk[l * NU + i] * x[l];
                        x[i] = s * Lk[i * NU + i];
                    }
                    const double* Kk = sb + S::ST_K;
                    SDF_LANES_FOR(i, NX + NU) {
                        if (i < NX) {
                            double s = rv[NU + i];       // pv was last read two barriers ago
#pragma unroll
                            for (int u = 0; u < NU; ++u) s += Kk[u * NX + i] * rv[u];
                            pv[i] = s;
                        } else {
                            dW[(i - NX) * KP + kk] = -x[i - NX];
                        }
                    }
                    L.sync();
                }
            }
            // ---------- forward roll-out: dW = step ----------
            SDF_LANES_FOR(i, NX) dW[(NU + i) * KP] = 0.0;      // dx_0 = 0 (w_x0 already equals x0 - xbar_0)
            L.sync();
            auto prefetch_f = [&](int kk, double* dst) {
                stage_async(L, dst + S::ST_AB, P.AB + (size_t)kk * NX * NW, NX * NW);
                stage_async(L, dst + S::ST_RD, rd0 + (size_t)(kk + 1) * NX, NX);
                stage_async(L, dst + S::ST_K, Kg + (size_t)kk * NU * NX, NU * NX);
                stage_commit();
            };
            prefetch_f(0, stg);
            for (int kk = 0; kk < N; ++kk) {
                const double* sb = stg + (kk & 1) * S::STAGE_DOUBLES;
                stage_wait_all();
                L.sync();
                if (kk + 1 < N) prefetch_f(kk + 1, stg + ((kk + 1) & 1) * S::STAGE_DOUBLES);
                const double* Kk = sb + S::ST_K;
                SDF_LANES_FOR(j, NU) {
                    double s = dW[j * KP + kk];
#pragma unroll
                    for (int i = 0; i < NX; ++i) s += Kk[j * NX + i] * dW[(NU + i) * KP + kk];
                    dW[j * KP + kk] = s;
                }
                L.sync();
                const double* ABk = sb + S::ST_AB;
                SDF_LANES_FOR(i, NX) {
                    double s = phi * sb[S::ST_RD + i];
#pragma unroll
                    for (int j = 0; j < NW; ++j) s += ABk[i * NW + j] * dW[j * KP + kk];
                    dW[(NU + i) * KP + kk + 1] = s;
                }
                L.sync();
            }
            // =========== row steps (lane = node): the node's (dt, dlam, ds, dmu) stay in registers ===========
            double amax = 1e30;
            SDF_NODE_LOOP(kk) {
                SDF_LOAD_STATE(kk);
#pragma unroll
                for (int a = 0; a < 2 * NU; ++a) { dtb[a] = 0.0; dlb[a] = 0.0; }
#pragma unroll
                for (int a = 0; a < 2 * NH; ++a) { dts[a] = 0.0; dls[a] = 0.0; dss[a] = 0.0; dms[a] = 0.0; }
                if (kk < N) {
#pragma unroll
                    for (int j = 0; j < NU; ++j) {
                        const double wu = Wsm[j * KP + kk], dwu = dW[j * KP + kk];
#pragma unroll
                        for (int side = 0; side < 2; ++side) {
                            const int a = side * NU + j;
                            const double sgn = side == 0 ? 1.0 : -1.0;
                            const double cr = pass == 0 ? 0.0 : tau - prod[a * KP + kk];
                            const HardEval ev = hard_eval(tb[a], lb[a], sgn * wu, hard_d(kk, a), cr);
                            const double kap = cr - lb[a] * tb[a];
                            dtb[a] = sgn * dwu - ev.r_p;
                            dlb[a] = (kap - lb[a] * dtb[a]) / tb[a];
                            amax = fmin(amax, fmin(step_to_boundary(tb[a], dtb[a]), step_to_boundary(lb[a], dlb[a])));
                        }
                    }
                }
                if (kk >= k0) {
#pragma unroll
                    for (int r = 0; r < NH; ++r) {
                        const double c0 = c_pos(kk, r, 0), c1 = c_pos(kk, r, 1), c2 = c_pos(kk, r, 2);
                        const double cw = c0 * Wsm[(HC0 + 0) * KP + kk] + c1 * Wsm[(HC0 + 1) * KP + kk] + c2 * Wsm[(HC0 + 2) * KP + kk];
                        const double cd = c0 * dW[(HC0 + 0) * KP + kk] + c1 * dW[(HC0 + 1) * KP + kk] + c2 * dW[(HC0 + 2) * KP + kk];
                        const double z = P.zs[r] * scale(kk), Z = P.Zs[r] * scale(kk);
#pragma unroll
                        for (int side = 0; side < 2; ++side) {
                            const int a = side * NH + r;
                            const double sgn = side == 0 ? 1.0 : -1.0;
                            const double ct = pass == 0 ? 0.0 : tau - prod[(2 * NU + a) * KP + kk];
                            const double cs = pass == 0 ? 0.0 : tau - prod[(2 * NU + 2 * NH + a) * KP + kk];
                            const SoftEval ev = soft_eval(ts[a], ls[a], ss[a], ms[a], sgn * cw, soft_d(kk, a), z, Z, ct, cs);
                            const double kap = ct - ls[a] * ts[a], nu = cs - ms[a] * ss[a];
                            const double gd = sgn * cd;
                            dss[a] = (ev.rho - ev.Th * gd) / ev.D;
                            dts[a] = gd + dss[a] - ev.r_p;
                            dls[a] = (kap - ls[a] * dts[a]) / ts[a];
                            dms[a] = (nu - ms[a] * dss[a]) / ss[a];
                            amax = fmin(amax, fmin(fmin(step_to_boundary(ts[a], dts[a]), step_to_boundary(ls[a], dls[a])),
                                                   fmin(step_to_boundary(ss[a], dss[a]), step_to_boundary(ms[a], dms[a]))));
                        }
                    }
                }
                SDF_STORE_DELTA(kk);
            }
            amax = L.min(amax);
            if (pass == 0) {
                // affine step length -> centring parameter; second-order products for the corrector
                const double aa = fmin(1.0, amax);
                const double so = safe ? 0.0 : 1.0;
                double csum_aff = 0.0;
                SDF_NODE_LOOP(kk) {
                    SDF_LOAD_STATE(kk);
                    SDF_LOAD_DELTA(kk);
                    if (kk < N) {
#pragma unroll
                        for (int a = 0; a < 2 * NU; ++a) {
                            csum_aff += (lb[a] + aa * dlb[a]) * (tb[a] + aa * dtb[a]);
                            prod[a * KP + kk] = so * dlb[a] * dtb[a];
                        }
                    }
                    if (kk >= k0) {
#pragma unroll
                        for (int a = 0; a < 2 * NH; ++a) {
                            csum_aff += (ls[a] + aa * dls[a]) * (ts[a] + aa * dts[a]) + (ms[a] + aa * dms[a]) * (ss[a] + aa * dss[a]);
                            prod[(2 * NU + a) * KP + kk] = so * dls[a] * dts[a];
                            prod[(2 * NU + 2 * NH + a) * KP + kk] = so * dms[a] * dss[a];
                        }
                    }
                }
                const double mu_aff = L.sum(csum_aff) / ncomp;
                double sig = mu_aff / fmax(mubar, 1e-300);
                sig = fmin(fmax(sig * sig * sig, safe ? 0.2 : 0.0), 1.0);
                tau = fmax(sig * mubar, 0.1 * P.tol);
                L.sync();
            } else {
                // =========== damped step ===========
                const double al = fmin(1.0, 0.995 * amax);
                SDF_NODE_LOOP(kk) {
                    SDF_LOAD_STATE(kk);
                    SDF_LOAD_DELTA(kk);
#pragma unroll
                    for (int a = 0; a < 2 * NU; ++a) { tb[a] += al * dtb[a]; lb[a] += al * dlb[a]; }
#pragma unroll
                    for (int a = 0; a < 2 * NH; ++a) { ts[a] += al * dts[a]; ls[a] += al * dls[a]; ss[a] += al * dss[a]; ms[a] += al * dms[a]; }
                    SDF_STORE_STATE(kk);
                }
                for (int e = lane; e < NW * KP; e += NL) Wsm[e] += al * dW[e];
                phi *= (1.0 - al);
                L.sync();
            }
        }
        if (stop || status == 4) break;
    }

    // ---- verification of the KKT residuals the iteration only tracked (costates by one backward sweep) ----
    double res_true = 0.0;
    {
        SDF_NODE_LOOP(kk) {
            SDF_LOAD_STATE(kk);
            for (int j = 0; j < NU; ++j) corr[j * KP + kk] = kk < N ? lb[j] - lb[NU + j] : 0.0;
            for (int r = 0; r < NH; ++r) corr[(NU + r) * KP + kk] = kk >= k0 ? ls[r] - ls[NH + r] : 0.0;
        }
        L.sync();
        for (int kk = N; kk >= 0; --kk) {
            for (int i = lane; i < NW; i += NL) {
                if (kk == N && i < NU) continue;
                double s = g[(size_t)kk * NW + i];
                for (int j = 0; j < NW; ++j) s += H[(size_t)kk * NW * NW + i * NW + j] * Wsm[j * KP + kk];
                if (i < NU) s -= corr[i * KP + kk];
                if (i >= HC0 && i < HC0 + 3)
                    for (int r = 0; r < NH; ++r) s -= corr[(NU + r) * KP + kk] * c_pos(kk, r, i - HC0);
                if (kk < N)
                    for (int l = 0; l < NX; ++l) s += P.AB[(size_t)kk * NX * NW + l * NW + i] * pig[(size_t)(kk + 1) * NX + l];
                if (i < NU) res_true = fmax(res_true, fabs(s));      // input stationarity
                else pig[(size_t)kk * NX + (i - NU)] = s;            // defines the costate of node kk (x-stationarity == 0)
            }
            if (kk < N)
                for (int i = lane; i < NX; i += NL) {
                    double s = P.bdyn[kk * NX + i] - Wsm[(NU + i) * KP + kk + 1];
                    for (int j = 0; j < NW; ++j) s += P.AB[(size_t)kk * NX * NW + i * NW + j] * Wsm[j * KP + kk];
                    res_true = fmax(res_true, fabs(s));
                }
            L.sync();
        }
        res_true = L.max(res_true);
        if (status == 0 && !(res_true <= 100.0 * P.tol_res)) status = 3;    // tracked residual was wrong: report, do not hide
    }

    double step_inf = 0.0;
    for (int e = lane; e < (N + 1) * NW; e += NL) { const int kk = e / NW, i = e - kk * NW; step_inf = fmax(step_inf, fabs(Wsm[i * KP + kk])); }
    step_inf = L.max(step_inf);
    // ---- full SQP step and outputs -----------------------------------------------------------
    if (status != 4) {
        for (int e = lane; e < (N + 1) * NX; e += NL) { const int kk = e / NX, i = e - kk * NX; P.xbar[e] += Wsm[(NU + i) * KP + kk]; }
        for (int e = lane; e < N * NU; e += NL) { const int kk = e / NU, j = e - kk * NU; P.ubar[e] += Wsm[j * KP + kk]; }
    }
    L.sync();
    for (int j = lane; j < NU; j += NL) P.u0[j] = P.ubar[j];
    SDF_NODE_LOOP(kk) {
        SDF_LOAD_STATE(kk);
        if (P.lam_hard && kk < N)
            for (int a = 0; a < 2 * NU; ++a) P.lam_hard[kk * 2 * NU + a] = lb[a];
        if (P.lam_soft)
            for (int a = 0; a < 2 * NH; ++a) P.lam_soft[kk * 2 * NH + a] = kk >= k0 ? ls[a] : 0.0;
    }
#if !defined(__CUDA_ARCH__)
    delete[] TB;
    delete[] TD;
#endif
    if (lane == 0) {
        *P.status = status;
        *P.iters = it;
        if (P.res_out) { P.res_out[0] = res_true; P.res_out[1] = step_inf; }
    }
}

#undef SDF_NODE_LOOP
#undef SDF_LANES_FOR
#undef SDF_LOAD_STATE
#undef SDF_STORE_STATE
#undef SDF_LOAD_DELTA
#undef SDF_STORE_DELTA

}  // namespace sdfnmpc
