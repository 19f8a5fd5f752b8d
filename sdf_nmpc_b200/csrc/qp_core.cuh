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
// Work mapping (v3).  Row phases: lane k owns node k and holds the node's (t, lam, s, mu) and
// their steps in registers INSIDE a row phase; between the phases (during the Riccati sweeps)
// they are parked in an L2-resident [item][node] block, so the sweeps have the register file
// to themselves.  Matrix phases (Riccati): lanes split the entries of the small dense products;
// the per-stage operands ([A B], H, gains) stream global -> shared through a two-slot ring of
// cp.async copies issued one stage ahead (the prologue and the final verification use the same ring).
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
    // Per-stage data travels global -> shared by bulk copies, so what one sweep needs of a stage is contiguous in the
    // global scratch:  record A_k = [ rv0_k | rd0_{k+1} | H_k ]  (written once per solve),
    //                  record F_k = [ PbM_k | L_k | K_k ]         (written by the factorising sweep of every iteration).
    static constexpr int RA_RV = 0, RA_RD = NW, RA_H = NW + NX, REC_A = NW + NX + NW * NW;
    static constexpr int RF_PB = 0, RF_L = NX, RF_K = NX + NU * NU, REC_F = NX + NU * NU + NU * NX;
    // shared stage slot: AB | rv0 | rd0 | { H (factorising sweep, verification)  or  record F (the other sweeps) }
    static constexpr int ST_AB = 0, ST_A = NX * NW, ST_RV = ST_A + RA_RV, ST_RD = ST_A + RA_RD, ST_H = ST_A + RA_H,
                         ST_F = ST_H, ST_PB = ST_F + RF_PB, ST_L = ST_F + RF_L, ST_K = ST_F + RF_K,
                         STAGE_DOUBLES = ST_A + REC_A;
    static_assert(REC_F <= NW * NW, "record F overlays H in a stage slot");
    static_assert(NX % 2 == 0 && NW % 2 == 0 && NU % 2 == 0 && REC_A % 2 == 0 && REC_F % 2 == 0 && STAGE_DOUBLES % 2 == 0,
                  "bulk copies move 16-byte units");
    // doubles of per-instance global scratch
    SDF_HD static size_t work_doubles(int N) {
        size_t n1 = (size_t)(N + 1);
        return n1 * REC_A                  // rv0, rd0, H per node
               + n1 * NW                   // g
               + (size_t)N * REC_F         // PbM, Linv, K per stage
               + ((n1 * NH * 3 + 1) & ~(size_t)1)    // position blocks of the constraint rows (even: instances stay 16-byte aligned)
               + 2 * (size_t)(4 * NU + 8 * NH) * kp(N);   // parked row state and row steps, [item][node]
    }
    // doubles of per-lane-group shared memory
    SDF_HD static int smem_doubles(int N) {
        return NX * NX + NX * (NW + 1) + NW * NW + NW + NU * NX + NU + 2 * NX       // Riccati workspace
               + 2 * NW * kp(N)                                                   // W (iterate), dW (step)
               + (NCORR + NPROD) * kp(N) + (((NCORR + NPROD) * kp(N)) & 1) + 8     // stage slots and the next lane group stay 16-byte aligned (odd N)
               + 2 * STAGE_DOUBLES + 2;                                           // double-buffered per-stage data, 2 mbarriers
    }
};

template <int NX, int NU, int NY, int NH>
struct QpProb {            // all pointers are offset to ONE instance
    int N, k0, max_iter, nyN;
    // Lifted terminal rows (flags.recursive_feasibility / stability, term_lift.cuh): nodes n_real+1..N are auxiliary nodes behind the
    // OCP's terminal node n_real (no cost, no regularisation of their state, rows with per-node bounds); n_real == N: plain OCP.
    // Scalar solver only (qp_mma.cuh ignores these fields).
    int n_real;
    const double *nlh, *nuh, *nzs, *nZs;   // [N+1][NH] per-node bounds / slack weights of the h rows, or null (lh / uh / zs / Zs apply)
    double lm, tol, tol_res, max_df, flag;
    double sdf_cost_w;     // weight of the optional stage cost row (1 - 0.5 sdf)^4 (flags.sdf_cost, gen_model.py:65-66); 0 = off
    const double* AB;      // [N][NX][NW]
    const double* ABt;     // [N][16][12] the same in the tensor-core operand layout (qp_mma.cuh); unused by the scalar solver
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
    int debug;             // development (-DSDFNMPC_QP_DEBUG_INST=b): per-iteration printf of instance b
    long long* trace;      // development (-DSDFNMPC_QP_TRACE): cycles per phase of IPM iteration 2, [pass][16]; null = off
};

// Initial point of the interior-point iteration.  tools/qp_init_sweep.py: smaller floors / mu0 or an 'inactive-row' split of
// the soft multipliers change the mean iteration count by less than +-0.3 (9.7 cold), so the plain choice stays.
#ifndef SDFNMPC_QP_T0
#define SDFNMPC_QP_T0 0.5          // floor of the initial constraint slacks t
#endif
#ifndef SDFNMPC_QP_MU0
#define SDFNMPC_QP_MU0 1.0         // initial complementarity lam * t of rows that start interior
#endif
#ifndef SDFNMPC_QP_LCAP
#define SDFNMPC_QP_LCAP 5.0        // largest initial multiplier of a soft row (the rest of z + Z s0 starts on the slack's multiplier)
#endif
#ifndef SDFNMPC_QP_LCAP_LIFT
#define SDFNMPC_QP_LCAP_LIFT 10.0  // the same cap for the lifted OCP (exact-penalty rows start thousands above their multipliers): 10 measured best
#endif
#ifndef SDFNMPC_QP_LCAP_ALL
#define SDFNMPC_QP_LCAP_ALL 1      // 0: only the lifted OCP's rows are capped (the plain OCP then starts at the even split)
#endif
#ifndef SDFNMPC_QP_S0
#define SDFNMPC_QP_S0 0.5          // initial soft-constraint slack s
#endif

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

// Reciprocal of a positive, normal double.  On the device: MUFU seed (rcp.approx.ftz.f64, ~2^-23) + two Newton steps = 5
// instructions and <= 1 ulp, instead of the ~20-instruction IEEE division with its slow-path call - the kernel has ~100
// inlined call sites (30 KB of a 171 KB kernel that already misses the instruction cache).
SDF_HD double rcp(double x) {
#if defined(__CUDA_ARCH__) && !defined(SDFNMPC_QP_EXACT_RCP)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
#else
    return 1.0 / x;
#endif
}
SDF_HD double rsqrt_d(double x) {
#if defined(__CUDA_ARCH__) && !defined(SDFNMPC_QP_EXACT_RCP)
    // MUFU seed + two Newton steps y += y/2 (1 - x y^2): 9 instructions instead of 17 (positive, normal pivots only)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-(x * y), y, 1.0);
    y = fma(0.5 * y, e, y);
    e = fma(-(x * y), y, 1.0);
    return fma(0.5 * y, e, y);
#elif defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}
// Step to the boundary x + alpha dx >= 0 without a division per row: rows report the ratio -dx/x (through the
// reciprocal of x they already hold); alpha_max = 1 / max ratio is formed once per lane.
SDF_HD double boundary_ratio(double dx, double rx) { return -dx * rx; }
// two consecutive doubles of a 16-byte aligned shared-memory row in one LDS.128
struct Pair { double x, y; };
SDF_HD Pair ld2(const double* p) {
#if defined(__CUDA_ARCH__)
    const double2 v = *reinterpret_cast<const double2*>(p);
    return Pair{v.x, v.y};
#else
    return Pair{p[0], p[1]};
#endif
}
SDF_HD int sym3(int a, int b) { return a >= b ? a * (a + 1) / 2 + b : b * (b + 1) / 2 + a; }

// Everything the elimination needs from one row; evaluated by ONE function so that the two places that use
// r_p / rho (right-hand side and step) see bit-identical values.
struct HardEval { double r_p, Th, coef, rt, rl; };                       // rt = 1/t, rl = 1/lam
struct SoftEval { double r_p, r_s, Th, rD, Tht, rho, coef, rt, rl, rs, rm; };   // rD = 1/D, rs = 1/s, rm = 1/mu

#if defined(__CUDA_ARCH__)
#define SDF_NOINLINE __device__ __noinline__
#else
#define SDF_NOINLINE inline
#endif

// Inlined: the independent rows of a node interleave their reciprocal / FMA chains (measured 5 % faster than calls;
// -DSDFNMPC_QP_NOINLINE_EVAL restores the out-of-line form for A/B runs, tools/qp_ab.py).
#if defined(SDFNMPC_QP_NOINLINE_EVAL)
#define SDF_HARD_LINKAGE SDF_NOINLINE
#define SDF_SOFT_LINKAGE SDF_NOINLINE
#else
#define SDF_HARD_LINKAGE SDF_HD
#define SDF_SOFT_LINKAGE SDF_HD
#endif

SDF_HARD_LINKAGE HardEval hard_eval(double t, double lam, double gw, double d, double corr) {
    HardEval r;
    const double c = lam * t, rc = rcp(c);        // one reciprocal serves both 1/t and 1/lam
    const double rt = lam * rc;
    r.rt = rt; r.rl = t * rc;
    r.r_p = t - (gw - d);
    r.Th = lam * rt;
    const double kap = corr - c;                  // corr = 0 (affine) or tau - dlam_aff*dt_aff
    r.coef = -kap * rt - r.Th * r.r_p;
    return r;
}

SDF_SOFT_LINKAGE SoftEval soft_eval(double t, double lam, double s, double mu, double gw, double d, double z, double Z,
                                double corr_t, double corr_s) {
    SoftEval r;
    const double c1 = lam * t, c2 = mu * s;
    const double rc1 = rcp(c1), rc2 = rcp(c2);
    const double rt = lam * rc1, rs = mu * rc2;
    r.rt = rt; r.rl = t * rc1; r.rs = rs; r.rm = s * rc2;
    r.r_p = t - (gw - d + s);
    r.r_s = z + Z * s - lam - mu;
    r.Th = lam * rt;
    const double zs = Z + mu * rs;
    const double rD = rcp(zs + r.Th);
    r.rD = rD;
    r.Tht = r.Th * zs * rD;
    const double kap = corr_t - c1, nu = corr_s - c2;
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
    const int NR = P.n_real;                      // terminal node of the OCP (== N unless terminal rows are lifted)
    const int NL = Lanes::count();
    const int lane = L.lane;
    const int ncomp = N * 2 * NU + 2 * (N + 1 - k0) * 2 * NH;

    // ---- global scratch ---------------------------------------------------------------------
    double* p = P.work;
    auto take = [&](size_t n) { double* r = p; p += n; return r; };
    const size_t n1 = (size_t)(N + 1);
    double* recA = take(n1 * S::REC_A);        // per node: rv0 | rd0 (of the NEXT node) | H
    double* g = take(n1 * NW);
    double* recF = take((size_t)N * S::REC_F);  // per stage: PbM | Linv | K
    double* cpos = take(n1 * NH * 3);          // per node: position block [NH][3] of the constraint rows
    // rv0: stationarity residual at the initial point (scaled by phi afterwards); rd0[k] = AB_{k-1} w_{k-1} + b_{k-1} - w_x,k
    // lives in record k-1 (the stage that produces node k); H: Gauss-Newton Hessian; L: Cholesky factor of Ruu, diagonal INVERTED
    auto Hk = [&](int k) { return recA + (size_t)k * S::REC_A + S::RA_H; };
    auto rv0k = [&](int k) { return recA + (size_t)k * S::REC_A + S::RA_RV; };
    auto rd0k = [&](int k) { return recA + (size_t)(k - 1) * S::REC_A + S::RA_RD; };   // k >= 1
    auto PbMk = [&](int k) { return recF + (size_t)k * S::REC_F + S::RF_PB; };
    auto Lgk = [&](int k) { return recF + (size_t)k * S::REC_F + S::RF_L; };
    auto Kgk = [&](int k) { return recF + (size_t)k * S::REC_F + S::RF_K; };

    // ---- shared workspace -----------------------------------------------------------------------
    double* Psm = sm;                    // [NX][NX]
    double* PAB = Psm + NX * NX;         // [NX][NW+1]  P [A B | phi rd]
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
    double* stg = prod + NPROD * KP + (((NCORR + NPROD) * KP) & 1);   // 2 x STAGE_DOUBLES (16-byte aligned: everything before `corr` is even)
    StagePipe pipe;
    pipe.init(L, reinterpret_cast<unsigned long long*>(stg + 2 * S::STAGE_DOUBLES));

    auto scale = [&](int k) { return k < N ? P.dt[k] : 1.0; };
    auto wrow = [&](int k, int r) { return k < NR ? P.W[k * NY + r] : ((k == NR && r < P.nyN) ? P.WN[r] : 0.0); };
    auto zrow = [&](int k, int r) { return (P.nzs ? P.nzs[k * NH + r] : P.zs[r]) * scale(k); };
    auto Zrow = [&](int k, int r) { return (P.nZs ? P.nZs[k * NH + r] : P.Zs[r]) * scale(k); };

    // ---- prologue: sdf row of C / h, Gauss-Newton Hessian and gradient ----------------------
    const bool sdf_cost = P.sdf_cost_w > 0.0;          // the cost row also needs the SDF at node 0
    for (int k = lane; k <= N; k += NL) {
        if (k >= (sdf_cost ? 0 : k0) && k <= NR) {
            const double* R = P.pose + 3;
            double gx = P.sdf_g[k * 3 + 0], gy = P.sdf_g[k * 3 + 1], gz = P.sdf_g[k * 3 + 2];
            double* Cr = P.C + ((size_t)k * NH + (NH - 1)) * NW;
            for (int jj = 0; jj < 3; ++jj) Cr[HC0 + jj] = P.flag * (R[jj * 3 + 0] * gx + R[jj * 3 + 1] * gy + R[jj * 3 + 2] * gz);
            P.hval[k * NH + (NH - 1)] = P.flag * (double)P.sdf_d[k] + (1.0 - P.flag) * P.max_df;
        }
        // compact copy of the rows' position blocks: the row phases read only these 3 of a row's NW entries, and through the
        // full-width C they touched every cache line of it (7 KB per instance, 8 instances per SM against 60 KB of L1)
        for (int r = 0; r < NH; ++r)
            for (int c = 0; c < 3; ++c) cpos[(k * NH + r) * 3 + c] = P.C[((size_t)k * NH + r) * NW + HC0 + c];
    }
    if (sdf_cost) L.sync();
    // optional cost row r = (1 - 0.5 d)^4 on the stage nodes: dr/dp = -2 (1 - 0.5 d)^3 dd/dp, only the position block
    auto cost_r = [&](int k, double& r, double* Jr) {
        const double q = 1.0 - 0.5 * P.hval[k * NH + (NH - 1)];
        const double* Cr = P.C + ((size_t)k * NH + (NH - 1)) * NW;
        r = q * q * q * q;
        for (int c = 0; c < 3; ++c) Jr[c] = -2.0 * q * q * q * Cr[HC0 + c];
    };
    // J_k travels through the (still idle) stage buffers, node k + 1's copy in flight while node k is reduced: read straight
    // from global memory the dependent loads of this loop were 13 % of the kernel's warp time (ncu source samples, r1g).
    {
        auto jslot = [&](int k) { return stg + (k & 1) * S::STAGE_DOUBLES; };
        auto fetch = [&](int k) {
            pipe.begin(L, k & 1, NY * NW * 8);
            pipe.template copy<NY * NW>(L, k & 1, jslot(k), P.J + (size_t)k * NY * NW);
            pipe.commit(L);
        };
        static_assert((NY * NW) % 2 == 0 && NY * NW + 2 * NY <= S::STAGE_DOUBLES, "J_k + weights + residual fit a stage slot");
        fetch(0);
        // the lane's (i, j) pairs of the lower triangle, packed 4 + 4 bits each: looked up once, not once per node
        constexpr int TRI_PER_LANE = (TRI + Lanes::count() - 1) / Lanes::count();
        constexpr bool TRI_PACK = TRI_PER_LANE <= 4 && NW <= 16;
        unsigned ijp = 0;
        if (TRI_PACK)
            for (int m = 0; m < TRI_PER_LANE; ++m) {
                const int t = lane + m * NL;
                if (t < TRI) ijp |= ((unsigned)tri.i[t] | ((unsigned)tri.j[t] << 4)) << (8 * m);
            }
        double w_nx = 0.0, y_nx = 0.0;                   // lane r < NY: weight and residual of row r of the next node
        if (lane < NY || NL == 1) { w_nx = wrow(0, NL == 1 ? 0 : lane); y_nx = P.yres[NL == 1 ? 0 : lane]; }
        for (int k = 0; k <= N; ++k) {
            pipe.wait(L, k & 1);
            const double* Jk = jslot(k);
            const double sck = scale(k);                // requested ahead of the lane barrier below
            double* wk = jslot(k) + NY * NW;            // [NY] weights | [NY] residual
            if (NL == 1) {
                for (int r = 0; r < NY; ++r) { wk[r] = wrow(k, r); wk[NY + r] = P.yres[(size_t)k * NY + r]; }
            } else if (lane < NY) {
                wk[lane] = w_nx; wk[NY + lane] = y_nx;
            }
            if (k < N) {
                fetch(k + 1);
                if (NL != 1 && lane < NY) { w_nx = wrow(k + 1, lane); y_nx = P.yres[(size_t)(k + 1) * NY + lane]; }
            }
            L.sync();
            for (int t = lane, m = 0; t < TRI; t += NL, ++m) {
                const int i = TRI_PACK ? (int)((ijp >> (8 * m)) & 15u) : (int)tri.i[t];
                const int j = TRI_PACK ? (int)((ijp >> (8 * m + 4)) & 15u) : (int)tri.j[t];
                double s = 0.0;
                for (int r = 0; r < NY; ++r) s += Jk[r * NW + i] * wk[r] * Jk[r * NW + j];
                if (i == j && (k <= NR || i < NU)) s += P.lm;      // auxiliary nodes: inputs only
                if (sdf_cost && k < NR && i >= HC0 && i < HC0 + 3 && j >= HC0 && j < HC0 + 3) {
                    double r, Jr[3];
                    cost_r(k, r, Jr);
                    s += P.sdf_cost_w * Jr[i - HC0] * Jr[j - HC0];
                }
                s *= sck;
                Hk(k)[i * NW + j] = s;
                Hk(k)[j * NW + i] = s;
            }
            for (int i = lane; i < NW; i += NL) {
                double s = 0.0;
                for (int r = 0; r < NY; ++r) s += Jk[r * NW + i] * wk[r] * wk[NY + r];
                if (sdf_cost && k < NR && i >= HC0 && i < HC0 + 3) {
                    double r, Jr[3];
                    cost_r(k, r, Jr);
                    s += P.sdf_cost_w * r * Jr[i - HC0];
                }
                g[k * NW + i] = sck * s;
                const double w0 = (k == 0 && i >= NU) ? P.x0[i - NU] - P.xbar[i - NU] : 0.0;
                Wsm[i * KP + k] = w0;
                dW[i * KP + k] = 0.0;
            }
        }
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
        return a < NH ? (P.nlh ? P.nlh[kk * NH + r] : P.lh[r]) - hv : hv - (P.nuh ? P.nuh[kk * NH + r] : P.uh[r]);
    };
    auto c_pos = [&](int kk, int r, int c) { return cpos[(kk * NH + r) * 3 + c]; };

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
#elif !defined(SDFNMPC_QP_NO_PARK)
    // The node's row state (40 doubles) and row steps (40) are live only inside the row phases; between them - during the
    // Riccati sweeps, three quarters of the time - they are parked in L2 ([item][node]: the lanes' accesses coalesce), which
    // leaves the sweeps the whole register file: 234 registers and no spills instead of 255 + 434 B, QP 5.21 -> 4.65 ms per
    // 4096 solves (-DSDFNMPC_QP_NO_PARK keeps the state in registers; 10 warps / SM at 200 registers: 4.82 ms).
    // Even the two back-to-back row phases (steps formed / steps consumed) go through L2: keeping the 80 doubles live across
    // them brings back 288 B of spills and costs 4.84 vs 3.93 ms.
    double* TB = take((size_t)(4 * NU + 8 * NH) * KP);
    double* TD = take((size_t)(4 * NU + 8 * NH) * KP);
#define SDF_PARK_LD(p_) __ldcg(p_)
#define SDF_PARK_ST(p_, v_) __stcg(p_, v_)
#define SDF_LOAD_DELTA(kk)                                                                                                   \
    {                                                                                                                        \
        const double* q_ = TD + (kk);                                                                                        \
        _Pragma("unroll") for (int a = 0; a < 2 * NU; ++a) { dtb[a] = SDF_PARK_LD(q_ + a * KP); dlb[a] = SDF_PARK_LD(q_ + (2 * NU + a) * KP); } \
        _Pragma("unroll") for (int a = 0; a < 2 * NH; ++a) {                                                                 \
            dts[a] = SDF_PARK_LD(q_ + (4 * NU + a) * KP); dls[a] = SDF_PARK_LD(q_ + (4 * NU + 2 * NH + a) * KP);              \
            dss[a] = SDF_PARK_LD(q_ + (4 * NU + 4 * NH + a) * KP); dms[a] = SDF_PARK_LD(q_ + (4 * NU + 6 * NH + a) * KP);     \
        }                                                                                                                    \
    }
#define SDF_STORE_DELTA(kk)                                                                                                  \
    {                                                                                                                        \
        double* q_ = TD + (kk);                                                                                              \
        _Pragma("unroll") for (int a = 0; a < 2 * NU; ++a) { SDF_PARK_ST(q_ + a * KP, dtb[a]); SDF_PARK_ST(q_ + (2 * NU + a) * KP, dlb[a]); } \
        _Pragma("unroll") for (int a = 0; a < 2 * NH; ++a) {                                                                 \
            SDF_PARK_ST(q_ + (4 * NU + a) * KP, dts[a]); SDF_PARK_ST(q_ + (4 * NU + 2 * NH + a) * KP, dls[a]);                \
            SDF_PARK_ST(q_ + (4 * NU + 4 * NH + a) * KP, dss[a]); SDF_PARK_ST(q_ + (4 * NU + 6 * NH + a) * KP, dms[a]);       \
        }                                                                                                                    \
    }
#define SDF_LOAD_STATE(kk)                                                                                                   \
    {                                                                                                                        \
        const double* q_ = TB + (kk);                                                                                        \
        _Pragma("unroll") for (int a = 0; a < 2 * NU; ++a) { tb[a] = SDF_PARK_LD(q_ + a * KP); lb[a] = SDF_PARK_LD(q_ + (2 * NU + a) * KP); } \
        _Pragma("unroll") for (int a = 0; a < 2 * NH; ++a) {                                                                 \
            ts[a] = SDF_PARK_LD(q_ + (4 * NU + a) * KP); ls[a] = SDF_PARK_LD(q_ + (4 * NU + 2 * NH + a) * KP);                \
            ss[a] = SDF_PARK_LD(q_ + (4 * NU + 4 * NH + a) * KP); ms[a] = SDF_PARK_LD(q_ + (4 * NU + 6 * NH + a) * KP);       \
        }                                                                                                                    \
    }
#define SDF_STORE_STATE(kk)                                                                                                  \
    {                                                                                                                        \
        double* q_ = TB + (kk);                                                                                              \
        _Pragma("unroll") for (int a = 0; a < 2 * NU; ++a) { SDF_PARK_ST(q_ + a * KP, tb[a]); SDF_PARK_ST(q_ + (2 * NU + a) * KP, lb[a]); } \
        _Pragma("unroll") for (int a = 0; a < 2 * NH; ++a) {                                                                 \
            SDF_PARK_ST(q_ + (4 * NU + a) * KP, ts[a]); SDF_PARK_ST(q_ + (4 * NU + 2 * NH + a) * KP, ls[a]);                  \
            SDF_PARK_ST(q_ + (4 * NU + 4 * NH + a) * KP, ss[a]); SDF_PARK_ST(q_ + (4 * NU + 6 * NH + a) * KP, ms[a]);         \
        }                                                                                                                    \
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
                t = fmax(-hard_d(kk, a), SDFNMPC_QP_T0);          // w_u = 0
                l = fmax(SDFNMPC_QP_MU0 / t, 1e-3 * SDFNMPC_QP_MU0);
            }
            tb[a] = t; lb[a] = l;
        }
#pragma unroll
        for (int a = 0; a < 2 * NH; ++a) {
            const int r = a < NH ? a : a - NH;
            double t = 1.0, l = 1.0, m = 1.0, s0 = SDFNMPC_QP_S0;
            if (kk >= k0) {
                const double z = zrow(kk, r), Z = Zrow(kk, r);
                double cw = 0.0;
                for (int c = 0; c < 3; ++c) cw += c_pos(kk, r, c) * Wsm[(HC0 + c) * KP + kk];
                const double gw = a < NH ? cw : -cw;
                t = fmax(gw - soft_d(kk, a) + s0, SDFNMPC_QP_T0);
                l = fmax(0.5 * (z + Z * s0), 1e-3);     // lam = mu: slack stationarity z + Z s - lam - mu = 0 split evenly
                m = l;
                // lifted rows with the exact-penalty weight z_hard: an even split starts the row multiplier at z_hard / 2, thousands of
                // times its value at the solution; start it at the cap and give the rest to the slack's multiplier (still stationary).
                // Host emulation, three variants of the lifted OCP: 12.7 -> 10.6 iterations.  On the plain OCP (terminal rows start at
                // 112, stage rows at 8) the cap saves 2 - 4 % of the iterations: 8.73 -> 8.48 at 4096 instances, 467 -> 423 us for one
                // control step at batch 1 (SDFNMPC_QP_LCAP_ALL=0 restores the even split there).
                const double lcap = P.nzs ? SDFNMPC_QP_LCAP_LIFT : SDFNMPC_QP_LCAP;
                if ((P.nzs || SDFNMPC_QP_LCAP_ALL) && l > lcap) { l = lcap; m = z + Z * s0 - l; }
            }
            ts[a] = t; ss[a] = s0; ls[a] = l; ms[a] = m;
        }
        SDF_STORE_STATE(kk);
        // rv0_k = H_k w_k + g_k - sum lam_i gvec_i   (w = 0 except dx_0; costates 0)
        for (int i = 0; i < NW; ++i) {
            double s = g[(size_t)kk * NW + i];
            if (kk == 0)
                for (int j = NU; j < NW; ++j) s += Hk(0)[i * NW + j] * Wsm[j * KP];
            if (kk < N && i < NU) s -= lb[i] - lb[NU + i];
            if (kk >= k0 && i >= HC0 && i < HC0 + 3)
                for (int r = 0; r < NH; ++r) s -= (ls[r] - ls[NH + r]) * c_pos(kk, r, i - HC0);
            if ((kk == 0 && i >= NU) || (kk == N && i < NU)) s = 0.0;
            rv0k(kk)[i] = s;
        }
        // rd0_{k} (k >= 1) = AB_{k-1} w_{k-1} + b_{k-1} - w_x,k ; only w_x,0 is non-zero
        if (kk >= 1)
            for (int i = 0; i < NX; ++i) {
                double s = P.bdyn[(kk - 1) * NX + i];
                if (kk == 1)
                    for (int j = NU; j < NW; ++j) s += P.AB[(size_t)i * NW + j] * Wsm[j * KP];
                rd0k(kk)[i] = s;
            }
    }
    L.sync();
    double res_v0 = 0.0;
    for (int e = lane; e < (N + 1) * NW; e += NL) { const int k = e / NW; res_v0 = fmax(res_v0, fabs(rv0k(k)[e - k * NW])); }
    for (int e = lane; e < N * NX; e += NL) { const int k = e / NX; res_v0 = fmax(res_v0, fabs(rd0k(k + 1)[e - k * NX])); }
    res_v0 = L.max(res_v0);

    int status = 2, it = 0;
    // (i, j) of the lane's entries of the lower triangle of M and the correction rows they receive (ca: diagonal input
    // entries, cb: position block), 4 bits each, resolved once instead of once per stage and iteration
    auto m_entry = [&](int e, int& i, int& j, int& ca, int& cb) {
        i = tri.i[e]; j = tri.j[e];
        ca = (i == j && i < NU) ? i : 15;
        const int pi_ = i - HC0, pj = j - HC0;
        cb = (pi_ >= 0 && pi_ < 3 && pj >= 0 && pj < 3) ? sym3(pi_, pj) : 15;
    };
    constexpr bool M_PACK = (TRI + NW + Lanes::count() - 1) / Lanes::count() <= 4 && NW <= 15;
    unsigned long long mpk = 0;
    if (M_PACK)
        for (int m = 0; m < 4; ++m) {
            const int e = lane + m * NL;
            if (e < TRI) {
                int i, j, ca, cb;
                m_entry(e, i, j, ca, cb);
                mpk |= (unsigned long long)(i | (j << 4) | (ca << 8) | (cb << 12)) << (16 * m);
            }
        }
#if defined(SDFNMPC_QP_TRACE) && defined(__CUDA_ARCH__)
    long long qpt_last = clock64();
    int qpt_pass = 0;
#define QPT(id)                                                                          \
    {                                                                                    \
        if (P.trace && lane == 0) {                                                      \
            const long long now_ = clock64();                                            \
            if (it == 2) P.trace[qpt_pass * 16 + (id)] += now_ - qpt_last;               \
            qpt_last = now_;                                                             \
        }                                                                                \
    }
#else
#define QPT(id)
#endif
    bool safe = false;              // conservative mode after a stall: more centering, no second-order term
    double cm_prev1 = 1e300, cm_prev2 = 1e300;
    double phi = 1.0;               // product of (1 - alpha): scale of the stationarity / dynamics residuals
    double res_track = 0.0;
    for (;; ++it) {
        double tau = 0.0, mubar = 0.0;
        bool stop = false;
        for (int pass = 0; pass < 2 && !stop; ++pass) {
#if defined(SDFNMPC_QP_TRACE) && defined(__CUDA_ARCH__)
            qpt_pass = pass;
#endif
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
                        const double z = zrow(kk, r), Z = Zrow(kk, r);
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
#elif defined(SDFNMPC_QP_DEBUG_INST)
                if (P.debug && lane == 0) printf("it %d res %.3e (rows %.3e) cmax %.3e phi %.3e mubar %.3e\n", it, res_track, res_rows, cmax, phi, mubar);
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
                QPT(0);
            // =========== backward Riccati sweep (pass 0: factorise + solve; pass 1: new right-hand side only) ===========
            if (pass == 0) {
                SDF_LANES_FOR(e, NX * NX) {
                    const int i = e / NX, j = e - i * NX;
                    double s = Hk(N)[(NU + i) * NW + NU + j];
                    const int pi_ = i + NU - HC0, pj = j + NU - HC0;
                    if (pi_ >= 0 && pi_ < 3 && pj >= 0 && pj < 3) s += corr[(NU + sym3(pi_, pj)) * KP + N];
                    Psm[e] = s;
                }
            }
            SDF_LANES_FOR(e, NX) {
                double s = phi * rv0k(N)[NU + e];
                const int pe = e + NU - HC0;
                if (pe >= 0 && pe < 3) s += corr[(2 * NU + 6 + pe) * KP + N];
                pv[e] = s;
            }
            L.sync();
                QPT(2);
            // stage data of node kk is prefetched into the other half of `stg` while node kk+1 is processed
            // (every stage body ends with a lane barrier, so the slot being refilled is no longer read)
            auto prefetch = [&](int kk, int slot) {
                double* dst = stg + slot * S::STAGE_DOUBLES;
                if (pass == 0) {
                    pipe.begin(L, slot, (NX * NW + S::REC_A) * 8);
                    pipe.template copy<NX * NW>(L, slot, dst + S::ST_AB, P.AB + (size_t)kk * NX * NW);
                    pipe.template copy<S::REC_A>(L, slot, dst + S::ST_A, recA + (size_t)kk * S::REC_A);
                } else {
                    pipe.begin(L, slot, (NX * NW + NW + S::REC_F) * 8);
                    pipe.template copy<NX * NW>(L, slot, dst + S::ST_AB, P.AB + (size_t)kk * NX * NW);
                    pipe.template copy<NW>(L, slot, dst + S::ST_RV, rv0k(kk));
                    pipe.template copy<S::REC_F>(L, slot, dst + S::ST_F, recF + (size_t)kk * S::REC_F);
                }
                pipe.commit(L);
            };
            pipe.publish(L);
            prefetch(N - 1, 0);
            for (int kk = N - 1; kk >= 0; --kk) {
                const int slot = (N - 1 - kk) & 1;
                const double* sb = stg + slot * S::STAGE_DOUBLES;
                pipe.wait(L, slot);
                QPT(3);
                if (kk > 0) prefetch(kk - 1, slot ^ 1);
                const double* ABk = sb + S::ST_AB;
                if (pass == 0) {
                    // PAB = P [A B | phi rd], read straight from the staged [A B] block and the staged residual
                    SDF_LANES_FOR(e, NX * NWB) {
                        double s = 0.0;
                        if (e < NX * NW) {
                            const int i = e / NW, j = e - i * NW;
#pragma unroll
                            for (int l = 0; l < NX; l += 2) {        // rows of P: 16-byte aligned pairs
                                const Pair pr = ld2(Psm + i * NX + l);
                                s += pr.x * ABk[l * NW + j];
                                s += pr.y * ABk[(l + 1) * NW + j];
                            }
                            PAB[i * NWB + j] = s;
                        } else {
                            const int i = e - NX * NW;
#pragma unroll
                            for (int l = 0; l < NX; l += 2) {
                                const Pair pr = ld2(Psm + i * NX + l), rd = ld2(sb + S::ST_RD + l);
                                s += pr.x * (phi * rd.x);
                                s += pr.y * (phi * rd.y);
                            }
                            PbMk(kk)[i] = s;
                            PAB[i * NWB + NW] = s + pv[i];
                        }
                    }
                    L.sync();
                QPT(5);
                    SDF_LANES_FOR(e, TRI + NW) {
                        if (e < TRI) {
                            int i, j, ca, cb;
                            if (M_PACK) {
                                const unsigned f = (unsigned)(mpk >> (16 * e_it));
                                i = f & 15u; j = (f >> 4) & 15u; ca = (f >> 8) & 15u; cb = (f >> 12) & 15u;
                            } else {
                                m_entry(e, i, j, ca, cb);
                            }
                            double s = sb[S::ST_H + i * NW + j];
#pragma unroll
                            for (int l = 0; l < NX; ++l) s += ABk[l * NW + i] * PAB[l * NWB + j];
                            if (ca != 15) s += corr[ca * KP + kk];
                            if (cb != 15) s += corr[(NU + cb) * KP + kk];
                            Msm[i * NW + j] = s;
                            Msm[j * NW + i] = s;
                        } else {
                            const int i = e - TRI;
                            double s = phi * sb[S::ST_RV + i];
#pragma unroll
                            for (int l = 0; l < NX; ++l) s += ABk[l * NW + i] * PAB[l * NWB + NW];
                            if (i < NU) s += corr[(NU + 6 + i) * KP + kk];
                            const int pi_ = i - HC0;
                            if (pi_ >= 0 && pi_ < 3) s += corr[(2 * NU + 6 + pi_) * KP + kk];
                            rv[i] = s;
                        }
                    }
                    L.sync();
                QPT(6);
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
                            if (i == j) { ok = ok && (s > 0.0); Lr[i][i] = rsqrt_d(s); }
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
                            if (c < NX) { Ksm[i * NX + c] = -x[i]; Kgk(kk)[i * NX + c] = -x[i]; }
                            else kfs[i] = -x[i];
                        }
                    }
                    if (lane == 0)
                        for (int i = 0; i < NU; ++i)
                            for (int j = 0; j < NU; ++j) Lgk(kk)[i * NU + j] = j <= i ? Lr[i][j] : 0.0;
                    L.sync();
                QPT(7);
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
                QPT(8);
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
                QPT(9);
                }
            }
            // ---------- forward roll-out: dW = step ----------
            SDF_LANES_FOR(i, NX) dW[(NU + i) * KP] = 0.0;      // dx_0 = 0 (w_x0 already equals x0 - xbar_0)
            L.sync();
                QPT(10);
            auto prefetch_f = [&](int kk, int slot) {
                double* dst = stg + slot * S::STAGE_DOUBLES;
                pipe.begin(L, slot, (NX * NW + NX + NU * NX) * 8);
                pipe.template copy<NX * NW>(L, slot, dst + S::ST_AB, P.AB + (size_t)kk * NX * NW);
                pipe.template copy<NX>(L, slot, dst + S::ST_RD, rd0k(kk + 1));
                pipe.template copy<NU * NX>(L, slot, dst + S::ST_K, Kgk(kk));
                pipe.commit(L);
            };
            pipe.publish(L);
            prefetch_f(0, 0);
            for (int kk = 0; kk < N; ++kk) {
                const int slot = kk & 1;
                const double* sb = stg + slot * S::STAGE_DOUBLES;
                pipe.wait(L, slot);
                QPT(11);
                if (kk + 1 < N) prefetch_f(kk + 1, slot ^ 1);
                const double* Kk = sb + S::ST_K;
                SDF_LANES_FOR(j, NU) {
                    double s = dW[j * KP + kk];
#pragma unroll
                    for (int i = 0; i < NX; i += 2) {
                        const Pair kr = ld2(Kk + j * NX + i);
                        s += kr.x * dW[(NU + i) * KP + kk];
                        s += kr.y * dW[(NU + i + 1) * KP + kk];
                    }
                    dW[j * KP + kk] = s;
                }
                L.sync();
                const double* ABk = sb + S::ST_AB;
                SDF_LANES_FOR(i, NX) {
                    double s = phi * sb[S::ST_RD + i];
#pragma unroll
                    for (int j = 0; j < NW; j += 2) {
                        const Pair ar = ld2(ABk + i * NW + j);
                        s += ar.x * dW[j * KP + kk];
                        s += ar.y * dW[(j + 1) * KP + kk];
                    }
                    dW[(NU + i) * KP + kk + 1] = s;
                }
                L.sync();
                QPT(12);
            }
            // =========== row steps (lane = node): the node's (dt, dlam, ds, dmu) stay in registers ===========
            double rmax = 0.0;              // largest -dx/x over the lane's rows
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
                            dlb[a] = (kap - lb[a] * dtb[a]) * ev.rt;
                            rmax = fmax(rmax, fmax(boundary_ratio(dtb[a], ev.rt), boundary_ratio(dlb[a], ev.rl)));
                        }
                    }
                }
                if (kk >= k0) {
#pragma unroll
                    for (int r = 0; r < NH; ++r) {
                        const double c0 = c_pos(kk, r, 0), c1 = c_pos(kk, r, 1), c2 = c_pos(kk, r, 2);
                        const double cw = c0 * Wsm[(HC0 + 0) * KP + kk] + c1 * Wsm[(HC0 + 1) * KP + kk] + c2 * Wsm[(HC0 + 2) * KP + kk];
                        const double cd = c0 * dW[(HC0 + 0) * KP + kk] + c1 * dW[(HC0 + 1) * KP + kk] + c2 * dW[(HC0 + 2) * KP + kk];
                        const double z = zrow(kk, r), Z = Zrow(kk, r);
#pragma unroll
                        for (int side = 0; side < 2; ++side) {
                            const int a = side * NH + r;
                            const double sgn = side == 0 ? 1.0 : -1.0;
                            const double ct = pass == 0 ? 0.0 : tau - prod[(2 * NU + a) * KP + kk];
                            const double cs = pass == 0 ? 0.0 : tau - prod[(2 * NU + 2 * NH + a) * KP + kk];
                            const SoftEval ev = soft_eval(ts[a], ls[a], ss[a], ms[a], sgn * cw, soft_d(kk, a), z, Z, ct, cs);
                            const double kap = ct - ls[a] * ts[a], nu = cs - ms[a] * ss[a];
                            const double gd = sgn * cd;
                            dss[a] = (ev.rho - ev.Th * gd) * ev.rD;
                            dts[a] = gd + dss[a] - ev.r_p;
                            dls[a] = (kap - ls[a] * dts[a]) * ev.rt;
                            dms[a] = (nu - ms[a] * dss[a]) * ev.rs;
                            rmax = fmax(rmax, fmax(fmax(boundary_ratio(dts[a], ev.rt), boundary_ratio(dls[a], ev.rl)),
                                                   fmax(boundary_ratio(dss[a], ev.rs), boundary_ratio(dms[a], ev.rm))));
                        }
                    }
                }
                SDF_STORE_DELTA(kk);
            }
            rmax = L.max(rmax);
            const double amax = rmax > 1e-30 ? 1.0 / rmax : 1e30;
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
                QPT(14);
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
                QPT(15);
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
        // stage data (H_k inside record A, [A B]_k) travels through the stage ring like in the sweeps; the costates ping-pong
        // between two shared vectors.  Not run after a numerical breakdown (copies of the aborted sweep may be in flight).
        auto fetch_v = [&](int kk, int slot) {
            double* dst = stg + slot * S::STAGE_DOUBLES;
            pipe.begin(L, slot, ((kk < N ? NX * NW : 0) + S::REC_A) * 8);
            if (kk < N) pipe.template copy<NX * NW>(L, slot, dst + S::ST_AB, P.AB + (size_t)kk * NX * NW);
            pipe.template copy<S::REC_A>(L, slot, dst + S::ST_A, recA + (size_t)kk * S::REC_A);
            pipe.commit(L);
        };
        if (status != 4) {
            pipe.publish(L);
            fetch_v(N, 0);
        }
        for (int kk = N; kk >= 0 && status != 4; --kk) {
            const int slot = (N - kk) & 1;
            const double* sb = stg + slot * S::STAGE_DOUBLES;
            pipe.wait(L, slot);
            if (kk > 0) fetch_v(kk - 1, slot ^ 1);
            const double* pin = slot ? pv : vv;       // costate of node kk + 1
            double* pout = slot ? vv : pv;            // costate of node kk
            const double bd = (kk < N && lane < NX) ? P.bdyn[kk * NX + lane] : 0.0;
            for (int i = lane; i < NW; i += NL) {
                if (kk == N && i < NU) continue;
                double s = g[(size_t)kk * NW + i];
                for (int j = 0; j < NW; ++j) s += sb[S::ST_H + i * NW + j] * Wsm[j * KP + kk];
                if (i < NU) s -= corr[i * KP + kk];
                if (i >= HC0 && i < HC0 + 3)
                    for (int r = 0; r < NH; ++r) s -= corr[(NU + r) * KP + kk] * c_pos(kk, r, i - HC0);
                if (kk < N)
                    for (int l = 0; l < NX; ++l) s += sb[S::ST_AB + l * NW + i] * pin[l];
                if (i < NU) res_true = fmax(res_true, fabs(s));      // input stationarity
                else pout[i - NU] = s;                               // defines the costate of node kk (x-stationarity == 0)
            }
            if (kk < N)
                for (int i = lane; i < NX; i += NL) {
                    double s = (NL == 1 ? P.bdyn[kk * NX + i] : bd) - Wsm[(NU + i) * KP + kk + 1];
                    for (int j = 0; j < NW; ++j) s += sb[S::ST_AB + i * NW + j] * Wsm[j * KP + kk];
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

#undef QPT
#undef SDF_NODE_LOOP
#undef SDF_LANES_FOR
#undef SDF_LOAD_STATE
#undef SDF_STORE_STATE
#undef SDF_LOAD_DELTA
#undef SDF_STORE_DELTA

}  // namespace sdfnmpc
