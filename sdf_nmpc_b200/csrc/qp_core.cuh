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
// Theta~_i g_i g_i' to the stage Hessian and c_i g_i to the stage gradient; the remaining
// equality-constrained QP is solved exactly by the Riccati recursion for the *target*
// point w+ (so the linear dynamics never need an explicit residual in the step).
#pragma once
#include "common.cuh"

namespace sdfnmpc {

template <int NX, int NU, int NY, int NH>
struct QpShape {
    static constexpr int NW = NX + NU;
    SDF_HD static int hard_rows(int N) { return N * 2 * NU; }
    SDF_HD static int soft_rows(int N) { return (N + 1) * 2 * NH; }
    // doubles of per-instance scratch
    SDF_HD static size_t work_doubles(int N) {
        size_t n1 = (size_t)(N + 1);
        return 2 * n1 * NW * NW + 2 * n1 * NW      // H, Ht, g, gt
               + 2 * n1 * NW + 2 * n1 * NX         // w, wp, pi, pip
               + 5 * (size_t)hard_rows(N)          // tb lb dtb dlb pab
               + 10 * (size_t)soft_rows(N)         // ts ls ss ms dts dls dss dms pas pbs
               + 2 * (size_t)N * NU + 2 * n1 * NH  // net Theta / c per node
               + (size_t)N * (NU * NX + NU * NU + NX + NU)    // K, Lc, PbM, kff
               + n1 * NW + n1 * NX                            // rvec, rdyn
               + (size_t)hard_rows(N) + 2 * (size_t)soft_rows(N);   // rpb, rps, rhos
    }
    static constexpr int SMEM_DOUBLES = NX * NX + 2 * NX * (NW + 1) + NW * NW + NW + NU * NX + NU + 2 * NX + 4;
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
    int* status;           // 0 ok, 2 max-iter, 4 NaN / factorisation failure (acados codes)
    int* iters;
};

namespace qpdetail {

SDF_HD void tri_decode(int e, int& i, int& j) {   // e -> (i >= j), row-major lower triangle
    int r = (int)((sqrt(8.0 * (double)e + 1.0) - 1.0) * 0.5);
    while ((r + 1) * (r + 2) / 2 <= e) ++r;
    while (r * (r + 1) / 2 > e) --r;
    i = r;
    j = e - r * (r + 1) / 2;
}

struct SoftRow {
    double r_p, r_s, Th, D, Tht, zs;
};

SDF_HD SoftRow soft_row(double t, double lam, double s, double mu, double gw, double d, double z, double Z) {
    SoftRow r;
    r.r_p = t - (gw - d + s);
    r.r_s = z + Z * s - lam - mu;
    r.Th = lam / t;
    r.zs = Z + mu / s;
    r.D = r.zs + r.Th;
    r.Tht = r.Th * r.zs / r.D;
    return r;
}

SDF_HD double step_to_boundary(double x, double dx) { return dx < 0.0 ? -x / dx : 1e30; }

}  // namespace qpdetail

template <int NX, int NU, int NY, int NH>
SDF_HD void qp_solve_instance(const Lanes& L, const QpProb<NX, NU, NY, NH>& P, double* sm) {
    using namespace qpdetail;
    using S = QpShape<NX, NU, NY, NH>;
    constexpr int NW = S::NW;
    constexpr int NWB = NW + 1;
    const int N = P.N, k0 = P.k0;
    const int NL = Lanes::count();
    const int lane = L.lane;
    const int nhr = S::hard_rows(N), nsr = S::soft_rows(N);
    const int ncomp = nhr + 2 * (N + 1 - k0) * 2 * NH;

    // ---- carve the per-instance scratch ------------------------------------------------
    double* p = P.work;
    auto take = [&](size_t n) { double* r = p; p += n; return r; };
    const size_t n1 = (size_t)(N + 1);
    double* H = take(n1 * NW * NW);
    double* Ht = take(n1 * NW * NW);
    double* g = take(n1 * NW);
    double* gt = take(n1 * NW);
    double* w = take(n1 * NW);
    double* wp = take(n1 * NW);
    double* pi = take(n1 * NX);
    double* pip = take(n1 * NX);
    double *tb = take(nhr), *lb = take(nhr), *dtb = take(nhr), *dlb = take(nhr), *pab = take(nhr);
    double *ts = take(nsr), *ls = take(nsr), *ss = take(nsr), *ms = take(nsr);
    double *dts = take(nsr), *dls = take(nsr), *dss = take(nsr), *dms = take(nsr), *pas = take(nsr), *pbs = take(nsr);
    double *thnb = take((size_t)N * NU), *cnb = take((size_t)N * NU);
    double *thns = take(n1 * NH), *cns = take(n1 * NH);
    double* Kg = take((size_t)N * NU * NX);
    double* Lg = take((size_t)N * NU * NU);
    double* PbM = take((size_t)N * NX);
    double* kffg = take((size_t)N * NU);
    double *rpb = take(nhr), *rps = take(nsr), *rhos = take(nsr);   // r_p / rho per row, shared by both uses
    double* rvec = take(n1 * NW);     // stationarity residual per node
    double* rdyn = take(n1 * NX);     // dynamics residual: rdyn[0] = dx0 - w_x0, rdyn[k+1] = AB_k w_k + b_k - w_x,k+1

    // ---- shared (per-lane-group) workspace ------------------------------------------------
    double* Psm = sm;                    // [NX][NX]
    double* ABb = Psm + NX * NX;         // [NX][NW+1]
    double* PAB = ABb + NX * NWB;        // [NX][NW+1]
    double* Msm = PAB + NX * NWB;        // [NW][NW]
    double* rv = Msm + NW * NW;          // [NW]
    double* Ksm = rv + NW;               // [NU][NX]
    double* kfs = Ksm + NU * NX;         // [NU]
    double* pv = kfs + NU;               // [NX]
    double* vv = pv + NX;                // [NX]

    auto scale = [&](int k) { return k < N ? P.dt[k] : 1.0; };
    auto wrow = [&](int k, int r) { return k < N ? P.W[k * NY + r] : (r < P.nyN ? P.WN[r] : 0.0); };
    // row data:   hard row index e = k*2NU + a  (a < NU lower, else upper);  soft e = k*2NH + a
    auto hard_d = [&](int k, int a) {
        int j = a < NU ? a : a - NU;
        double ub = P.ubar[k * NU + j];
        return a < NU ? P.lbu[j] - ub : ub - P.ubu[j];
    };
    auto soft_d = [&](int k, int a) {
        int r = a < NH ? a : a - NH;
        double hv = P.hval[k * NH + r];
        return a < NH ? P.lh[r] - hv : hv - P.uh[r];
    };
    auto c_dot = [&](int k, int r, const double* vec) {   // C_r . vec_k
        const double* Cr = P.C + ((size_t)k * NH + r) * NW;
        const double* v = vec + (size_t)k * NW;
        double s = 0.0;
        for (int i = 0; i < NW; ++i) s += Cr[i] * v[i];
        return s;
    };

    // ---- prologue: sdf row of C / h, Gauss-Newton Hessian and gradient ----------------------
    for (int k = k0 + lane; k <= N; k += NL) {
        const double* R = P.pose + 3;
        double gx = P.sdf_g[k * 3 + 0], gy = P.sdf_g[k * 3 + 1], gz = P.sdf_g[k * 3 + 2];
        double* Cr = P.C + ((size_t)k * NH + (NH - 1)) * NW;
        for (int jj = 0; jj < 3; ++jj) Cr[NU + jj] = P.flag * (R[jj * 3 + 0] * gx + R[jj * 3 + 1] * gy + R[jj * 3 + 2] * gz);
        P.hval[k * NH + (NH - 1)] = P.flag * (double)P.sdf_d[k] + (1.0 - P.flag) * P.max_df;
    }
    for (int e = lane; e < (N + 1) * NW * NW; e += NL) {
        int k = e / (NW * NW), ij = e % (NW * NW), i = ij / NW, j = ij % NW;
        const double* Jk = P.J + (size_t)k * NY * NW;
        double s = 0.0;
        for (int r = 0; r < NY; ++r) s += Jk[r * NW + i] * wrow(k, r) * Jk[r * NW + j];
        if (i == j) s += P.lm;
        H[e] = scale(k) * s;
    }
    for (int e = lane; e < (N + 1) * NW; e += NL) {
        int k = e / NW, i = e % NW;
        const double* Jk = P.J + (size_t)k * NY * NW;
        const double* yr = P.yres + (size_t)k * NY;
        double s = 0.0;
        for (int r = 0; r < NY; ++r) s += Jk[r * NW + i] * wrow(k, r) * yr[r];
        g[e] = scale(k) * s;
        w[e] = (k == 0 && i >= NU) ? P.x0[i - NU] - P.xbar[i - NU] : 0.0;
    }
    for (int e = lane; e < (N + 1) * NX; e += NL) pi[e] = 0.0;
    L.sync();
    // ---- initial point -----------------------------------------------------------------------
    for (int e = lane; e < nhr; e += NL) {
        int k = e / (2 * NU), a = e % (2 * NU);
        double r = -hard_d(k, a);          // w_u = 0
        double t = fmax(r, 0.5);
        tb[e] = t;
        lb[e] = fmax(1.0 / t, 1e-3);
    }
    for (int e = lane; e < nsr; e += NL) {
        int k = e / (2 * NH), a = e % (2 * NH);
        int r = a < NH ? a : a - NH;
        double sgn = a < NH ? 1.0 : -1.0;
        double z = P.zs[r] * scale(k), Z = P.Zs[r] * scale(k);
        double s0 = 0.5;
        double gw = k >= k0 ? sgn * c_dot(k, r, w) : 0.0;
        double t = fmax(gw - soft_d(k, a) + s0, 0.5);
        double l0 = fmax(0.5 * (z + Z * s0), 1e-3);
        ts[e] = t; ss[e] = s0; ls[e] = l0; ms[e] = l0;
    }
    L.sync();

    int status = 2, it = 0;
    bool safe = false;              // conservative mode after a stall: more centering, no second-order term
    double cm_prev1 = 1e300, cm_prev2 = 1e300;
    for (;; ++it) {
        // ================= residuals and convergence =================
        double res = 0.0, cmax = 0.0, csum = 0.0;
        for (int e = lane; e < nhr; e += NL) {
            int k = e / (2 * NU), a = e % (2 * NU);
            int j = a < NU ? a : a - NU;
            double gw = (a < NU ? 1.0 : -1.0) * w[k * NW + j];
            res = fmax(res, fabs(tb[e] - (gw - hard_d(k, a))));
            double c = lb[e] * tb[e];
            cmax = fmax(cmax, c); csum += c;
        }
        for (int e = lane + k0 * 2 * NH; e < nsr; e += NL) {
            int k = e / (2 * NH), a = e % (2 * NH);
            int r = a < NH ? a : a - NH;
            double gw = (a < NH ? 1.0 : -1.0) * c_dot(k, r, w);
            SoftRow sr = soft_row(ts[e], ls[e], ss[e], ms[e], gw, soft_d(k, a), P.zs[r] * scale(k), P.Zs[r] * scale(k));
            res = fmax(res, fmax(fabs(sr.r_p), fabs(sr.r_s)));
            double c1 = ls[e] * ts[e], c2 = ms[e] * ss[e];
            cmax = fmax(cmax, fmax(c1, c2)); csum += c1 + c2;
        }
        // stationarity (free variables only: dx_0 is fixed) and dynamics
        for (int e = lane; e < (N + 1) * NW; e += NL) {
            int k = e / NW, i = e % NW;
            if ((k == 0 && i >= NU) || (k == N && i < NU)) { rvec[e] = 0.0; continue; }
            const double* Hk = H + (size_t)k * NW * NW + (size_t)i * NW;
            const double* wk = w + (size_t)k * NW;
            double s = g[e];
            for (int j = 0; j < NW; ++j) s += Hk[j] * wk[j];
            if (k < N) {
                if (i < NU) s -= lb[k * 2 * NU + i] - lb[k * 2 * NU + NU + i];
                const double* ABk = P.AB + (size_t)k * NX * NW;
                const double* pn = pi + (size_t)(k + 1) * NX;
                for (int l = 0; l < NX; ++l) s += ABk[l * NW + i] * pn[l];
            }
            if (k >= k0)
                for (int r = 0; r < NH; ++r)
                    s -= (ls[k * 2 * NH + r] - ls[k * 2 * NH + NH + r]) * P.C[((size_t)k * NH + r) * NW + i];
            if (i >= NU) s -= pi[(size_t)k * NX + (i - NU)];
            rvec[e] = s;
            res = fmax(res, fabs(s));
        }
        for (int e = lane; e < (N + 1) * NX; e += NL) {
            int k = e / NX, i = e % NX;
            double s;
            if (k == 0) {
                s = (P.x0[i] - P.xbar[i]) - w[NU + i];
            } else {
                const double* ABk = P.AB + (size_t)(k - 1) * NX * NW + (size_t)i * NW;
                const double* wk = w + (size_t)(k - 1) * NW;
                s = P.bdyn[(k - 1) * NX + i] - w[(size_t)k * NW + NU + i];
                for (int j = 0; j < NW; ++j) s += ABk[j] * wk[j];
            }
            rdyn[e] = s;
            res = fmax(res, fabs(s));
        }
#if defined(SDFNMPC_QP_DEBUG) && !defined(__CUDA_ARCH__)
        printf("it %d res %.3e cmax %.3e\n", it, res, cmax);
#endif
        res = L.max(res);
        cmax = L.max(cmax);
        const double mubar = L.sum(csum) / ncomp;
        if (!(res == res) || !(mubar == mubar) || res > 1e300) { status = 4; break; }
        if (res <= P.tol_res && cmax <= P.tol) { status = 0; break; }
        if (it >= P.max_iter) { status = 2; break; }

        L.sync();
        // two passes: 0 = affine (predictor, with factorisation), 1 = corrector (vector only).
        // The Newton system is solved in DELTA form (right-hand sides are residuals), so the step keeps
        // full relative accuracy when the iterate is close to the solution.
        // Mehrotra's corrector can cycle on poorly centred iterates; if the largest complementarity product
        // has not dropped by 30 % over two iterations switch (for good) to plain damped path following.
        if (it >= 2 && cmax > 0.7 * cm_prev2) safe = true;
        cm_prev2 = cm_prev1; cm_prev1 = cmax;
        const double so = safe ? 0.0 : 1.0;
        double tau = 0.0;
        for (int pass = 0; pass < 2; ++pass) {
            // ---------- per-row elimination coefficients -> net Theta~ and c per node ----------
            for (int e = lane; e < N * NU; e += NL) {
                int k = e / NU, j = e % NU;
                double thn = 0.0, cn = 0.0;
                for (int side = 0; side < 2; ++side) {
                    int row = k * 2 * NU + side * NU + j;
                    double sgn = side == 0 ? 1.0 : -1.0;
                    double t = tb[row], lam = lb[row];
                    double gw = sgn * w[k * NW + j];
                    double r_p = t - (gw - hard_d(k, side * NU + j));
                    rpb[row] = r_p;
                    double Th = lam / t;
                    double kap = (pass == 0 ? 0.0 : tau - pab[row]) - lam * t;
                    double coef = -kap / t - Th * r_p;
                    thn += Th;
                    cn += sgn * coef;
                }
                thnb[e] = thn;
                cnb[e] = cn;
            }
            for (int e = lane; e < (N + 1) * NH; e += NL) {
                int k = e / NH, r = e % NH;
                double thn = 0.0, cn = 0.0;
                if (k >= k0) {
                    double cw = c_dot(k, r, w);
                    double z = P.zs[r] * scale(k), Z = P.Zs[r] * scale(k);
                    for (int side = 0; side < 2; ++side) {
                        int row = k * 2 * NH + side * NH + r;
                        double sgn = side == 0 ? 1.0 : -1.0;
                        double gw = sgn * cw;
                        SoftRow sr = soft_row(ts[row], ls[row], ss[row], ms[row], gw, soft_d(k, side * NH + r), z, Z);
                        double kap = (pass == 0 ? 0.0 : tau - pas[row]) - ls[row] * ts[row];
                        double nu = (pass == 0 ? 0.0 : tau - pbs[row]) - ms[row] * ss[row];
                        double rho = -sr.r_s + kap / ts[row] + nu / ss[row] + sr.Th * sr.r_p;
                        rps[row] = sr.r_p;
                        rhos[row] = rho;
                        double coef = -kap / ts[row] + sr.Th * (rho / sr.D - sr.r_p);
                        thn += sr.Tht;
                        cn += sgn * coef;
                    }
                }
                thns[e] = thn;
                cns[e] = cn;
            }
            L.sync();
            // ---------- modified Hessian (pass 0 only) and gradient ----------
            if (pass == 0) {
                for (int e = lane; e < (N + 1) * NW * NW; e += NL) {
                    int k = e / (NW * NW), ij = e % (NW * NW), i = ij / NW, j = ij % NW;
                    double s = H[e];
                    if (k < N && i == j && i < NU) s += thnb[k * NU + i];
                    if (k >= k0)
                        for (int r = 0; r < NH; ++r) {
                            const double* Cr = P.C + ((size_t)k * NH + r) * NW;
                            s += thns[k * NH + r] * Cr[i] * Cr[j];
                        }
                    Ht[e] = s;
                }
            }
            for (int e = lane; e < (N + 1) * NW; e += NL) {
                int k = e / NW, i = e % NW;
                double s = rvec[e];
                if (k < N && i < NU) s += cnb[k * NU + i];
                if (k >= k0)
                    for (int r = 0; r < NH; ++r) s += cns[k * NH + r] * P.C[((size_t)k * NH + r) * NW + i];
                gt[e] = s;
            }
            L.sync();
            // ---------- backward Riccati sweep ----------
            if (pass == 0) {
                for (int e = lane; e < NX * NX; e += NL) Psm[e] = Ht[(size_t)N * NW * NW + (size_t)(NU + e / NX) * NW + NU + e % NX];
            }
            for (int e = lane; e < NX; e += NL) pv[e] = gt[(size_t)N * NW + NU + e];
            L.sync();
            for (int k = N - 1; k >= 0; --k) {
                const double* ABk = P.AB + (size_t)k * NX * NW;
                if (pass == 0) {
                    for (int e = lane; e < NX * NW; e += NL) ABb[(e / NW) * NWB + e % NW] = ABk[e];
                    for (int e = lane; e < NX; e += NL) ABb[e * NWB + NW] = rdyn[(k + 1) * NX + e];
                    L.sync();
                    for (int e = lane; e < NX * NWB; e += NL) {
                        int i = e / NWB, j = e % NWB;
                        double s = 0.0;
                        for (int l = 0; l < NX; ++l) s += Psm[i * NX + l] * ABb[l * NWB + j];
                        if (j == NW) { PbM[k * NX + i] = s; s += pv[i]; }
                        PAB[e] = s;
                    }
                    L.sync();
                    constexpr int TRI = NW * (NW + 1) / 2;
                    for (int e = lane; e < TRI + NW; e += NL) {
                        if (e < TRI) {
                            int i, j;
                            tri_decode(e, i, j);
                            double s = Ht[(size_t)k * NW * NW + i * NW + j];
                            for (int l = 0; l < NX; ++l) s += ABb[l * NWB + i] * PAB[l * NWB + j];
                            Msm[i * NW + j] = s;
                            Msm[j * NW + i] = s;
                        } else {
                            int i = e - TRI;
                            double s = gt[(size_t)k * NW + i];
                            for (int l = 0; l < NX; ++l) s += ABb[l * NWB + i] * PAB[l * NWB + NW];
                            rv[i] = s;
                        }
                    }
                    L.sync();
                    // Cholesky of Ruu, redundantly in every lane
                    double Lr[NU][NU];
                    bool ok = true;
                    for (int i = 0; i < NU; ++i)
                        for (int j = 0; j <= i; ++j) {
                            double s = Msm[i * NW + j];
                            for (int l = 0; l < j; ++l) s -= Lr[i][l] * Lr[j][l];
                            if (i == j) { ok = ok && (s > 0.0); Lr[i][i] = sqrt(s); }
                            else Lr[i][j] = s / Lr[j][j];
                        }
                    if (!ok) status = 4;
                    for (int c = lane; c < NX + 1; c += NL) {
                        double x[NU];
                        for (int i = 0; i < NU; ++i) {
                            double s = c < NX ? Msm[i * NW + NU + c] : rv[i];
                            for (int l = 0; l < i; ++l) s -= Lr[i][l] * x[l];
                            x[i] = s / Lr[i][i];
                        }
                        for (int i = NU - 1; i >= 0; --i) {
                            double s = x[i];
                            for (int l = i + 1; l < NU; ++l) s -= Lr[l][i] * x[l];
                            x[i] = s / Lr[i][i];
                        }
                        for (int i = 0; i < NU; ++i) {
                            if (c < NX) { Ksm[i * NX + c] = -x[i]; Kg[(size_t)k * NU * NX + i * NX + c] = -x[i]; }
                            else { kfs[i] = -x[i]; kffg[k * NU + i] = -x[i]; }
                        }
                    }
                    if (lane == 0)
                        for (int i = 0; i < NU; ++i)
                            for (int j = 0; j < NU; ++j) Lg[(size_t)k * NU * NU + i * NU + j] = j <= i ? Lr[i][j] : 0.0;
                    L.sync();
                    for (int e = lane; e < NX * NX + NX; e += NL) {
                        if (e < NX * NX) {
                            int i = e / NX, j = e % NX;
                            double s = Msm[(NU + i) * NW + NU + j];
                            for (int u = 0; u < NU; ++u) s += Msm[(NU + i) * NW + u] * Ksm[u * NX + j];
                            Psm[e] = s;
                        } else {
                            int i = e - NX * NX;
                            double s = rv[NU + i];
                            for (int u = 0; u < NU; ++u) s += Msm[(NU + i) * NW + u] * kfs[u];
                            pv[i] = s;
                        }
                    }
                    L.sync();
                    // keep P symmetric against round-off drift (Msm is free again: use it as the temporary)
                    for (int e = lane; e < NX * NX; e += NL) {
                        int i = e / NX, j = e % NX;
                        if (i > j) Msm[i * NW + j] = 0.5 * (Psm[i * NX + j] + Psm[j * NX + i]);
                    }
                    L.sync();
                    for (int e = lane; e < NX * NX; e += NL) {
                        int i = e / NX, j = e % NX;
                        if (i > j) { double a = Msm[i * NW + j]; Psm[i * NX + j] = a; Psm[j * NX + i] = a; }
                    }
                    L.sync();
                } else {
                    for (int e = lane; e < NX; e += NL) vv[e] = PbM[k * NX + e] + pv[e];
                    L.sync();
                    for (int i = lane; i < NW; i += NL) {
                        double s = gt[(size_t)k * NW + i];
                        for (int l = 0; l < NX; ++l) s += ABk[l * NW + i] * vv[l];
                        rv[i] = s;
                    }
                    L.sync();
                    const double* Lk = Lg + (size_t)k * NU * NU;
                    double x[NU];
                    for (int i = 0; i < NU; ++i) {
                        double s = rv[i];
                        for (int l = 0; l < i; ++l) s -= Lk[i * NU + l] * x[l];
                        x[i] = s / Lk[i * NU + i];
                    }
                    for (int i = NU - 1; i >= 0; --i) {
                        double s = x[i];
                        for (int l = i + 1; l < NU; ++l) s -= Lk[l * NU + i] * x[l];
                        x[i] = s / Lk[i * NU + i];
                    }
                    if (lane == 0)
                        for (int i = 0; i < NU; ++i) kffg[k * NU + i] = -x[i];
                    const double* Kk = Kg + (size_t)k * NU * NX;
                    for (int i = lane; i < NX; i += NL) {   // pv was last read two barriers ago
                        double s = rv[NU + i];
                        for (int u = 0; u < NU; ++u) s += Kk[u * NX + i] * rv[u];
                        pv[i] = s;
                    }
                    L.sync();
                }
            }
            // ---------- forward roll-out -> wp (= delta w) ----------
            for (int i = lane; i < NX; i += NL) wp[NU + i] = rdyn[i];
            L.sync();
            for (int k = 0; k < N; ++k) {
                const double* Kk = Kg + (size_t)k * NU * NX;
                double* wk = wp + (size_t)k * NW;
                for (int j = lane; j < NU; j += NL) {
                    double s = kffg[k * NU + j];
                    for (int i = 0; i < NX; ++i) s += Kk[j * NX + i] * wk[NU + i];
                    wk[j] = s;
                }
                L.sync();
                const double* ABk = P.AB + (size_t)k * NX * NW;
                for (int i = lane; i < NX; i += NL) {
                    double s = rdyn[(k + 1) * NX + i];
                    for (int j = 0; j < NW; ++j) s += ABk[i * NW + j] * wk[j];
                    wk[NW + NU + i] = s;
                }
                L.sync();
            }
            for (int j = lane; j < NU; j += NL) wp[(size_t)N * NW + j] = 0.0;
            L.sync();
            // ---------- row steps ----------
            double amax = 1e30, csum_aff = 0.0;
            for (int e = lane; e < nhr; e += NL) {
                int k = e / (2 * NU), a = e % (2 * NU);
                int j = a < NU ? a : a - NU;
                double sgn = a < NU ? 1.0 : -1.0;
                double t = tb[e], lam = lb[e];
                double gd = sgn * wp[k * NW + j];
                double r_p = rpb[e];
                double kap = (pass == 0 ? 0.0 : tau - pab[e]) - lam * t;
                double dt_ = gd - r_p;
                double dl = (kap - lam * dt_) / t;
                dtb[e] = dt_; dlb[e] = dl;
                amax = fmin(amax, fmin(step_to_boundary(t, dt_), step_to_boundary(lam, dl)));
            }
            for (int e = lane + k0 * 2 * NH; e < nsr; e += NL) {
                int k = e / (2 * NH), a = e % (2 * NH);
                int r = a < NH ? a : a - NH;
                double sgn = a < NH ? 1.0 : -1.0;
                const double* Cr = P.C + ((size_t)k * NH + r) * NW;
                double cd = 0.0;
                for (int i = 0; i < NW; ++i) cd += Cr[i] * wp[(size_t)k * NW + i];
                double gd = sgn * cd;
                double kap = (pass == 0 ? 0.0 : tau - pas[e]) - ls[e] * ts[e];
                double nu = (pass == 0 ? 0.0 : tau - pbs[e]) - ms[e] * ss[e];
                double Th = ls[e] / ts[e];
                double Dd = (P.Zs[r] * scale(k) + ms[e] / ss[e]) + Th;
                double rho = rhos[e];
                double ds = (rho - Th * gd) / Dd;
                double dt_ = gd + ds - rps[e];
                double dl = (kap - ls[e] * dt_) / ts[e];
                double dm = (nu - ms[e] * ds) / ss[e];
                dts[e] = dt_; dls[e] = dl; dss[e] = ds; dms[e] = dm;
                amax = fmin(amax, fmin(fmin(step_to_boundary(ts[e], dt_), step_to_boundary(ls[e], dl)),
                                       fmin(step_to_boundary(ss[e], ds), step_to_boundary(ms[e], dm))));
            }
            amax = L.min(amax);
            if (pass == 0) {
                const double aa = fmin(1.0, amax);
                for (int e = lane; e < nhr; e += NL) {
                    csum_aff += (lb[e] + aa * dlb[e]) * (tb[e] + aa * dtb[e]);
                    pab[e] = so * dlb[e] * dtb[e];
                }
                for (int e = lane + k0 * 2 * NH; e < nsr; e += NL) {
                    csum_aff += (ls[e] + aa * dls[e]) * (ts[e] + aa * dts[e]) + (ms[e] + aa * dms[e]) * (ss[e] + aa * dss[e]);
                    pas[e] = so * dls[e] * dts[e];
                    pbs[e] = so * dms[e] * dss[e];
                }
                const double mu_aff = L.sum(csum_aff) / ncomp;
                double sig = mu_aff / fmax(mubar, 1e-300);
                sig = fmin(fmax(sig * sig * sig, safe ? 0.2 : 0.0), 1.0);
                tau = fmax(sig * mubar, 0.1 * P.tol);
                L.sync();
            } else {
                // ---------- costate step (pip = delta pi), then the damped step ----------
                for (int i = lane; i < NX; i += NL) {
                    const double* Hk = Ht + (size_t)N * NW * NW + (size_t)(NU + i) * NW;
                    double s = gt[(size_t)N * NW + NU + i];
                    for (int j = NU; j < NW; ++j) s += Hk[j] * wp[(size_t)N * NW + j];
                    pip[(size_t)N * NX + i] = s;
                }
                L.sync();
                for (int k = N - 1; k >= 1; --k) {
                    const double* ABk = P.AB + (size_t)k * NX * NW;
                    for (int i = lane; i < NX; i += NL) {
                        const double* Hk = Ht + (size_t)k * NW * NW + (size_t)(NU + i) * NW;
                        double s = gt[(size_t)k * NW + NU + i];
                        for (int j = 0; j < NW; ++j) s += Hk[j] * wp[(size_t)k * NW + j];
                        for (int l = 0; l < NX; ++l) s += ABk[l * NW + NU + i] * pip[(size_t)(k + 1) * NX + l];
                        pip[(size_t)k * NX + i] = s;
                    }
                    L.sync();
                }
                const double al = fmin(1.0, 0.995 * amax);
#if defined(SDFNMPC_QP_DEBUG) && !defined(__CUDA_ARCH__)
                {
                    int wb = -1, ws = -1; double mb = 1e30, msv = 1e30; int which = 0;
                    for (int e = 0; e < nhr; ++e) { double a1 = step_to_boundary(tb[e], dtb[e]), a2 = step_to_boundary(lb[e], dlb[e]); if (fmin(a1,a2) < mb) { mb = fmin(a1,a2); wb = e; } }
                    for (int e = k0 * 2 * NH; e < nsr; ++e) { double a[4] = {step_to_boundary(ts[e], dts[e]), step_to_boundary(ls[e], dls[e]), step_to_boundary(ss[e], dss[e]), step_to_boundary(ms[e], dms[e])}; for (int q = 0; q < 4; ++q) if (a[q] < msv) { msv = a[q]; ws = e; which = q; } }
                    printf("   alpha %.3e tau %.3e mubar %.3e | hard row %d a=%.3e (t %.3e l %.3e) | soft row %d var %d a=%.3e (t %.3e l %.3e s %.3e m %.3e)\n", al, tau, mubar, wb, mb, tb[wb], lb[wb], ws, which, msv, ts[ws], ls[ws], ss[ws], ms[ws]);
                }
#endif
                for (int e = lane; e < (N + 1) * NW; e += NL) w[e] += al * wp[e];
                for (int e = lane + NX; e < (N + 1) * NX; e += NL) pi[e] += al * pip[e];
                for (int e = lane; e < nhr; e += NL) { tb[e] += al * dtb[e]; lb[e] += al * dlb[e]; }
                for (int e = lane + k0 * 2 * NH; e < nsr; e += NL) {
                    ts[e] += al * dts[e]; ls[e] += al * dls[e]; ss[e] += al * dss[e]; ms[e] += al * dms[e];
                }
                L.sync();
            }
        }
        if (status == 4) break;
    }

    // ---- full SQP step and outputs -----------------------------------------------------------
    if (status != 4) {
        for (int e = lane; e < (N + 1) * NX; e += NL) P.xbar[e] += w[(size_t)(e / NX) * NW + NU + e % NX];
        for (int e = lane; e < N * NU; e += NL) P.ubar[e] += w[(size_t)(e / NU) * NW + e % NU];
    }
    L.sync();
    for (int j = lane; j < NU; j += NL) P.u0[j] = P.ubar[j];
    if (P.lam_hard)
        for (int e = lane; e < nhr; e += NL) P.lam_hard[e] = lb[e];
    if (P.lam_soft)
        for (int e = lane; e < nsr; e += NL) P.lam_soft[e] = e >= k0 * 2 * NH ? ls[e] : 0.0;
    if (lane == 0) { *P.status = status; *P.iters = it; }
}

}  // namespace sdfnmpc
