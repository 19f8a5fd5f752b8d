// Terminal ingredients of flags.recursive_feasibility / flags.stability (sdf_nmpc/gen_model.py:73-149), model 'att'.
//
// The reference adds, at the terminal node x_N = [p, q, v]:
//   rec_feas = sdf_flag(p) - flag * poly(v)            in [size.xy, max_df]   (soft with weights.slack_brake, else hard)
//   hfov, vfov at the braking end point  p + flag * poly(v) v / sqrt(v.v + 1e-4)          (hard, gen_model.py:105-118)
//   |v_i| <= limits.v_i                                (hard, stability only; cost_const_helpers.py:93-102)
// with poly the 3-variate braking-distance polynomial (utils/math.py:294-321).  These rows depend on position AND velocity,
// while the QP kernels carry the h rows of a node as 3-column blocks on the position (qp_core.cuh, "Assumption encoded in
// HC0/HCN").  They are therefore LIFTED: behind the terminal node n_real the horizon is extended by auxiliary nodes whose
// "position" entries ARE the row functions,
//   x_{n_real+1} = Phi_1(x_N) = [hfov_E, vfov_E, rec_feas, (1,0,0,0), v],      x_{n_real+2} = Phi_2(x_{n_real+1}) = [v, (1,0,0,0), 0]
// with "dynamics" A = dPhi/dx, B = 0, zero defect (the auxiliary states are re-derived from x_N at every linearisation), no cost
// and no regularisation of their state, so that row r of an auxiliary node is the unit row e_r on the position block with the
// bounds of the lifted row.  On the QP level this is the same problem: dx_{aux} = A dx_N turns  lh <= h + e_r.dx_{aux} <= uh
// into  lh <= h + grad h . dx_N <= uh.  Hard rows use the exact L1 penalty (slack weight `z_hard` far above any multiplier).
#pragma once
#include "model_att.cuh"
#include "linearize_core.cuh"   // mma_wpos

namespace sdfnmpc {

constexpr int TERM_MAX_COEFFS = 56;        // degree <= 5

struct TermConst {
    int n_lift;                            // 0 off, 1 recursive feasibility, 2 + stability
    int degree;
    double coeffs[TERM_MAX_COEFFS];        // ordered by total degree, then a, then b (c = td - a - b): utils/math.py:306-312
};

template <class T>
SDF_HD T term_poly3(const TermConst& tc, const T* v) {
    T px[6], py[6], pz[6];
    px[0] = T(1.0); py[0] = T(1.0); pz[0] = T(1.0);
    for (int i = 1; i <= tc.degree; ++i) { px[i] = px[i - 1] * v[0]; py[i] = py[i - 1] * v[1]; pz[i] = pz[i - 1] * v[2]; }
    T s = T(0.0);
    int i = 0;
    for (int td = 0; td <= tc.degree; ++td)
        for (int a = 0; a <= td; ++a)
            for (int b = 0; b <= td - a; ++b) s = s + tc.coeffs[i++] * (px[a] * py[b] * pz[td - a - b]);
    return s;
}

// Direction d (state entry d, 0 <= d < NX) of one instance: column NU + d of the auxiliary stages n_real..N-1 of [B A] (and, for
// d < NU, the zero input column d), the matching rows of the operand layout `ABt` when given (LinOut::ABt: [16 positions][12 slots]);
// direction 0 also writes what exists once per instance - the zero defects, the auxiliary states, the unit rows of the auxiliary
// nodes with the row values, the affine / pad rows of `ABt`.  The directions of an instance write disjoint entries and read only
// node n_real of xbar, so they may run as separate threads (nmpc_term_lift_kernel) or as a loop (term_lift_instance).
// Pointers are offset to the instance; AB is [N][NX][NW] with columns [u; x].  sdf_d / sdf_g: network value and gradient (camera
// frame) at node n_real.
SDF_HD void term_lift_dir(const ModelConst& mc, const TermConst& tc, int N, int NR, const double* pose, double flag, double max_df,
                          float sdf_d, const float* sdf_g, double* xbar, double* AB, double* bdyn, double* C, double* hval,
                          double* ABt, int d) {
    constexpr int NX = ATT_NX, NU = ATT_NU, NW = ATT_NW, NH = ATT_NH;
    const double* x = xbar + (size_t)NR * NX;
    const double* R = pose + 3;
    const bool two = tc.n_lift >= 2 && NR + 2 <= N;
    // ---- stage n_real: x_{n_real+1} = Phi_1(x_N) = [hfov_E, vfov_E, rec_feas, (1,0,0,0), v] ----
    Dual xd[NX];
    for (int i = 0; i < NX; ++i) xd[i] = Dual(x[i], i == d ? 1.0 : 0.0);
    const Dual* v = xd + 7;
    const Dual dist = flag * term_poly3(tc, v);
    const Dual inv_n = 1.0 / dsqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + 1e-4);
    Dual xe[NX];
    for (int i = 0; i < NX; ++i) xe[i] = xd[i];
    for (int i = 0; i < 3; ++i) xe[i] = xd[i] + dist * (v[i] * inv_n);
    Dual h[2];
    hfov_att(mc, pose, flag, xe, h);
    // rec_feas: flag * sdf(Co_p_B) + (1 - flag) * max_df - flag * poly(v); d sdf / d p = R g (camera_point: c = R^T (p - p_Co))
    double rec_d = -dist.d;
    if (d < 3) rec_d += flag * (R[d * 3 + 0] * (double)sdf_g[0] + R[d * 3 + 1] * (double)sdf_g[1] + R[d * 3 + 2] * (double)sdf_g[2]);
    double col1[NX], col2[NX];
    for (int i = 0; i < NX; ++i) { col1[i] = 0.0; col2[i] = 0.0; }
    col1[0] = h[0].d; col1[1] = h[1].d; col1[2] = rec_d;
    if (d >= 7) { col1[d] = 1.0; col2[d - 7] = 1.0; }      // v carried along; Phi_2 = [v, (1,0,0,0), 0]
    auto put_col = [&](int k, int dir, const double* col) {
        double* A = AB + (size_t)k * NX * NW;
        for (int i = 0; i < NX; ++i) A[i * NW + dir] = col ? col[i] : 0.0;
        if (ABt) {
            double* row = ABt + (size_t)k * 192 + mma_wpos(dir) * 12;
            for (int i = 0; i < NX; ++i) row[i] = col ? col[i] : 0.0;
            row[10] = 0.0; row[11] = 0.0;
        }
    };
    put_col(NR, NU + d, col1);
    if (d < NU) put_col(NR, d, nullptr);
    if (two) {
        put_col(NR + 1, NU + d, col2);
        if (d < NU) put_col(NR + 1, d, nullptr);
    }
    if (d != 0) return;
    // ---- once per instance ----
    const double val[3] = {h[0].v, h[1].v, flag * (double)sdf_d + (1.0 - flag) * max_df - dist.v};
    auto once = [&](int k, const double* hv, const double* xa) {      // stage k -> node k + 1
        for (int i = 0; i < NX; ++i) bdyn[(size_t)k * NX + i] = 0.0;
        double* xn = xbar + (size_t)(k + 1) * NX;
        for (int i = 0; i < NX; ++i) xn[i] = xa[i];
        double* Ck = C + (size_t)(k + 1) * NH * NW;
        for (int i = 0; i < NH * NW; ++i) Ck[i] = 0.0;
        for (int r = 0; r < NH; ++r) { Ck[r * NW + NU + r] = 1.0; hval[(k + 1) * NH + r] = hv[r]; }
        if (ABt) {      // affine position 12 = zero defect and the 1 that carries the affine entry, pad position 14 zero
            double* T = ABt + (size_t)k * 192;
            for (int i = 0; i < 12; ++i) { T[12 * 12 + i] = i == 10 ? 1.0 : 0.0; T[14 * 12 + i] = 0.0; }
        }
    };
    const double x1[NX] = {val[0], val[1], val[2], 1.0, 0.0, 0.0, 0.0, x[7], x[8], x[9]};
    once(NR, val, x1);
    if (two) {
        const double vv[3] = {x[7], x[8], x[9]};
        const double x2[NX] = {vv[0], vv[1], vv[2], 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        once(NR + 1, vv, x2);
    }
}

// One instance, all directions in turn (host emulation; the device runs the directions as threads).
SDF_HD void term_lift_instance(const ModelConst& mc, const TermConst& tc, int N, int NR, const double* pose, double flag, double max_df,
                               float sdf_d, const float* sdf_g, double* xbar, double* AB, double* bdyn, double* C, double* hval,
                               double* ABt = nullptr) {
    // direction 0 last: it overwrites x_{n_real+1..}, which no direction reads, but keeps the order obvious
    for (int d = ATT_NX - 1; d >= 0; --d) term_lift_dir(mc, tc, N, NR, pose, flag, max_df, sdf_d, sdf_g, xbar, AB, bdyn, C, hval, ABt, d);
}

}  // namespace sdfnmpc
