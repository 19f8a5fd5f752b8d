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

// One instance: stages n_real..N-1 of [B A], their (zero) defects, the auxiliary states, and the rows of the auxiliary nodes.
// Pointers are offset to the instance; AB is [N][NX][NW] with columns [u; x].  sdf_d / sdf_g: network value and gradient (camera
// frame) at node n_real.
SDF_HD void term_lift_instance(const ModelConst& mc, const TermConst& tc, int N, int NR, const double* pose, double flag, double max_df,
                               float sdf_d, const float* sdf_g, double* xbar, double* AB, double* bdyn, double* C, double* hval,
                               double* ABt = nullptr) {
    constexpr int NX = ATT_NX, NU = ATT_NU, NW = ATT_NW, NH = ATT_NH;
    const double* x = xbar + (size_t)NR * NX;
    double* A1 = AB + (size_t)NR * NX * NW;
    for (int i = 0; i < NX * NW; ++i) A1[i] = 0.0;
    for (int i = 0; i < NX; ++i) bdyn[(size_t)NR * NX + i] = 0.0;
    double val[3] = {0.0, 0.0, 0.0};
    const double* R = pose + 3;
    for (int d = 0; d < NX; ++d) {
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
        A1[0 * NW + NU + d] = h[0].d;
        A1[1 * NW + NU + d] = h[1].d;
        A1[2 * NW + NU + d] = rec_d;
        if (d >= 7) A1[d * NW + NU + d] = 1.0;
        if (d == 0) {
            val[0] = h[0].v;
            val[1] = h[1].v;
            val[2] = flag * (double)sdf_d + (1.0 - flag) * max_df - dist.v;
        }
    }
    auto unit_rows = [&](int k, const double* hv) {
        double* Ck = C + (size_t)k * NH * NW;
        for (int i = 0; i < NH * NW; ++i) Ck[i] = 0.0;
        for (int r = 0; r < NH; ++r) { Ck[r * NW + NU + r] = 1.0; hval[k * NH + r] = hv[r]; }
    };
    double* x1 = xbar + (size_t)(NR + 1) * NX;
    x1[0] = val[0]; x1[1] = val[1]; x1[2] = val[2];
    x1[3] = 1.0; x1[4] = 0.0; x1[5] = 0.0; x1[6] = 0.0;
    x1[7] = x[7]; x1[8] = x[8]; x1[9] = x[9];
    unit_rows(NR + 1, val);
    if (tc.n_lift >= 2 && NR + 2 <= N) {
        double* A2 = AB + (size_t)(NR + 1) * NX * NW;
        for (int i = 0; i < NX * NW; ++i) A2[i] = 0.0;
        for (int i = 0; i < 3; ++i) A2[i * NW + NU + 7 + i] = 1.0;
        for (int i = 0; i < NX; ++i) bdyn[(size_t)(NR + 1) * NX + i] = 0.0;
        double* x2 = xbar + (size_t)(NR + 2) * NX;
        const double vv[3] = {x[7], x[8], x[9]};
        x2[0] = vv[0]; x2[1] = vv[1]; x2[2] = vv[2];
        x2[3] = 1.0; x2[4] = 0.0; x2[5] = 0.0; x2[6] = 0.0;
        x2[7] = 0.0; x2[8] = 0.0; x2[9] = 0.0;
        unit_rows(NR + 2, vv);
    }
    // the same stages in the operand layout of the tensor-core sweeps (LinOut::ABt: [16 positions][12 slots], affine position 12
    // = zero defect and the 1 that carries the affine entry, pad position 14 zero)
    if (ABt) {
        for (int k = NR; k < N && k < NR + tc.n_lift; ++k) {
            const double* A = AB + (size_t)k * NX * NW;
            double* T = ABt + (size_t)k * 192;
            for (int dir = 0; dir < NW; ++dir) {
                double* row = T + mma_wpos(dir) * 12;
                for (int i = 0; i < NX; ++i) row[i] = A[i * NW + dir];
                row[10] = 0.0; row[11] = 0.0;
            }
            for (int i = 0; i < 12; ++i) { T[12 * 12 + i] = i == 10 ? 1.0 : 0.0; T[14 * 12 + i] = 0.0; }
        }
    }
}

}  // namespace sdfnmpc
