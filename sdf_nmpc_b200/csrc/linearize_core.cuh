// Stage-wise linearisation of the sdf-nmpc OCP around the current SQP iterate.
//
// One call handles ONE tangent direction `dir` (column of [du; dx]) of ONE shooting
// node of ONE instance: it propagates a dual number through the ERK4 step, the NLS
// outputs and the FoV rows and writes column `dir` of [B|A], J and C.  This is the
// role of acados' per-node calls to the CasADi-generated expl_vde_forw,
// cost_y_fun_jac_ut_xt and constr_h_fun_jac_uxt_zt (sdf_nmpc/ocp.py:106-120).
#pragma once
#include "model_att.cuh"

namespace sdfnmpc {

struct LinIn {           // device pointers, whole batch
    double* xbar;        // [B][N+1][NX]   node 0 is overwritten with x0 (see linearize_dir)
    const double* ubar;  // [B][N][NU]
    const double* yref;  // [B][N][NY]
    const double* yrefN; // [B][NYN]
    const double* qd;    // [B][N+1][4]
    const double* pose;  // [B][12]
    const double* flag;  // [B]
    const double* dt;    // [N]
    int nyN;             // terminal outputs: 4 (att, att_tau) or 7 (acc)
    const double* x0;    // [B][NX] measured state: node 0 is linearised THERE (ocp.py:161 `solver.set(0, 'x', x0)` precedes
                         // the RTI step, so dx_0 = 0); null = linearise node 0 at xbar_0 (the QP then carries dx_0 = x0 - xbar_0)
    // Lifted terminal rows (term_lift.cuh): the OCP ends at node n_real < N, nodes n_real+1..N are auxiliary (their dynamics and
    // rows are written by term_lift_instance after the SDF evaluation).  The caller's arrays yref / qd keep the OCP's horizon
    // ([B][n_real][NY], [B][n_real+1][4]); xbar / ubar and every output are laid out for N.  -1: n_real = N.
    int n_real = -1;
    int stab = 0;        // flags.stability: terminal outputs flag * [p, q_e3] and flag * v.v (quad_rollpitchyawrate.py:52-55, gen_model.py:139-147)
};

struct LinOut {
    double* AB;    // [B][N][NX][NW]   columns ordered [u; x]; optional (null: only ABt is written)
    double* bdyn;  // [B][N][NX]       phi(xbar_k, ubar_k) - xbar_{k+1}
    double* J;     // [B][N+1][NY][NW]
    double* yres;  // [B][N+1][NY]     y - yref (terminal rows >= NYN are zero)
    double* C;     // [B][N+1][NH][NW] rows 0,1 = FoV; row 2 (sdf) is filled by the QP prologue
    double* hval;  // [B][N+1][NH]
    float* qpts;   // [B][N+1][3]      SDF query points Co_p_B (float32, as L4CasADi casts)
    double* ABt;   // [B][N][16][12]   optional: the same [A B | b] in the operand layout of the tensor-core QP sweeps
                   //                  (qp_mma.cuh): rows = 16 positions of [u; x; 1; pad], columns = 12 slots of [x; 1; pad]
};

// position of entry i of w~ = [u (4); x (10); 1; pad] in the tensor-core layout (qp_mma.cuh): x0..x7 -> 0..7, x8 / x9 / affine /
// pad -> 8 / 10 / 12 / 14, u0..u3 -> 9 / 11 / 13 / 15
SDF_HD constexpr int mma_wpos(int i) { return i < ATT_NU ? 9 + 2 * i : (i < ATT_NU + 8 ? i - ATT_NU : 8 + 2 * (i - ATT_NU - 8)); }

// ERK4 step of a dual number whose state is the measured / iterate state `xg` (values in memory) seeded with the unit tangent
// e_dx (dx < 0: no state tangent).  Same arithmetic as rk4_step<MODEL, Dual>; the state is re-read instead of kept.
// (Measured: spills 552 -> 352 B per thread with this and the running sum of rk4_step, 0.279 -> 0.273 ms per 4096 instances; keeping
// the running sum in shared memory on top does not reduce them further - the rest of the pressure is inside the model function.)
template <int MODEL>
SDF_HD void rk4_step_seeded(const ModelConst& c, const double* xg, int dx, const Dual* u, double dt, Dual* xn) {
    constexpr int NX_ = ModelDims<MODEL>::NX;
    Dual k[NX_], xs[NX_];
    AttInputTrig<Dual> trig;
    const AttInputTrig<Dual>* it = nullptr;
    if (MODEL == MODEL_ATT) { trig = att_input_trig(c, u); it = &trig; }
    auto X = [&](int i) { return Dual(xg[i], dx == i ? 1.0 : 0.0); };
    for (int i = 0; i < NX_; ++i) xs[i] = X(i);
    f_model<MODEL>(c, xs, u, k, it);
    for (int i = 0; i < NX_; ++i) { xn[i] = k[i]; xs[i] = X(i) + (0.5 * dt) * k[i]; }
    f_model<MODEL>(c, xs, u, k, it);
    for (int i = 0; i < NX_; ++i) { xn[i] = xn[i] + 2.0 * k[i]; xs[i] = X(i) + (0.5 * dt) * k[i]; }
    f_model<MODEL>(c, xs, u, k, it);
    for (int i = 0; i < NX_; ++i) { xn[i] = xn[i] + 2.0 * k[i]; xs[i] = X(i) + dt * k[i]; }
    f_model<MODEL>(c, xs, u, k, it);
    for (int i = 0; i < NX_; ++i) xn[i] = X(i) + (dt / 6.0) * (xn[i] + k[i]);
}

template <int MODEL>
SDF_HD void linearize_dir(const ModelConst& mc, int N, const LinIn& in, const LinOut& out, int b, int k, int dir) {
    using D_ = ModelDims<MODEL>;
    constexpr int NX = D_::NX, NU = D_::NU, NW = D_::NW, NY = D_::NY, NH = D_::NH;
    const int NR = in.n_real >= 0 ? in.n_real : N;
    const size_t node = (size_t)b * (N + 1) + k;
    if (k > NR) {        // auxiliary node: no cost; a harmless SDF query point
        double* J = out.J + node * NY * NW;
        for (int r = 0; r < NY; ++r) J[r * NW + dir] = 0.0;
        if (dir == 0) {
            for (int r = 0; r < NY; ++r) out.yres[node * NY + r] = 0.0;
            for (int c = 0; c < 3; ++c) out.qpts[node * 3 + c] = 0.f;
        }
        return;
    }
    double* xb = in.xbar + ((size_t)b * (N + 1) + k) * NX;
    const double* x = (k == 0 && in.x0) ? in.x0 + (size_t)b * NX : xb;
    // only the threads of node 0 read x0 and none of them reads xbar_0, so the direction-0 thread may store it
    if (k == 0 && in.x0 && dir == 0)
        for (int i = 0; i < NX; ++i) xb[i] = x[i];
#if defined(__CUDA_ARCH__)
    // The direction-0 thread of a node reads the node's reference row and the next node's state AFTER the RK4 step and the
    // outputs; cold, those loads were 25 % of the kernel's stall samples (the other lanes of the warp wait for them).  Ask for
    // the lines now.
    if (dir == 0) {
        const double* r_ = k < NR ? in.yref + ((size_t)b * NR + k) * NY : in.yrefN + (size_t)b * in.nyN;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(r_));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(r_ + (k < NR ? NY : in.nyN) - 1));
        if (k < NR) {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(xb + NX));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(xb + 2 * NX - 1));
        }
    }
#endif
    // The seeded state is rebuilt from memory wherever it is needed (the 10-14 values of a node sit in L1 for its 14-18 threads)
    // instead of being held in registers across the RK4 step: 20 fewer live doubles where the pressure is highest.
    Dual ud[NU];
    auto seed_x = [&](Dual* xd_) {
        for (int i = 0; i < NX; ++i) xd_[i] = Dual(x[i], (dir - NU) == i ? 1.0 : 0.0);
    };
    if (k < NR) {
        const double* u = in.ubar + ((size_t)b * N + k) * NU;
        for (int j = 0; j < NU; ++j) ud[j] = Dual(u[j], dir == j ? 1.0 : 0.0);
    } else {
        for (int j = 0; j < NU; ++j) ud[j] = Dual(0.0, 0.0);
    }
    if (k < NR) {
        Dual xn[NX];
        rk4_step_seeded<MODEL>(mc, x, dir - NU, ud, in.dt[k], xn);
        if (out.AB) {       // null when the solver reads the operand layout `ABt` only
            double* AB = out.AB + ((size_t)b * N + k) * NX * NW;
            for (int i = 0; i < NX; ++i) AB[i * NW + dir] = xn[i].d;
        }
        double* T = (out.ABt && NX == ATT_NX) ? out.ABt + ((size_t)b * N + k) * 192 : nullptr;     // 10-state models only
        if (T) {
            double* row = T + mma_wpos(dir) * 12;
            for (int i = 0; i < NX; ++i) row[i] = xn[i].d;
            row[10] = 0.0; row[11] = 0.0;                    // affine slot, pad slot
        }
        if (dir == 0) {
            const double* x1 = xb + NX;
            double* bd = out.bdyn + ((size_t)b * N + k) * NX;
            for (int i = 0; i < NX; ++i) bd[i] = xn[i].v - x1[i];
            if (T) {
                // affine column (position 12): the dynamics defect, and the 1 that carries the affine entry along; pad row zero
                for (int i = 0; i < NX; ++i) T[12 * 12 + i] = xn[i].v - x1[i];
                T[12 * 12 + 10] = 1.0; T[12 * 12 + 11] = 0.0;
                for (int i = 0; i < 12; ++i) T[14 * 12 + i] = 0.0;
            }
        }
    }
    Dual xd[NX];
    seed_x(xd);
    {
        Dual y[NY];
        y_model<MODEL>(mc, xd, ud, in.qd + ((size_t)b * (NR + 1) + k) * 4, y);
        if (k == NR && in.stab) {
            const double f = in.flag[b];
            for (int r = 0; r < 4; ++r) y[r] = f * y[r];
            y[4] = f * (xd[7] * xd[7] + xd[8] * xd[8] + xd[9] * xd[9]);
        }
        double* J = out.J + node * NY * NW;
        for (int r = 0; r < NY; ++r) J[r * NW + dir] = y[r].d;
        if (dir == 0) {
            double* yr = out.yres + node * NY;
            if (k < NR) {
                const double* ref = in.yref + ((size_t)b * NR + k) * NY;
                for (int r = 0; r < NY; ++r) yr[r] = y[r].v - ref[r];
            } else {
                const double* ref = in.yrefN + (size_t)b * in.nyN;
                for (int r = 0; r < NY; ++r) yr[r] = r < in.nyN ? y[r].v - ref[r] : 0.0;
            }
        }
    }
    {
        const double* pose = in.pose + (size_t)b * 12;
        Dual h[2];
        hfov_att(mc, pose, in.flag[b], xd, h);
        double* C = out.C + node * NH * NW;
        C[0 * NW + dir] = h[0].d;
        C[1 * NW + dir] = h[1].d;
        C[2 * NW + dir] = 0.0;
        if (dir == 0) {
            out.hval[node * NH + 0] = h[0].v;
            out.hval[node * NH + 1] = h[1].v;
            double c0[3];
            camera_point(pose, x, c0);
            float* q = out.qpts + node * 3;
            q[0] = (float)c0[0];
            q[1] = (float)c0[1];
            q[2] = (float)c0[2];
        }
    }
}

}  // namespace sdfnmpc
