// Quadrotor model 'att' (thrust / roll / pitch / yaw-rate inputs), templated on the
// scalar type so that it runs on double (plant, values) and on Dual (derivatives).
//
// Replaces the CasADi graphs of
//   sdf_nmpc/model/quad_rollpitchyawrate.py:20-42  (dynamics)      -> f_att
//   sdf_nmpc/model/quad_rollpitchyawrate.py:49-55  (NLS outputs)   -> y_att
//   sdf_nmpc/model/cost_const_helpers.py:56-65     (FoV rows)      -> hfov_att
//   sdf_nmpc/gen_model.py:46-51                    (SDF query)     -> camera_point
// and acados' sim_erk (ERK4, one step; ocp.py:106) -> rk4_step.
// Closed forms: SURVEY.md Appendix A.
#pragma once
#include "dual.cuh"

namespace sdfnmpc {

constexpr int ATT_NX = 10;
constexpr int ATT_NU = 4;
constexpr int ATT_NW = 14;   // [u; x]
constexpr int ATT_NY = 11;
constexpr int ATT_NYN = 4;
constexpr int ATT_NH = 3;    // hfov, vfov, sdf  (gen_model.py:42,68)

enum QuadModel { MODEL_ATT = 0, MODEL_ATT_TAU = 1, MODEL_ACC = 2, MODEL_PROPS = 3 };   // mpc.model: att / att_tau / acc / props (gen_model.py:15-20)

// Stage dimensions by model.  'props' (sdf_nmpc/model/quad_props.py:12-15) has 13 states [p, q, v, w], 4 rotor-speed inputs,
// 16 / 12 outputs; on the device its state carries ONE PAD ENTRY (index 13: zero dynamics, not in any output or constraint, so
// its step is exactly zero) because the QP's staging copies move 16-byte units and need even row lengths.
template <int MODEL>
struct ModelDims {
    static constexpr int NX = ATT_NX, NU = ATT_NU, NW = ATT_NW, NY = ATT_NY, NH = ATT_NH;
};
template <>
struct ModelDims<MODEL_PROPS> {
    static constexpr int NX = 14, NU = 4, NW = 18, NY = 16, NH = 3;
};
constexpr int PROPS_NX_REF = 13, PROPS_NYN = 12;

struct ModelConst {
    double g, gamma, roll, pitch, wz, mass;
    double cam_off[3];    // B_R_C^T B_p_C + (fov_const_offset, 0, 0)
    double ax, ay, az;    // acceleration limits (model 'acc', default.yaml:101-103)
    double tau_roll, tau_pitch;   // first-order attitude lag (model 'att_tau', quad_rollpitchyawrate_tau.py:19-20)
    // model 'props' (quad_props.py:18-38): rotor speed scale, allocation matrices with cf folded in - Gf / mass and J^-1 Gt,
    // row-major [3][4] - and the diagonal inertia
    double wp_max, Gf_m[12], JGt[12], inertia[3];
};

// Input-only trigonometry of model 'att': the inputs are constant over the RK4 stages of a node, so cos / sin of the roll
// and pitch set-points are evaluated once per node instead of once per stage.
template <class T>
struct AttInputTrig {
    T cr, sr, cp, sp;
    bool valid;
};
template <class T>
SDF_HD AttInputTrig<T> att_input_trig(const ModelConst& c, const T* u) {
    AttInputTrig<T> t;
    const T r = c.roll * u[1], pt = c.pitch * u[2];
    t.cr = dcos(r); t.sr = dsin(r); t.cp = dcos(pt); t.sp = dsin(pt);
    t.valid = true;
    return t;
}

template <class T>
SDF_HD void f_att(const ModelConst& c, const T* x, const T* u, T* dx, const AttInputTrig<T>* it = nullptr) {
    T q0 = x[3], q1 = x[4], q2 = x[5], q3 = x[6];
    T inv_n = 1.0 / dsqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
    T h0 = q0 * inv_n, h1 = q1 * inv_n, h2 = q2 * inv_n, h3 = q3 * inv_n;
    // yaw-only rotation: theta_z = atan2(q3, q0), W_R_V = Rz(2 theta_z).  With rho^2 = q0^2 + q3^2 the double-angle
    // identities give cos(2 theta_z) = (q0^2 - q3^2) / rho^2 and sin(2 theta_z) = 2 q0 q3 / rho^2 without atan2 / sin / cos.
    T irho2 = 1.0 / (q0 * q0 + q3 * q3);
    T cps = (q0 * q0 - q3 * q3) * irho2, sps = (2.0 * (q0 * q3)) * irho2;
    T gam = c.gamma * u[0];
    T wz = c.wz * u[3];
    T cr, sr, cp, sp;
    if (it) { cr = it->cr; sr = it->sr; cp = it->cp; sp = it->sp; }
    else { T r = c.roll * u[1], pt = c.pitch * u[2]; cr = dcos(r); sr = dsin(r); cp = dcos(pt); sp = dsin(pt); }
    dx[0] = x[7];
    dx[1] = x[8];
    dx[2] = x[9];
    T hw = 0.5 * wz;
    dx[3] = -(hw * h3);
    dx[4] = hw * h2;
    dx[5] = -(hw * h1);
    dx[6] = hw * h0;
    dx[7] = gam * (cps * cr * sp + sps * sr);
    dx[8] = gam * (sps * cr * sp - cps * sr);
    dx[9] = gam * cr * cp - c.g;
}

// Model 'att_tau' (sdf_nmpc/model/quad_rollpitchyawrate_tau.py:22-42): full-quaternion attitude, roll/pitch follow
// their set-points with a first-order lag through the Euler-rate map of utils/math.py:209-226 (restated literally,
// including its second row [0, cos(roll), -sin(pitch)]).
template <class T>
SDF_HD void f_att_tau(const ModelConst& c, const T* x, const T* u, T* dx) {
    T q0 = x[3], q1 = x[4], q2 = x[5], q3 = x[6];
    T inv_n = 1.0 / dsqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
    q0 = q0 * inv_n; q1 = q1 * inv_n; q2 = q2 * inv_n; q3 = q3 * inv_n;
    T roll = datan2(2.0 * (q0 * q1 + q2 * q3), 1.0 - 2.0 * (q1 * q1 + q2 * q2));     // quat2euler, utils/math.py:57-70
    T pitch = dasin(2.0 * (q0 * q2 - q3 * q1));
    T gam = c.gamma * u[0];
    T wz = c.wz * u[3];
    T dot_roll = (c.roll * u[1] - roll) / c.tau_roll;
    T dot_pitch = (c.pitch * u[2] - pitch) / c.tau_pitch;
    T sr = dsin(roll), cr = dcos(roll), sp = dsin(pitch), cp = dcos(pitch);
    T w0 = dot_roll + (sp * sr / cp) * dot_pitch;
    T w1 = cr * dot_pitch;
    dx[0] = x[7]; dx[1] = x[8]; dx[2] = x[9];
    dx[3] = 0.5 * (-(q1 * w0) - q2 * w1 - q3 * wz);
    dx[4] = 0.5 * (q0 * w0 + q2 * wz - q3 * w1);
    dx[5] = 0.5 * (q0 * w1 - q1 * wz + q3 * w0);
    dx[6] = 0.5 * (q0 * wz + q1 * w1 - q2 * w0);
    dx[7] = gam * (2.0 * (q1 * q3 + q0 * q2));                                       // third column of quat2rot
    dx[8] = gam * (2.0 * (q2 * q3 - q0 * q1));
    dx[9] = gam * (q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3) - c.g;
}

// Model 'acc' (sdf_nmpc/model/quad_acc.py:20-33): world-frame acceleration inputs and yaw rate.
template <class T>
SDF_HD void f_acc(const ModelConst& c, const T* x, const T* u, T* dx) {
    T q0 = x[3], q1 = x[4], q2 = x[5], q3 = x[6];
    T inv_n = 1.0 / dsqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
    T hw = 0.5 * (c.wz * u[3]);
    dx[0] = x[7]; dx[1] = x[8]; dx[2] = x[9];
    dx[3] = -(hw * (q3 * inv_n));
    dx[4] = hw * (q2 * inv_n);
    dx[5] = -(hw * (q1 * inv_n));
    dx[6] = hw * (q0 * inv_n);
    dx[7] = c.ax * u[0];
    dx[8] = c.ay * u[1];
    dx[9] = c.az * u[2];
}

// Model 'props' (sdf_nmpc/model/quad_props.py:29-50): rigid body driven by four rotor speeds wp = u * limits.wp through the
// allocation matrices of a generically tilted multi-rotor (utils/math.py:385-396), quaternion normalised first (line 30).
template <class T>
SDF_HD void f_props(const ModelConst& c, const T* x, const T* u, T* dx) {
    T q0 = x[3], q1 = x[4], q2 = x[5], q3 = x[6];
    T inv_n = 1.0 / dsqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
    q0 = q0 * inv_n; q1 = q1 * inv_n; q2 = q2 * inv_n; q3 = q3 * inv_n;
    const T w0 = x[10], w1 = x[11], w2 = x[12];
    T s[4];
    for (int j = 0; j < 4; ++j) { T wp = c.wp_max * u[j]; s[j] = wp * wp; }
    T F[3], M[3];
    for (int i = 0; i < 3; ++i) {
        F[i] = c.Gf_m[i * 4 + 0] * s[0] + c.Gf_m[i * 4 + 1] * s[1] + c.Gf_m[i * 4 + 2] * s[2] + c.Gf_m[i * 4 + 3] * s[3];
        M[i] = c.JGt[i * 4 + 0] * s[0] + c.JGt[i * 4 + 1] * s[1] + c.JGt[i * 4 + 2] * s[2] + c.JGt[i * 4 + 3] * s[3];
    }
    dx[0] = x[7]; dx[1] = x[8]; dx[2] = x[9];
    dx[3] = 0.5 * (-(q1 * w0) - q2 * w1 - q3 * w2);                                  // hamilton_prod(q, (0, w)) / 2
    dx[4] = 0.5 * (q0 * w0 + q2 * w2 - q3 * w1);
    dx[5] = 0.5 * (q0 * w1 - q1 * w2 + q3 * w0);
    dx[6] = 0.5 * (q0 * w2 + q1 * w1 - q2 * w0);
    // W_R_B (quat2rot, utils/math.py:7-23) times the body-frame specific force
    dx[7] = (q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3) * F[0] + (2.0 * (q1 * q2 - q0 * q3)) * F[1] + (2.0 * (q1 * q3 + q0 * q2)) * F[2];
    dx[8] = (2.0 * (q1 * q2 + q0 * q3)) * F[0] + (q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3) * F[1] + (2.0 * (q2 * q3 - q0 * q1)) * F[2];
    dx[9] = (2.0 * (q1 * q3 - q0 * q2)) * F[0] + (2.0 * (q2 * q3 + q0 * q1)) * F[1] + (q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3) * F[2] - c.g;
    // J^-1 (Gt wp^2 - w x J w), J diagonal
    const double J0 = c.inertia[0], J1 = c.inertia[1], J2 = c.inertia[2];
    dx[10] = M[0] - ((J2 - J1) / J0) * (w1 * w2);
    dx[11] = M[1] - ((J0 - J2) / J1) * (w2 * w0);
    dx[12] = M[2] - ((J1 - J0) / J2) * (w0 * w1);
    dx[13] = T(0.0);                                                                 // pad state
}

template <int MODEL, class T>
SDF_HD void f_model(const ModelConst& c, const T* x, const T* u, T* dx, const AttInputTrig<T>* it = nullptr) {
    if (MODEL == MODEL_ATT) f_att(c, x, u, dx, it);
    else if (MODEL == MODEL_ATT_TAU) f_att_tau(c, x, u, dx);
    else if (MODEL == MODEL_PROPS) f_props(c, x, u, dx);
    else f_acc(c, x, u, dx);
}

template <int MODEL, class T>
SDF_HD void rk4_step(const ModelConst& c, const T* x, const T* u, double dt, T* xn) {
    constexpr int NX_ = ModelDims<MODEL>::NX;
    // The weighted sum of the stage derivatives is accumulated as the stages come ((k1 + 2 k2) + 2 k3) + k4, the order of the
    // plain formula), so only one stage derivative is live at a time: with all four kept to the end the dual-number
    // instantiation holds 100 doubles and spills (552 B per thread, more than the L1 holds for 16 resident warps).
    T k[NX_], xs[NX_];
    AttInputTrig<T> trig;
    const AttInputTrig<T>* it = nullptr;
    if (MODEL == MODEL_ATT) { trig = att_input_trig(c, u); it = &trig; }
    f_model<MODEL>(c, x, u, k, it);
    for (int i = 0; i < NX_; ++i) { xn[i] = k[i]; xs[i] = x[i] + (0.5 * dt) * k[i]; }
    f_model<MODEL>(c, xs, u, k, it);
    for (int i = 0; i < NX_; ++i) { xn[i] = xn[i] + 2.0 * k[i]; xs[i] = x[i] + (0.5 * dt) * k[i]; }
    f_model<MODEL>(c, xs, u, k, it);
    for (int i = 0; i < NX_; ++i) { xn[i] = xn[i] + 2.0 * k[i]; xs[i] = x[i] + dt * k[i]; }
    f_model<MODEL>(c, xs, u, k, it);
    for (int i = 0; i < NX_; ++i) xn[i] = x[i] + (dt / 6.0) * (xn[i] + k[i]);
}

// y = [p, q_e[3], v, roll, pitch, wz, W_a[2]];  q_e = q_d (x) conj(q/|q|)
template <class T>
SDF_HD void y_att(const ModelConst& c, const T* x, const T* u, const double* qd, T* y) {
    T q0 = x[3], q1 = x[4], q2 = x[5], q3 = x[6];
    T inv_n = 1.0 / dsqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
    T h0 = q0 * inv_n, h1 = q1 * inv_n, h2 = q2 * inv_n, h3 = q3 * inv_n;
    T r = c.roll * u[1], pt = c.pitch * u[2];
    y[0] = x[0];
    y[1] = x[1];
    y[2] = x[2];
    y[3] = qd[2] * h1 + qd[3] * h0 - qd[0] * h3 - qd[1] * h2;
    y[4] = x[7];
    y[5] = x[8];
    y[6] = x[9];
    y[7] = r;
    y[8] = pt;
    y[9] = c.wz * u[3];
    y[10] = (c.gamma * u[0]) * dcos(r) * dcos(pt) - c.g;
}

// att_tau: y = [p, q_e[3], v, roll_des, pitch_des, wz, W_a[2]] with the full-quaternion W_a (quad_rollpitchyawrate_tau.py:49-52)
// acc:     y = [p, q_e[3], v, W_a(3), wz], yN = first 7                               (quad_acc.py:40-45)
// props:   y = [p, roll, pitch, q_e[3], v, w, wp(4)], yN = first 12                    (quad_props.py:56-59)
template <int MODEL, class T>
SDF_HD void y_model(const ModelConst& c, const T* x, const T* u, const double* qd, T* y) {
    if (MODEL == MODEL_ATT) { y_att(c, x, u, qd, y); return; }
    T q0 = x[3], q1 = x[4], q2 = x[5], q3 = x[6];
    T inv_n = 1.0 / dsqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
    T h0 = q0 * inv_n, h1 = q1 * inv_n, h2 = q2 * inv_n, h3 = q3 * inv_n;
    if (MODEL == MODEL_PROPS) {
        y[0] = x[0]; y[1] = x[1]; y[2] = x[2];
        y[3] = datan2(2.0 * (h0 * h1 + h2 * h3), 1.0 - 2.0 * (h1 * h1 + h2 * h2));     // quat2euler, utils/math.py:57-70
        y[4] = dasin(2.0 * (h0 * h2 - h3 * h1));
        y[5] = qd[2] * h1 + qd[3] * h0 - qd[0] * h3 - qd[1] * h2;
        for (int i = 0; i < 6; ++i) y[6 + i] = x[7 + i];
        for (int j = 0; j < 4; ++j) y[12 + j] = c.wp_max * u[j];
        return;
    }
    y[0] = x[0]; y[1] = x[1]; y[2] = x[2];
    y[3] = qd[2] * h1 + qd[3] * h0 - qd[0] * h3 - qd[1] * h2;
    y[4] = x[7]; y[5] = x[8]; y[6] = x[9];
    if (MODEL == MODEL_ATT_TAU) {
        y[7] = c.roll * u[1];
        y[8] = c.pitch * u[2];
        y[9] = c.wz * u[3];
        y[10] = (c.gamma * u[0]) * (h0 * h0 - h1 * h1 - h2 * h2 + h3 * h3) - c.g;
    } else {
        y[7] = c.ax * u[0];
        y[8] = c.ay * u[1];
        y[9] = c.az * u[2];
        y[10] = c.wz * u[3];
    }
}

// Co_p_B = W_R_Co^T (p - W_p_Co); pose = [W_p_Co(3), W_R_Co row-major(9)]  (controller.py:50-54)
template <class T>
SDF_HD void camera_point(const double* pose, const T* x, T* c0) {
    T d0 = x[0] - pose[0], d1 = x[1] - pose[1], d2 = x[2] - pose[2];
    const double* R = pose + 3;
    c0[0] = R[0] * d0 + R[3] * d1 + R[6] * d2;
    c0[1] = R[1] * d0 + R[4] * d1 + R[7] * d2;
    c0[2] = R[2] * d0 + R[5] * d1 + R[8] * d2;
}

// h1 = flag*atan2(c_y, c_x), h2 = flag*atan2(c_z, |c_xy|), c = Co_p_B + cam_off
template <class T>
SDF_HD void hfov_att(const ModelConst& c, const double* pose, double flag, const T* x, T* h) {
    T c0[3];
    camera_point(pose, x, c0);
    T cx = c0[0] + c.cam_off[0], cy = c0[1] + c.cam_off[1], cz = c0[2] + c.cam_off[2];
    h[0] = flag * datan2(cy, cx);
    h[1] = flag * datan2(cz, dsqrt(cx * cx + cy * cy));
}

// robot commands (quad_rollpitchyawrate.py:45-46)
SDF_HD void u_to_TRPYr(const ModelConst& c, const double* u, double* cmd) {
    cmd[0] = c.gamma * u[0] * c.mass;
    cmd[1] = c.roll * u[1];
    cmd[2] = c.pitch * u[2];
    cmd[3] = c.wz * u[3];
}

// [W_R_B^T W_a, wz]: u_to_acc of quad_rollpitchyawrate.py:45, quad_rollpitchyawrate_tau.py:46, quad_acc.py:36.
// For the thrust models W_a = W_R_B (0,0,gamma) - g e_z, hence W_R_B^T W_a = (0,0,gamma) - g * (third row of W_R_B).
// props: W_a = W_R_B F - g e_z with F = Gf wp^2 / m, hence W_R_B^T W_a = F - g * (third row of W_R_B); fourth entry w_z.
template <int MODEL>
SDF_HD void u_to_acc(const ModelConst& c, const double* x, const double* u, double* cmd) {
    cmd[3] = MODEL == MODEL_PROPS ? x[12] : c.wz * u[3];
    if (MODEL == MODEL_ATT) {
        double r = c.roll * u[1], pt = c.pitch * u[2];
        cmd[0] = c.g * sin(pt);                       // third row of Rz(psi) V_R_B(roll, pitch) is (-sp, sr cp, cr cp)
        cmd[1] = -c.g * sin(r) * cos(pt);
        cmd[2] = c.gamma * u[0] - c.g * cos(r) * cos(pt);
        return;
    }
    double inv_n = 1.0 / sqrt(x[3] * x[3] + x[4] * x[4] + x[5] * x[5] + x[6] * x[6]);
    double q0 = x[3] * inv_n, q1 = x[4] * inv_n, q2 = x[5] * inv_n, q3 = x[6] * inv_n;
    double R[9] = {q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3, 2 * (q1 * q2 - q0 * q3), 2 * (q1 * q3 + q0 * q2),      // utils/math.py:7-23
                   2 * (q1 * q2 + q0 * q3), q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3, 2 * (q2 * q3 - q0 * q1),
                   2 * (q1 * q3 - q0 * q2), 2 * (q2 * q3 + q0 * q1), q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3};
    if (MODEL == MODEL_ATT_TAU) {
        cmd[0] = -c.g * R[6];
        cmd[1] = -c.g * R[7];
        cmd[2] = c.gamma * u[0] - c.g * R[8];
    } else if (MODEL == MODEL_PROPS) {
        double s[4];
        for (int j = 0; j < 4; ++j) { const double wp = c.wp_max * u[j]; s[j] = wp * wp; }
        for (int i = 0; i < 3; ++i)
            cmd[i] = c.Gf_m[i * 4 + 0] * s[0] + c.Gf_m[i * 4 + 1] * s[1] + c.Gf_m[i * 4 + 2] * s[2] + c.Gf_m[i * 4 + 3] * s[3] - c.g * R[6 + i];
    } else {
        double a0 = c.ax * u[0], a1 = c.ay * u[1], a2 = c.az * u[2];
        cmd[0] = R[0] * a0 + R[3] * a1 + R[6] * a2;
        cmd[1] = R[1] * a0 + R[4] * a1 + R[7] * a2;
        cmd[2] = R[2] * a0 + R[5] * a1 + R[8] * a2;
    }
}

}  // namespace sdfnmpc
