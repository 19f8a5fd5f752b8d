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

struct ModelConst {
    double g, gamma, roll, pitch, wz, mass;
    double cam_off[3];    // B_R_C^T B_p_C + (fov_const_offset, 0, 0)
};

template <class T>
SDF_HD void f_att(const ModelConst& c, const T* x, const T* u, T* dx) {
    T q0 = x[3], q1 = x[4], q2 = x[5], q3 = x[6];
    T inv_n = 1.0 / dsqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
    T h0 = q0 * inv_n, h1 = q1 * inv_n, h2 = q2 * inv_n, h3 = q3 * inv_n;
    // yaw-only rotation: theta_z = atan2(q3, q0), W_R_V = Rz(2 theta_z)
    T psi = 2.0 * datan2(h3, h0);
    T cps = dcos(psi), sps = dsin(psi);
    T gam = c.gamma * u[0];
    T r = c.roll * u[1], pt = c.pitch * u[2], wz = c.wz * u[3];
    T cr = dcos(r), sr = dsin(r), cp = dcos(pt), sp = dsin(pt);
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

template <class T>
SDF_HD void rk4_step(const ModelConst& c, const T* x, const T* u, double dt, T* xn) {
    T k1[ATT_NX], k2[ATT_NX], k3[ATT_NX], k4[ATT_NX], xs[ATT_NX];
    f_att(c, x, u, k1);
    for (int i = 0; i < ATT_NX; ++i) xs[i] = x[i] + (0.5 * dt) * k1[i];
    f_att(c, xs, u, k2);
    for (int i = 0; i < ATT_NX; ++i) xs[i] = x[i] + (0.5 * dt) * k2[i];
    f_att(c, xs, u, k3);
    for (int i = 0; i < ATT_NX; ++i) xs[i] = x[i] + dt * k3[i];
    f_att(c, xs, u, k4);
    for (int i = 0; i < ATT_NX; ++i) xn[i] = x[i] + (dt / 6.0) * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
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

}  // namespace sdfnmpc
