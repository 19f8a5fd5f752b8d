// Forward-mode dual numbers (value + one tangent) in double precision.
//
// The reference obtains every derivative of its model from CasADi's symbolic AD
// at code-generation time (sdf_nmpc/model/base_model.py:83-109 feeds MX graphs
// to acados, which generates expl_vde_forw / cost_y_fun_jac_ut_xt / constr_h_fun_jac).
// Here one GPU thread propagates ONE tangent direction through the templated
// model functions, so the derivative of the *discrete RK4 map* (identical to
// integrating the variational ODE with the same tableau) comes out exactly.
#pragma once
#include "common.cuh"

namespace sdfnmpc {

struct Dual {
    double v, d;
    SDF_HD Dual() : v(0.0), d(0.0) {}
    SDF_HD Dual(double v_) : v(v_), d(0.0) {}
    SDF_HD Dual(double v_, double d_) : v(v_), d(d_) {}
};

SDF_HD Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
SDF_HD Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
SDF_HD Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
SDF_HD Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, a.d * b.v + a.v * b.d); }
SDF_HD Dual operator/(Dual a, Dual b) {
    double inv = 1.0 / b.v;
    double q = a.v * inv;
    return Dual(q, (a.d - q * b.d) * inv);
}
SDF_HD Dual operator+(Dual a, double b) { return Dual(a.v + b, a.d); }
SDF_HD Dual operator+(double a, Dual b) { return Dual(a + b.v, b.d); }
SDF_HD Dual operator-(Dual a, double b) { return Dual(a.v - b, a.d); }
SDF_HD Dual operator-(double a, Dual b) { return Dual(a - b.v, -b.d); }
SDF_HD Dual operator*(Dual a, double b) { return Dual(a.v * b, a.d * b); }
SDF_HD Dual operator*(double a, Dual b) { return Dual(a * b.v, a * b.d); }
SDF_HD Dual operator/(Dual a, double b) { return Dual(a.v / b, a.d / b); }

SDF_HD Dual dsqrt(Dual a) {
    double s = sqrt(a.v);
    return Dual(s, 0.5 * a.d / s);
}
SDF_HD Dual dsin(Dual a) { return Dual(sin(a.v), cos(a.v) * a.d); }
SDF_HD Dual dcos(Dual a) { return Dual(cos(a.v), -sin(a.v) * a.d); }
SDF_HD Dual datan2(Dual y, Dual x) {
    double r2 = x.v * x.v + y.v * y.v;
    return Dual(atan2(y.v, x.v), (x.v * y.d - y.v * x.d) / r2);
}

SDF_HD Dual dasin(Dual a) { return Dual(asin(a.v), a.d / sqrt(1.0 - a.v * a.v)); }

// scalar overloads so that the model templates also instantiate on plain double
SDF_HD double dasin(double a) { return asin(a); }
SDF_HD double dsqrt(double a) { return sqrt(a); }
SDF_HD double dsin(double a) { return sin(a); }
SDF_HD double dcos(double a) { return cos(a); }
SDF_HD double datan2(double y, double x) { return atan2(y, x); }
SDF_HD double value(double a) { return a; }
SDF_HD double value(Dual a) { return a.v; }
SDF_HD double tangent(double) { return 0.0; }
SDF_HD double tangent(Dual a) { return a.d; }

}  // namespace sdfnmpc
