"""YAML configuration with attribute access - the host-side mirror of
``sdf_nmpc/utils/config.py:9-44`` (``AttrDict`` / ``Config``), including the derived sensor fields
``sensor.B_p_C`` / ``sensor.B_R_C`` (``config.py:43-44``) and the vfov sanity check (``config.py:38-41``,
``utils/math.py:286-291``).  Any YAML written for the reference loads here unchanged; the optional
``b200`` block carries the solver numerics of this implementation.
"""
import os

import numpy as np
import yaml

DEFAULT_YAML = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'config', 'quad_att.yaml')


class AttrDict(dict):
    def __init__(self, dictionary=()):
        super().__init__()
        for k, v in dict(dictionary).items():
            self[k] = v

    def __setitem__(self, key, value):
        if isinstance(value, dict) and not isinstance(value, AttrDict):
            value = AttrDict(value)
        elif isinstance(value, list):
            value = [AttrDict(v) if isinstance(v, dict) else v for v in value]
        super().__setitem__(key, value)
        super().__setattr__(key, value)

    __setattr__ = __setitem__


def euler2rot(e):
    """ZYX rotation matrix of [roll, pitch, yaw] (utils/math.py:26-54)."""
    r, p, y = e
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([[cp * cy, sr * sp * cy - cr * sy, cr * sp * cy + sr * sy],
                     [cp * sy, sr * sp * sy + cr * cy, cr * sp * sy - sr * cy],
                     [-sp, sr * cp, cr * cp]])


def get_vfov(hfov, aspect_ratio, is_spherical):
    return hfov / aspect_ratio if is_spherical else float(np.arctan(np.tan(hfov) / aspect_ratio))


class Config(AttrDict):
    def __init__(self, config_file=None):
        with open(config_file or DEFAULT_YAML, 'r') as f:
            super().__init__(yaml.load(f, Loader=yaml.FullLoader))
        s = self.sensor
        assert abs(get_vfov(s.hfov, s.aspect_ratio, s.is_spherical) - s.vfov) < 0.1, 'check sensor fov in config file'
        s.B_p_C = np.asarray(self.robot.sensor_extrinsics.position, float)
        s.B_R_C = euler2rot(self.robot.sensor_extrinsics.orientation)
        if 'b200' not in self:
            self['b200'] = dict(qp_iter_max=100, qp_tol=1e-8, qp_tol_res=1e-7, h_at_node0=False)

    # ---- derived quantities of the OCP (ocp.py:21-27, gen_model.py:35, cost_const_helpers.py:67-75) ----
    def shooting_dt(self):
        m = self.mpc
        if m.uniform_dt:
            nodes = np.linspace(0, m.T, m.N + 1)
        else:
            ns, dts = m.nb_short_nodes, m.control_loop_time * 1e-3
            nodes = np.hstack([np.linspace(0, dts * (ns - 1), ns), np.linspace(dts * ns, m.T, m.N - ns + 1)])
        return np.diff(nodes)

    def h_bounds(self, max_df):
        hl, vl = self.sensor.hfov * self.mpc.fov_ratio, self.sensor.vfov * self.mpc.fov_ratio
        return (np.array([-hl, -vl, self.robot.size.xy + self.mpc.bound_margin]), np.array([hl, vl, max_df + 0.2]))

    def slack_weights(self):
        w = self.mpc.weights
        return (np.array([w.slack_fov[0], w.slack_fov[0], w.slack_df[0]], float),
                np.array([w.slack_fov[1], w.slack_fov[1], w.slack_df[1]], float))

    def cam_offset(self):
        return self.sensor.B_R_C.T @ self.sensor.B_p_C + np.array([self.mpc.fov_const_offset, 0.0, 0.0])
