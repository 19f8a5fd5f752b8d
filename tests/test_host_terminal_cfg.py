"""Host logic of the terminal ingredients (flags.recursive_feasibility / stability, gen_model.py:73-149) that needs no device: where
the braking-distance coefficients come from, how many a degree needs, and the oracle's polynomial in the reference's term order."""
import os

import numpy as np
import pytest


def test_braking_coefficients_from_cache_dir_and_inline(tmp_path, monkeypatch):
    from sdf_nmpc_b200 import synth
    from sdf_nmpc_b200.config import Config
    from sdf_nmpc_b200.nmpc import braking_coeffs
    cfg = Config()
    co = synth.braking_coeffs(4)
    assert co.size == 35                                   # (d + 1)(d + 2)(d + 3) / 6 terms of a 3-variate polynomial of degree 4
    # gen_model.py:76-77: np.load(os.path.join(cache_dir(), cfg.mpc.braking_dist.coeff_file))
    path = tmp_path / cfg.mpc.braking_dist.coeff_file
    os.makedirs(path.parent)
    np.save(path, co)
    monkeypatch.setenv('SDFNMPC_CACHE_DIR', str(tmp_path))
    deg, got = braking_coeffs(cfg)
    assert deg == 4 and np.array_equal(got, co)
    monkeypatch.setenv('SDFNMPC_CACHE_DIR', str(tmp_path / 'elsewhere'))
    with pytest.raises(FileNotFoundError):
        braking_coeffs(cfg)
    cfg.mpc.braking_dist = dict(degree=2, coeffs=list(range(10)))
    deg, got = braking_coeffs(cfg)
    assert deg == 2 and got.tolist() == list(range(10))
    cfg.mpc.braking_dist = dict(degree=6, coeffs=[0.0] * 84)      # the descriptor carries at most degree 5 (56 coefficients)
    with pytest.raises(ValueError):
        braking_coeffs(cfg)


def test_oracle_polynomial_follows_the_reference_term_order():
    """utils/math.py:306-312: total degree, then a, then b, c = td - a - b; value and gradient against a direct sum."""
    from oracle import rti
    rng = np.random.default_rng(5)
    deg = 3
    terms = [(a, b, td - a - b) for td in range(deg + 1) for a in range(td + 1) for b in range(td + 1 - a)]
    co = rng.standard_normal(len(terms))
    v = rng.standard_normal((7, 3))
    val, grad = rti.poly3(v, deg, co)
    ref = sum(c * v[:, 0] ** a * v[:, 1] ** b * v[:, 2] ** cc for c, (a, b, cc) in zip(co, terms))
    assert np.abs(val - ref).max() < 1e-12
    eps = 1e-6
    for i in range(3):
        dv = np.zeros(3); dv[i] = eps
        fd = (rti.poly3(v + dv, deg, co)[0] - rti.poly3(v - dv, deg, co)[0]) / (2 * eps)
        assert np.abs(grad[:, i] - fd).max() < 1e-7
    # the first terms of the order: 1, z, y, x, z^2, y z, y^2, x z, x y, x^2
    assert terms[:10] == [(0, 0, 0), (0, 0, 1), (0, 1, 0), (1, 0, 0), (0, 0, 2), (0, 1, 1), (0, 2, 0), (1, 0, 1), (1, 1, 0), (2, 0, 0)]
