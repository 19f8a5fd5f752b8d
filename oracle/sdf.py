"""Oracle: neural SDF value + input gradient (numpy, float64 by default).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.

Restates
  * ``sdf_nmpc/network/neural_df.py:91-103``  (NeuralDF.forward)
  * ``sdf_nmpc/utils/embeddings.py:21,37-111`` (PositionEmbedding)
  * ``sdf_nmpc/utils/activation.py:12-13``     (Sine)
and the analytic reverse sweep of SURVEY.md Appendix A (what L4CasADi's traced
``jacrev`` returns in columns 0..2 of ``jac_sdf_l4c``; ``gen_model.py:38-39``).

A network is a plain dict of numpy arrays / scalars (``net``):
  dirs [3,D], nb_freqs, w0, act in {'sin','relu','softplus'},
  res in {'full','state','latent','none'}, max_df, size_latent,
  W1 [h1,E+L] b1, W2 [h2,h1] b2, W3 [h3,h2(+E)(+L)] b3, W4 [h4,h3] b4, W5 [1,h4] b5.
"""
import numpy as np


def embedding_dirs(proj):
    """Projection directions, columns unit-norm (embeddings.py:24-100)."""
    phi = (1 + np.sqrt(5)) / 2
    if proj in ('none', 'pos'):
        d = np.eye(3)
    elif proj == 'cube':
        d = np.array([[-1, 0, 0], [1, 0, 0], [0, -1, 0], [0, 1, 0], [0, 0, -1], [0, 0, 1]], float).T
    elif proj in ('oct', 'octohedron'):
        d = np.array([[a, b, c] for a in (-1, 1) for b in (-1, 1) for c in (-1, 1)], float).T
    elif proj in ('dod', 'dodecahedron'):
        d = np.array([[0, -1, -phi], [0, 1, -phi], [0, -1, phi], [0, 1, phi],
                      [-1, 0, -phi], [1, 0, -phi], [-1, 0, phi], [1, 0, phi],
                      [-1, -phi, 0], [1, -phi, 0], [-1, phi, 0], [1, phi, 0]], float).T
    elif proj in ('ico', 'icosahedron'):
        h = 1 / phi
        d = np.array([[1, 1, 1], [1, 1, -1], [1, -1, 1], [1, -1, -1],
                      [-1, 1, 1], [-1, 1, -1], [-1, -1, 1], [-1, -1, -1],
                      [0, phi, h], [0, phi, -h], [0, -phi, h], [0, -phi, -h],
                      [h, 0, phi], [h, 0, -phi], [-h, 0, phi], [-h, 0, -phi],
                      [phi, h, 0], [phi, -h, 0], [-phi, h, 0], [-phi, -h, 0]], float).T
    else:
        raise AttributeError('Unknown off-axis projection mode')
    # the reference builds the table in float32 and normalises there
    d = d.astype(np.float32)
    if proj not in ('none', 'pos', 'cube'):
        # torch.linalg.vector_norm over dim 0 accumulates the three squares with fp32 FMAs; reproduce bit-exactly
        acc = np.zeros(d.shape[1], np.float32)
        for r in range(3):
            acc = (d[r].astype(np.float64) ** 2 + acc.astype(np.float64)).astype(np.float32)
        d = d / np.sqrt(acc)
    return d.astype(np.float32)


def _act(net, a):
    k = net['act']
    if k == 'sin':
        return np.sin(net['w0'] * a)
    if k == 'relu':
        return np.maximum(a, 0)
    if k == 'softplus':
        return np.logaddexp(0, a)
    raise AttributeError(k)


def _dact(net, a):
    k = net['act']
    if k == 'sin':
        return net['w0'] * np.cos(net['w0'] * a)
    if k == 'relu':
        return (a > 0).astype(a.dtype)
    if k == 'softplus':
        return 1 / (1 + np.exp(-a))
    raise AttributeError(k)


def embed(net, s):
    """[n,3] -> [n,E] exactly as embeddings.py:106-111 (d-major, f-minor)."""
    dt = s.dtype
    dirs = net['dirs'].astype(dt)
    F = int(net['nb_freqs'])
    if F == 0 or net.get('embed', 'x') == 'none':
        return s
    freq = (2.0 ** np.linspace(0, F - 1, F)).astype(dt)
    proj = s @ dirs
    xb = (proj[..., None] * freq).reshape(s.shape[0], -1)
    half_pi = dt.type(0.5 * np.pi)
    return np.concatenate([s, np.sin(xb), np.sin(xb + half_pi)], axis=1)


def forward(net, s, z, dtype=np.float64):
    """SDF value d [n] for points s [n,3] and latents z [n,L] (or [L])."""
    s = np.asarray(s, dtype)
    z = np.broadcast_to(np.asarray(z, dtype), (s.shape[0], net['size_latent']))
    W = {k: np.asarray(net[k], dtype) for k in ('W1', 'b1', 'W2', 'b2', 'W3', 'b3', 'W4', 'b4', 'W5', 'b5')}
    e = embed(net, s)
    x = np.concatenate([e, z], 1)
    x = _act(net, x @ W['W1'].T + W['b1'])
    x = _act(net, x @ W['W2'].T + W['b2'])
    if net['res'] in ('full', 'state'):
        x = np.concatenate([x, e], 1)
    if net['res'] in ('full', 'latent'):
        x = np.concatenate([x, z], 1)
    x = _act(net, x @ W['W3'].T + W['b3'])
    x = _act(net, x @ W['W4'].T + W['b4'])
    return (x @ W['W5'].T + W['b5'])[:, 0]


def value_and_grad(net, s, z, dtype=np.float64):
    """(d [n], grad_s d [n,3]) by the hand-derived reverse sweep."""
    s = np.asarray(s, dtype)
    n = s.shape[0]
    L = net['size_latent']
    z = np.broadcast_to(np.asarray(z, dtype), (n, L))
    W = {k: np.asarray(net[k], dtype) for k in ('W1', 'b1', 'W2', 'b2', 'W3', 'b3', 'W4', 'b4', 'W5', 'b5')}
    dirs = net['dirs'].astype(dtype)
    F = int(net['nb_freqs'])
    D = dirs.shape[1]
    freq = (2.0 ** np.linspace(0, F - 1, F)).astype(dtype)
    proj = s @ dirs
    xb = (proj[..., None] * freq).reshape(n, D * F)
    half_pi = dtype(0.5 * np.pi)
    e = np.concatenate([s, np.sin(xb), np.sin(xb + half_pi)], 1)
    E = e.shape[1]
    h2n = W['W2'].shape[0]
    # forward
    a1 = np.concatenate([e, z], 1) @ W['W1'].T + W['b1']
    h1 = _act(net, a1)
    a2 = h1 @ W['W2'].T + W['b2']
    h2 = _act(net, a2)
    x3 = [h2]
    if net['res'] in ('full', 'state'):
        x3.append(e)
    if net['res'] in ('full', 'latent'):
        x3.append(z)
    a3 = np.concatenate(x3, 1) @ W['W3'].T + W['b3']
    h3 = _act(net, a3)
    a4 = h3 @ W['W4'].T + W['b4']
    h4 = _act(net, a4)
    d = (h4 @ W['W5'].T + W['b5'])[:, 0]
    # reverse
    g4 = W['W5'] * _dact(net, a4)                     # [n,h4]
    g3 = (g4 @ W['W4']) * _dact(net, a3)              # [n,h3]
    gx3 = g3 @ W['W3']                                # [n, h2 (+E)(+L)]
    g2 = gx3[:, :h2n] * _dact(net, a2)
    g1 = (g2 @ W['W2']) * _dact(net, a1)
    ge = g1 @ W['W1'][:, :E]
    if net['res'] in ('full', 'state'):
        ge = ge + gx3[:, h2n:h2n + E]
    gxb = ge[:, 3:3 + D * F] * np.cos(xb) + ge[:, 3 + D * F:] * np.cos(xb + half_pi)
    gproj = (gxb.reshape(n, D, F) * freq).sum(-1)
    grad = ge[:, :3] + gproj @ dirs.T
    return d, grad


def fold_latent(net, z, dtype=np.float64):
    """Per-latent folded biases (SURVEY.md 8a-1 'latent folding'):
    b1' = W1[:,E:] z + b1 [n,h1];  b3' = W3[:,h2+E:] z + b3 [n,h3]."""
    z = np.atleast_2d(np.asarray(z, dtype))
    E = embedding_size(net)
    W1, W3 = np.asarray(net['W1'], dtype), np.asarray(net['W3'], dtype)
    b1 = z @ W1[:, E:].T + np.asarray(net['b1'], dtype)
    b3 = np.broadcast_to(np.asarray(net['b3'], dtype), (z.shape[0], W3.shape[0])).copy()
    if net['res'] in ('full', 'latent'):
        off = np.asarray(net['W2']).shape[0] + (E if net['res'] == 'full' else 0)
        b3 = b3 + z @ W3[:, off:].T
    return b1, b3


def embedding_size(net):
    return 3 + 2 * int(net['nb_freqs']) * net['dirs'].shape[1]


def l4casadi_call(net, i0):
    """What ``sdf_l4c`` / ``jac_sdf_l4c`` return for one column vector i0 [3+L]
    (double in, fp32 model, double out).  The latent
    columns of the Jacobian are filled for completeness (reverse sweep on z)."""
    i0 = np.asarray(i0, np.float64).reshape(-1)
    x = i0.astype(np.float32)
    d, g = value_and_grad(net, x[None, :3], x[None, 3:], dtype=np.float32)
    return np.float64(d[0]), g[0].astype(np.float64)
