"""Off-axis positional-embedding direction tables.

Same tables as the reference's ``PositionEmbedding`` registers as the buffer
``dirs`` (``sdf_nmpc/utils/embeddings.py:24-100``): unit columns, float32,
shape [3, D].  When a real TorchScript SDF is loaded (``weights.sdf_from_module``) the
buffer is read from the module instead; this table is for synthetic networks.
"""
import itertools
import numpy as np

_PHI = (1.0 + np.sqrt(5.0)) / 2.0
ALIASES = {'pos': 'none', 'none': 'none', 'cube': 'cube', 'oct': 'octohedron', 'octohedron': 'octohedron',
           'dod': 'dodecahedron', 'dodecahedron': 'dodecahedron', 'ico': 'icosahedron', 'icosahedron': 'icosahedron'}


def directions(name):
    kind = ALIASES[name]
    if kind == 'none':
        pts = np.eye(3)
    elif kind == 'cube':
        pts = np.array([s * np.eye(3)[i] for i in range(3) for s in (-1.0, 1.0)])
    elif kind == 'octohedron':
        pts = np.array(list(itertools.product((-1.0, 1.0), repeat=3)))
    elif kind == 'dodecahedron':
        # vertices of a regular icosahedron; ordering as the reference table
        yz = [(-1.0, -_PHI), (1.0, -_PHI), (-1.0, _PHI), (1.0, _PHI)]
        pts = np.array([[0, a, b] for a, b in yz] + [[a, 0, b] for a, b in yz] +
                       [[a, b, 0] for a, b in [(-1.0, -_PHI), (1.0, -_PHI), (-1.0, _PHI), (1.0, _PHI)]])
    elif kind == 'icosahedron':
        h = 1.0 / _PHI
        cube = list(itertools.product((1.0, -1.0), repeat=3))
        pts = np.array(cube +
                       [[0, a, b] for a in (_PHI, -_PHI) for b in (h, -h)] +
                       [[b, 0, a] for b in (h, -h) for a in (_PHI, -_PHI)] +
                       [[a, b, 0] for a in (_PHI, -_PHI) for b in (h, -h)])
    else:  # pragma: no cover
        raise AttributeError('Unknown off-axis projection mode')
    d = pts.astype(np.float32).T
    if kind not in ('none', 'cube'):
        # torch.linalg.vector_norm over dim 0 accumulates the three squares with fp32 FMAs; reproduce bit-exactly
        acc = np.zeros(d.shape[1], np.float32)
        for r in range(3):
            acc = (d[r].astype(np.float64) ** 2 + acc.astype(np.float64)).astype(np.float32)
        d = d / np.sqrt(acc)
    return np.ascontiguousarray(d, dtype=np.float32)
