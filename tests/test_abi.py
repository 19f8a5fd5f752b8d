"""The C-ABI library loads on a CPU-only box and exports every symbol include/sdfnmpc_b200.h declares
(no compute call is made); the product refuses to run without CUDA instead of falling back."""
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'sdfnmpc_b200.h')).read()
    body = src.split('#define SDFNMPC_DECLARE_CASADI')[0]
    names = set(re.findall(r'\b(sdfnmpc_[a-z0-9_]+)\s*\(', body))
    macro = src.split('#define SDFNMPC_DECLARE_CASADI(F)')[1].split('/* sdf_l4c:')[0]
    suffixes = [''] + re.findall(r'F##(_[a-z_]+)\(', macro)
    for f in ('sdf_l4c', 'jac_sdf_l4c'):
        names |= {f + s for s in suffixes}
    return names


def test_library_exports_every_declared_symbol():
    from sdf_nmpc_b200 import _lib, build
    build.build()
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 24 + 32
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert set(_lib.EXPORTS) <= declared | {'sdf_l4c', 'jac_sdf_l4c'}
    assert b'sm_100a' in lib.sdfnmpc_version()
    # metadata entry points of the CasADi ABI need no device
    assert lib.sdf_l4c_n_in() == 1 and lib.jac_sdf_l4c_n_in() == 2 and lib.jac_sdf_l4c_n_out() == 1
    assert lib.sdf_l4c_name_in(0) == b'i0' and lib.jac_sdf_l4c_name_in(1) == b'out_o0'


def test_descriptor_layouts_agree_between_header_library_and_ctypes(tmp_path):
    """The header compiles as plain C (gcc, no CUDA headers) and its descriptor structs have the size the library reports
    (``sdfnmpc_sizeof``) and the size of the ctypes restatements in ``_lib.py`` - a binding with another layout would hand the
    create functions garbage."""
    import ctypes as C
    import subprocess
    from sdf_nmpc_b200 import _lib, build
    build.build()
    lib = _lib.load()
    names = ['sdfnmpc_sdf_desc', 'sdfnmpc_solver_desc', 'sdfnmpc_encoder_desc', 'sdfnmpc_encoder_weights', 'sdfnmpc_decoder_desc',
             'sdfnmpc_decoder_weights']
    src = tmp_path / 'sizes.c'
    src.write_text('#include <stdio.h>\n#include "sdfnmpc_b200.h"\nint main(void) { printf("' + ' '.join(['%zu'] * len(names)) + '\\n", '
                   + ', '.join(f'sizeof({n})' for n in names) + '); return 0; }\n')
    exe = tmp_path / 'sizes'
    subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)], check=True)
    c_sizes = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    py = (_lib.SdfDesc, _lib.SolverDesc, _lib.EncoderDesc, _lib.EncoderWeights, _lib.DecoderDesc, _lib.DecoderWeights)
    for which, (n, cls) in enumerate(zip(names, py)):
        assert c_sizes[which] == lib.sdfnmpc_sizeof(which) == C.sizeof(cls), (n, c_sizes[which], lib.sdfnmpc_sizeof(which), C.sizeof(cls))
    assert lib.sdfnmpc_sizeof(6) == -1


def test_sass_is_sm100a_only():
    import subprocess
    from sdf_nmpc_b200 import _lib
    out = subprocess.run(['cuobjdump', '-lelf', _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip('cuobjdump unavailable')
    archs = set(re.findall(r'sm_\d+a?', out.stdout))
    assert archs == {'sm_100a'}, archs


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from sdf_nmpc_b200 import _lib, synth
    from sdf_nmpc_b200.sdf import SdfEvaluator
    with pytest.raises(_lib.SdfNmpcError):
        SdfEvaluator(synth.make_sdf_net(seed=0, **synth.ARCH_A))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'sdf_nmpc_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, re.M), f
