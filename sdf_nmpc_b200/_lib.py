"""ctypes binding of libsdfnmpc_b200.so (the C ABI declared in include/sdfnmpc_b200.h).

There is no CPU path: if the library is missing (not built) or CUDA is unavailable the product
fails loudly.  ``load()`` only dlopen()s and sets prototypes; it makes no compute call, so the
CPU-only test-suite can check the exported symbols.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SDFNMPC_LIB selects another build of the SAME library (kernel A/B experiments, tools/qp_ab.py); default = the in-tree build
LIB_PATH = os.environ.get('SDFNMPC_LIB') or os.path.join(HERE, 'lib', 'libsdfnmpc_b200.so')
MAX_N = 64

_lib = None


class SdfDesc(C.Structure):
    _fields_ = [('n_dirs', C.c_int32), ('nb_freqs', C.c_int32), ('hidden', C.c_int32 * 4), ('size_latent', C.c_int32),
                ('res_state', C.c_int32), ('res_latent', C.c_int32), ('act', C.c_int32), ('w0', C.c_float),
                ('max_df', C.c_float)]


class SolverDesc(C.Structure):
    _fields_ = [('N', C.c_int32), ('max_batch', C.c_int32), ('qp_iter_max', C.c_int32), ('h_at_node0', C.c_int32),
                ('lm', C.c_double), ('qp_tol', C.c_double), ('qp_tol_res', C.c_double),
                ('g', C.c_double), ('gamma', C.c_double), ('roll', C.c_double), ('pitch', C.c_double),
                ('wz', C.c_double), ('mass', C.c_double), ('cam_off', C.c_double * 3),
                ('lbu', C.c_double * 4), ('ubu', C.c_double * 4), ('lh', C.c_double * 3), ('uh', C.c_double * 3),
                ('zl', C.c_double * 3), ('Zl', C.c_double * 3), ('dt', C.c_double * MAX_N),
                ('model', C.c_int32), ('reserved', C.c_int32), ('ax', C.c_double), ('ay', C.c_double), ('az', C.c_double),
                ('tau_roll', C.c_double), ('tau_pitch', C.c_double), ('sdf_cost_w', C.c_double),
                ('wp_max', C.c_double), ('Gf_m', C.c_double * 12), ('JGt', C.c_double * 12), ('inertia', C.c_double * 3),
                ('u_hover_props', C.c_double),
                ('term_lift', C.c_int32), ('brake_degree', C.c_int32), ('brake_coeffs', C.c_double * 56),
                ('brake_bounds', C.c_double * 2), ('slack_brake', C.c_double * 2), ('fov_term', C.c_double * 2),
                ('vlim', C.c_double * 3), ('z_hard', C.c_double)]


class EncoderDesc(C.Structure):
    _fields_ = [('height', C.c_int32), ('width', C.c_int32), ('size_latent', C.c_int32), ('max_batch', C.c_int32),
                ('is_depth', C.c_int32), ('is_normalized', C.c_int32), ('dmax_scale', C.c_float),
                ('hfov', C.c_float), ('vfov', C.c_float)]


FP = C.POINTER(C.c_float)


class EncoderWeights(C.Structure):
    _fields_ = [('conv0_w', FP), ('conv0_b', FP),
                ('conv1_w', FP * 4), ('bn1', (FP * 4) * 4),
                ('conv2_w', FP * 4), ('bn2', (FP * 4) * 4),
                ('sc_w', FP * 4), ('sc_bn', (FP * 4) * 4),
                ('mean_w', FP), ('mean_b', FP), ('bn_eps', C.c_float)]


class DecoderDesc(C.Structure):
    _fields_ = [('height', C.c_int32), ('width', C.c_int32), ('size_latent', C.c_int32), ('max_batch', C.c_int32)]


class DecoderWeights(C.Structure):
    _fields_ = [('lin_w', FP), ('lin_b', FP),
                ('ct1_w', FP * 4), ('bn1', (FP * 4) * 4),
                ('ct2_w', FP * 4), ('bn2', (FP * 4) * 4),
                ('sc_w', FP * 4), ('sc_bn', (FP * 4) * 4),
                ('out_w', FP), ('out_b', FP), ('bn_eps', C.c_float)]


CASADI_COMPANIONS = ('', '_alloc_mem', '_init_mem', '_free_mem', '_checkout', '_release', '_incref', '_decref', '_n_in',
                     '_n_out', '_default_in', '_name_in', '_name_out', '_sparsity_in', '_sparsity_out', '_work')

EXPORTS = (['sdfnmpc_version', 'sdfnmpc_launch_count', 'sdfnmpc_sizeof', 'sdfnmpc_sdf_create', 'sdfnmpc_sdf_destroy',
            'sdfnmpc_sdf_set_latents', 'sdfnmpc_sdf_eval', 'sdfnmpc_sdf_eval_host', 'sdfnmpc_solver_create',
            'sdfnmpc_solver_destroy', 'sdfnmpc_solver_init_traj', 'sdfnmpc_solver_shift', 'sdfnmpc_solver_rti_step',
            'sdfnmpc_solver_state', 'sdfnmpc_solver_diag', 'sdfnmpc_solver_qp_residual', 'sdfnmpc_solver_timing_begin', 'sdfnmpc_solver_timing_end', 'sdfnmpc_plant_step', 'sdfnmpc_cmd_trpyr', 'sdfnmpc_cmd_acc', 'sdfnmpc_cmd_props', 'sdfnmpc_solver_dims', 'sdfnmpc_solver_horizon', 'sdfnmpc_colcheck', 'sdfnmpc_df_signed', 'sdfnmpc_df_unsigned', 'sdfnmpc_fail_count_update',
            'sdfnmpc_encoder_create', 'sdfnmpc_encoder_destroy', 'sdfnmpc_encoder_encode',
            'sdfnmpc_encoder_encode_host', 'sdfnmpc_decoder_create', 'sdfnmpc_decoder_destroy', 'sdfnmpc_decoder_decode',
            'sdfnmpc_decoder_decode_host', 'sdfnmpc_l4c_bind', 'sdfnmpc_l4c_load_blob']
           + [f + s for f in ('sdf_l4c', 'jac_sdf_l4c') for s in CASADI_COMPANIONS])


class SdfNmpcError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        names = {-1: 'invalid argument', -2: 'no CUDA device (this library has no CPU path)', -3: 'call order violated'}
        raise SdfNmpcError(f'{what} failed: {names.get(rc, "cudaError %d" % rc)}')


def load():
    """dlopen the library and declare prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SdfNmpcError(f'{LIB_PATH} is missing - run `python -m sdf_nmpc_b200.build` '
                           '(the CUDA extension is mandatory; there is no CPU fallback)')
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dp, fp, ip = C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p
    lib.sdfnmpc_version.restype = C.c_char_p
    lib.sdfnmpc_launch_count.restype = C.c_uint64
    lib.sdfnmpc_sizeof.argtypes, lib.sdfnmpc_sizeof.restype = [C.c_int32], C.c_int32
    # the ctypes restatements of the descriptor structs must have the library's layout: refuse loudly otherwise
    for which, cls in enumerate((SdfDesc, SolverDesc, EncoderDesc, EncoderWeights, DecoderDesc, DecoderWeights)):
        if lib.sdfnmpc_sizeof(which) != C.sizeof(cls):
            raise SdfNmpcError(f'{cls.__name__}: ctypes layout is {C.sizeof(cls)} bytes, {LIB_PATH} has {lib.sdfnmpc_sizeof(which)} '
                               '(stale build? run `python -m sdf_nmpc_b200.build --force`)')
    lib.sdfnmpc_sdf_create.argtypes = [C.POINTER(vp), C.POINTER(SdfDesc)] + [fp] * 11
    lib.sdfnmpc_sdf_destroy.argtypes = [vp]
    lib.sdfnmpc_sdf_set_latents.argtypes = [vp, fp, i32, vp]
    lib.sdfnmpc_sdf_eval.argtypes = [vp, fp, ip, i32, i64, fp, fp, vp]
    lib.sdfnmpc_sdf_eval_host.argtypes = [vp, fp, fp, i64, fp, fp]
    lib.sdfnmpc_solver_create.argtypes = [C.POINTER(vp), C.POINTER(SolverDesc), vp]
    lib.sdfnmpc_solver_destroy.argtypes = [vp]
    lib.sdfnmpc_solver_init_traj.argtypes = [vp, dp, i32, vp]
    lib.sdfnmpc_solver_shift.argtypes = [vp, i32, i32, vp]
    lib.sdfnmpc_solver_rti_step.argtypes = [vp, i32] + [dp] * 8 + [fp, dp, ip, ip, vp]
    lib.sdfnmpc_solver_state.argtypes = [vp] + [C.POINTER(vp)] * 4
    lib.sdfnmpc_solver_diag.argtypes = [vp] + [C.POINTER(vp)] * 3
    lib.sdfnmpc_solver_qp_residual.argtypes = [vp, C.POINTER(vp)]
    lib.sdfnmpc_solver_timing_begin.argtypes = [vp, i32]
    lib.sdfnmpc_solver_timing_end.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_int32)]
    lib.sdfnmpc_plant_step.argtypes = [vp, i32, dp, dp, C.c_double, vp]
    lib.sdfnmpc_cmd_trpyr.argtypes = [vp, i32, dp, dp, vp]
    lib.sdfnmpc_cmd_acc.argtypes = [vp, i32, dp, dp, dp, vp]
    lib.sdfnmpc_cmd_props.argtypes = [vp, i32, dp, dp, vp]
    lib.sdfnmpc_solver_dims.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    lib.sdfnmpc_solver_horizon.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
    lib.sdfnmpc_fail_count_update.argtypes = [ip, vp, i32, vp]
    lib.sdfnmpc_colcheck.argtypes = [fp, i32, i32, i32, fp, ip, i64] + [C.c_float] * 4 + [i32, i32, i32, vp, vp]
    lib.sdfnmpc_df_signed.argtypes = [fp, i32, i32, i32, fp, ip, i64, fp, fp, i32] + [C.c_float] * 3 + [i32, i32, C.c_float, C.c_float, fp, fp, vp]
    lib.sdfnmpc_df_unsigned.argtypes = [fp, i32, i32, i32, i32, fp, fp, ip, i64] + [C.c_float] * 3 + [i32, i32, C.c_float, fp, fp, vp]
    lib.sdfnmpc_encoder_create.argtypes = [C.POINTER(vp), C.POINTER(EncoderDesc), C.POINTER(EncoderWeights)]
    lib.sdfnmpc_encoder_destroy.argtypes = [vp]
    lib.sdfnmpc_encoder_encode.argtypes = [vp, fp, i32, fp, vp]
    lib.sdfnmpc_encoder_encode_host.argtypes = [vp, fp, i32, fp]
    lib.sdfnmpc_decoder_create.argtypes = [C.POINTER(vp), C.POINTER(DecoderDesc), C.POINTER(DecoderWeights)]
    lib.sdfnmpc_decoder_destroy.argtypes = [vp]
    lib.sdfnmpc_decoder_decode.argtypes = [vp, fp, i32, fp, vp]
    lib.sdfnmpc_decoder_decode_host.argtypes = [vp, fp, i32, fp]
    lib.sdfnmpc_l4c_bind.argtypes = [vp]
    lib.sdfnmpc_l4c_load_blob.argtypes = [C.c_char_p]
    for f in ('sdf_l4c', 'jac_sdf_l4c'):
        getattr(lib, f).argtypes = [C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.POINTER(C.c_double)), vp, vp, C.c_int]
        getattr(lib, f + '_n_in').restype = C.c_longlong
        getattr(lib, f + '_n_out').restype = C.c_longlong
        getattr(lib, f + '_name_in').restype = C.c_char_p
        getattr(lib, f + '_name_in').argtypes = [C.c_longlong]
        getattr(lib, f + '_name_out').restype = C.c_char_p
        getattr(lib, f + '_name_out').argtypes = [C.c_longlong]
        getattr(lib, f + '_sparsity_in').restype = C.POINTER(C.c_longlong)
        getattr(lib, f + '_sparsity_in').argtypes = [C.c_longlong]
        getattr(lib, f + '_sparsity_out').restype = C.POINTER(C.c_longlong)
        getattr(lib, f + '_sparsity_out').argtypes = [C.c_longlong]
        getattr(lib, f + '_work').argtypes = [C.POINTER(C.c_longlong)] * 4
    _lib = lib
    return lib


def launch_count():
    return int(load().sdfnmpc_launch_count())
