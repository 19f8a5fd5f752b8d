/* sdfnmpc_b200 - C ABI of the B200-native sdf-nmpc hot path.
 *
 * Plain C: pointers, sizes and POD descriptors only (no torch / C++ types).  All
 * `*_dev` pointers are CUDA device pointers owned by the caller; `stream` is a
 * cudaStream_t passed as void* (NULL = default stream).  Calls are stream-ordered
 * and do not synchronise unless their name ends in `_host`.  Every function
 * returns 0 on success, a cudaError_t value or a negative SDFNMPC_E_* code
 * otherwise.  One handle per GPU; handles are not thread-safe.
 *
 * What each entry point replaces in the reference (ntnu-arl/sdf-nmpc):
 *
 *   sdfnmpc_sdf_*      libsdf_l4c.so produced by L4CasADi 1.4.1 from the TorchScript SDF
 *                      (sdf_nmpc/gen_model.py:32-39,60; network/neural_df.py:91-103),
 *                      batched over query points and latents instead of batch-1.
 *   sdf_l4c, jac_sdf_l4c (+ the CasADi external-function companions, declared at the end)
 *                      the exact symbols acados links through
 *                      model_external_shared_lib_name='sdf_l4c' (sdf_nmpc/ocp.py:100-102).
 *   sdfnmpc_solver_*   AcadosOcpSolver.reset/set/cost_set/solve_for_x0/get for the default
 *                      OCP (sdf_nmpc/ocp.py:144-170, options :96-120), batched over instances.
 *   sdfnmpc_encoder_*  VaeWrapper.set_img + encode (sdf_nmpc/vae.py:15-40; network/vae.py:19-38;
 *                      utils/preprocessing.py:5-30,67-96), batched over frames.
 */
#ifndef SDFNMPC_B200_H
#define SDFNMPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDFNMPC_MAX_N 64           /* size of the dt array; the QP kernel additionally needs N <= 31 (node k <-> lane k) */
#define SDFNMPC_E_INVALID (-1)     /* bad argument / unsupported configuration */
#define SDFNMPC_E_NODEVICE (-2)    /* no CUDA device: this library has no CPU path */
#define SDFNMPC_E_STATE (-3)       /* call order violated (e.g. eval before set_latents) */

const char* sdfnmpc_version(void);
/* number of kernels this library has launched since load (all handles) */
uint64_t sdfnmpc_launch_count(void);
/* sizeof of the descriptor structs below as this build sees them, for bindings that restate them (ctypes, cgo, JNI): 0 sdf_desc,
 * 1 solver_desc, 2 encoder_desc, 3 encoder_weights, 4 decoder_desc, 5 decoder_weights; anything else -1.  A binding whose own
 * layout differs must refuse to call the create functions. */
int32_t sdfnmpc_sizeof(int32_t which);

/* ------------------------------------------------------------------ neural SDF ---------- */
typedef struct sdfnmpc_sdf* sdfnmpc_sdf_t;

typedef struct {
    int32_t n_dirs;          /* D: projection directions (embeddings.py:24-100)            */
    int32_t nb_freqs;        /* F                                                           */
    int32_t hidden[4];       /* layer_sizes (neural_df.py:23)                               */
    int32_t size_latent;     /* L                                                           */
    int32_t res_state;       /* res in {'full','state'}  (neural_df.py:97-98)               */
    int32_t res_latent;      /* res in {'full','latent'} (neural_df.py:99-100)              */
    int32_t act;             /* 0 sin(w0 x), 1 relu, 2 softplus (neural_df.py:40-47)        */
    float w0;
    float max_df;
} sdfnmpc_sdf_desc;

/* Weights are HOST pointers in torch's nn.Linear layout ([out][in], row-major):
 * dirs [3][D]; W1 [h1][E+L]; W2 [h2][h1]; W3 [h3][h2 + E*res_state + L*res_latent]; W4 [h4][h3]; W5 [1][h4]. */
int sdfnmpc_sdf_create(sdfnmpc_sdf_t* out, const sdfnmpc_sdf_desc* desc, const float* dirs,
                       const float* W1, const float* b1, const float* W2, const float* b2,
                       const float* W3, const float* b3, const float* W4, const float* b4,
                       const float* W5, const float* b5);
int sdfnmpc_sdf_destroy(sdfnmpc_sdf_t h);
/* Fold n_inst latents [n_inst][L] (device) into per-instance biases kept inside the handle. */
int sdfnmpc_sdf_set_latents(sdfnmpc_sdf_t h, const float* latents_dev, int32_t n_inst, void* stream);
/* d [n] and grad [n][3] for points [n][3]; the latent of point i is pt2inst_dev[i] if given,
 * else i / pts_per_inst.  sdf_dev or grad_dev may be NULL. */
int sdfnmpc_sdf_eval(sdfnmpc_sdf_t h, const float* pts_dev, const int32_t* pt2inst_dev, int32_t pts_per_inst,
                     int64_t n, float* sdf_dev, float* grad_dev, void* stream);
/* Host-buffer convenience (one latent): H2D, fold, eval, D2H, synchronise. */
int sdfnmpc_sdf_eval_host(sdfnmpc_sdf_t h, const float* pts, const float* latent, int64_t n, float* sdf, float* grad);

/* ------------------------------------------------------------------ NMPC solver --------- */
typedef struct sdfnmpc_solver* sdfnmpc_solver_t;

typedef struct {
    int32_t N;               /* shooting intervals (default.yaml:52)                        */
    int32_t max_batch;       /* capacity in instances                                       */
    int32_t qp_iter_max;     /* ocp.py:115                                                  */
    int32_t h_at_node0;      /* 0: nonlinear rows at nodes 1..N (acados 0.3.x), 1: also node 0 */
    double lm;               /* Levenberg-Marquardt (default.yaml:58)                       */
    double qp_tol;           /* complementarity tolerance                                   */
    double qp_tol_res;       /* residual tolerance                                          */
    double g, gamma, roll, pitch, wz, mass;   /* base_model.py:15, default.yaml:94-106,73   */
    double cam_off[3];       /* B_R_C^T B_p_C + (fov_const_offset,0,0)                      */
    double lbu[4], ubu[4];   /* quad_rollpitchyawrate.py:58-59                              */
    double lh[3], uh[3];     /* [hfov, vfov, sdf] bounds                                    */
    double zl[3], Zl[3];     /* slack L1 / L2 weights in h order (ocp.py:85-92)             */
    double dt[SDFNMPC_MAX_N];/* stage lengths (ocp.py:21-27)                                */
    int32_t model;           /* mpc.model (gen_model.py:15-20): 0 att, 1 att_tau, 2 acc, 3 props; yrefN/WN are [B][4] (7 for acc, 12 for props) */
    int32_t reserved;
    double ax, ay, az;       /* robot.limits.ax/ay/az (model acc)                           */
    double tau_roll, tau_pitch; /* attitude lag of att_tau (0 -> 0.12 s, quad_rollpitchyawrate_tau.py:19-20) */
    double sdf_cost_w;       /* flags.sdf_cost: weight of the stage cost row (1 - 0.5 sdf)^4 (gen_model.py:65-66: 20); 0 = off */
    /* model props (quad_props.py:18-52): rotor-speed scale limits.wp, allocation matrices with cf folded in as Gf / mass and
     * J^-1 Gt (row-major [3][4]), diagonal inertia, hover input of Ocp.init */
    double wp_max, Gf_m[12], JGt[12], inertia[3], u_hover_props;
    /* flags.recursive_feasibility / flags.stability (gen_model.py:73-149; model att only).  term_lift: 0 off, 1 recursive
     * feasibility (terminal sdf row replaced by rec_feas = sdf_flag(p) - flag * poly(v) in brake_bounds, soft with slack_brake
     * or hard when slack_brake[0] < 0; hard FoV rows at the braking end point, fov_term = (hfov, vfov) limits, vfov <= 0: no
     * vfov row), 2 also stability (hard velocity box +- vlim, terminal outputs flag * [p, q_e3] and flag * v.v: yrefN / WN
     * are then [B][5]).  brake_coeffs: 3-variate polynomial of degree brake_degree <= 5 in the term order of
     * utils/math.py:306-312.  The rows are carried by term_lift auxiliary nodes behind node N (csrc/term_lift.cuh), so
     * N + term_lift <= 31 and the state arrays of sdfnmpc_solver_state / _diag have N + term_lift (+ 1) nodes per instance
     * (sdfnmpc_solver_horizon).  Hard rows use an exact L1 penalty of weight z_hard (0 -> 1e4), which must exceed their
     * multipliers. */
    int32_t term_lift, brake_degree;
    double brake_coeffs[56];
    double brake_bounds[2], slack_brake[2], fov_term[2], vlim[3], z_hard;
} sdfnmpc_solver_desc;

/* Stage dimensions by model: att / att_tau / acc  nx 10, ny 11;  props  nx 14 (the reference's 13 states [p, q, v, w] + ONE PAD
 * entry that callers keep at 0: the QP's staging copies move 16-byte units), ny 16.  nu = 4 and nh = 3 for all.  Wherever a
 * comment below says [10] / [11] read nx / ny of the handle's model. */
int sdfnmpc_solver_dims(sdfnmpc_solver_t h, int32_t* nx, int32_t* ny, int32_t* nyN);
/* N of the OCP (stride of yref / W / qd) and the node stride N_alloc = N + term_lift of the handle's own arrays. */
int sdfnmpc_solver_horizon(sdfnmpc_solver_t h, int32_t* N, int32_t* N_alloc);

int sdfnmpc_solver_create(sdfnmpc_solver_t* out, const sdfnmpc_solver_desc* desc, sdfnmpc_sdf_t sdf);
int sdfnmpc_solver_destroy(sdfnmpc_solver_t h);
/* Ocp.init (ocp.py:144-149): x_k = x0 for all k, u_k = u_hover, for instances [0,B). */
int sdfnmpc_solver_init_traj(sdfnmpc_solver_t h, const double* x0_dev, int32_t B, void* stream);
/* Ocp.shift (ocp.py:152-156). */
int sdfnmpc_solver_shift(sdfnmpc_solver_t h, int32_t B, int32_t k, void* stream);
/* One SQP-RTI iteration for B instances (Ocp.solve, ocp.py:159-170).
 *   x0 [B][10]; yref [B][N][11]; W [B][N][11]; yrefN [B][4]; WN [B][4]; qd [B][N+1][4];
 *   pose [B][12] = W_p_Co(3), W_R_Co row-major(9); flag [B]; latents [B][L] float (NULL = keep the
 *   latents of the previous call); outputs u0 [B][4], status [B] (0 ok, 2 max-iter, 4 QP failure),
 *   qp_iters [B]. */
int sdfnmpc_solver_rti_step(sdfnmpc_solver_t h, int32_t B, const double* x0_dev, const double* yref_dev,
                            const double* W_dev, const double* yrefN_dev, const double* WN_dev,
                            const double* qd_dev, const double* pose_dev, const double* flag_dev,
                            const float* latents_dev, double* u0_dev, int32_t* status_dev,
                            int32_t* qp_iters_dev, void* stream);
/* Device pointers of the persistent warm start x [cap][N+1][10], u [cap][N][4] and of the QP
 * multipliers of the last step: lam_hard [cap][N][8] (lbu(4), ubu(4)), lam_soft [cap][N+1][6]. */
int sdfnmpc_solver_state(sdfnmpc_solver_t h, double** xbar_dev, double** ubar_dev, double** lam_hard_dev,
                         double** lam_soft_dev);
/* h values of the last linearisation [cap][N+1][3] and SDF query results d [cap][N+1], grad [cap][N+1][3]. */
int sdfnmpc_solver_diag(sdfnmpc_solver_t h, double** hval_dev, float** sdf_d_dev, float** sdf_g_dev);
/* Per instance [cap][2]: verified KKT residual of the last QP (inf-norm of input stationarity and dynamics) and
 * inf-norm of the full SQP step [du; dx] it produced (the full-SQP termination test of BatchedNmpc.solve_sqp). */
int sdfnmpc_solver_qp_residual(sdfnmpc_solver_t h, double** res_dev);
/* Per-kernel device timing of the next <= max_steps rti_step calls (CUDA events on the launching stream):
 * ms3 = summed milliseconds of {linearise, sdf, qp}; steps = number of timed calls. */
int sdfnmpc_solver_timing_begin(sdfnmpc_solver_t h, int32_t max_steps);
int sdfnmpc_solver_timing_end(sdfnmpc_solver_t h, float* ms3, int32_t* steps);
/* controller.py:74-81 bookkeeping for a batch: fail_count[b] = status[b] == 0 ? 0 : fail_count[b] + 1 (device pointers). */
int sdfnmpc_fail_count_update(const int32_t* status_dev, int64_t* fail_count_dev, int32_t B, void* stream);
/* RK4 plant step with the same model: x <- phi(x, u, dt) for closed-loop rollouts. */
int sdfnmpc_plant_step(sdfnmpc_solver_t h, int32_t B, double* x_dev, const double* u_dev, double dt, void* stream);
/* u_to_TRPYr + clip (controller.py:106-108): cmd [B][4]. */
int sdfnmpc_cmd_trpyr(sdfnmpc_solver_t h, int32_t B, const double* u_dev, double* cmd_dev, void* stream);
/* u_to_acc + clip (controller.py:101-103): cmd [B][4] = [W_R_B^T W_a, wz] from x0 [B][10] and u [B][4]; all models. */
int sdfnmpc_cmd_acc(sdfnmpc_solver_t h, int32_t B, const double* x0_dev, const double* u_dev, double* cmd_dev, void* stream);
/* u_to_props + clip (controller.py:27-29,111-113; quad_props.py:52): cmd [B][4] = rotor speeds u * limits.wp clipped to
 * [0, limits.wp]; model props only (SDFNMPC_E_INVALID otherwise). */
int sdfnmpc_cmd_props(sdfnmpc_solver_t h, int32_t B, const double* u_dev, double* cmd_dev, void* stream);

/* ------------------------------------------------------------------ collision labels ----- */
/* ColChecker.check_image_points (sdf_nmpc/utils/collision_checker.py:23-125): ground-truth collision label of each point
 * (Cartesian, sensor frame, metres) against the range / depth image it belongs to.  imgs [B][H][W] normalised by dmax;
 * p2i [n] image index per point or NULL (points are then split evenly, in order, over the B images);
 * outside: 0 free, 1 collision, 2 extrapolate (clamp to the frustum); out [n] 0 / 1.  All pointers are device pointers. */
int sdfnmpc_colcheck(const float* imgs_dev, int32_t B, int32_t H, int32_t W, const float* pts_dev, const int32_t* p2i_dev, int64_t n,
                     float dmax, float hfov, float vfov, float safe_ball, int32_t outside, int32_t is_depth, int32_t is_spherical,
                     uint8_t* out_dev, void* stream);

/* DfComputer.get_df (sdf_nmpc/utils/df_computer.py:56-221): ground-truth distance-field labels of points against the images.
 * Signed (get_sdf, :198-221): sign = the point's own collision label ('extrapolate', no safe ball); distance = the smallest
 * voxel offset |grid[g]| (grid_dist[g]) among the G voxels of the multi-resolution ball (generate_dist_grid, :31-53; built by
 * the host, sdf_nmpc_b200/collision.py) whose label differs from the point's; sdf = clamp(sign * dist, min_df, max_df);
 * grad [n][3] = -sign * grid[g*] / |grid[g*]|, zero where sdf saturates.  One CTA per point, nothing materialised. */
int sdfnmpc_df_signed(const float* imgs_dev, int32_t B, int32_t H, int32_t W, const float* pts_dev, const int32_t* p2i_dev, int64_t n,
                      const float* grid_dev, const float* grid_dist_dev, int32_t G, float dmax, float hfov, float vfov,
                      int32_t is_depth, int32_t is_spherical, float min_df, float max_df, float* sdf_dev, float* grad_dev,
                      void* stream);
/* Unsigned (get_udf, :138-176 with the Warp kernel :84-136): pool x pool min-pool of the image that ignores zeros (H, W
 * multiples of pool; pooled_dev [B][H/pool][W/pool] is caller-provided scratch), distance of each point to the 3-D position of
 * every pooled pixel or to the virtual wall at dmax, minimum over the pixels; udf = clamp(min, 0, max_df), grad = -direction. */
int sdfnmpc_df_unsigned(const float* imgs_dev, int32_t B, int32_t H, int32_t W, int32_t pool, float* pooled_dev, const float* pts_dev,
                        const int32_t* p2i_dev, int64_t n, float dmax, float hfov, float vfov, int32_t is_depth,
                        int32_t is_spherical, float max_df, float* udf_dev, float* grad_dev, void* stream);

/* ------------------------------------------------------------------ depth encoder ------- */
typedef struct sdfnmpc_encoder* sdfnmpc_encoder_t;

typedef struct {
    int32_t height, width;   /* sensor.shape_imgs[1:]                                       */
    int32_t size_latent;
    int32_t max_batch;
    int32_t is_depth;        /* apply Depth2Range (vae.py:21-23)                            */
    int32_t is_normalized;   /* skip ClipDistance (vae.py:18-20)                            */
    float dmax_scale;        /* dmax / mm_resolution * 1000 (preprocessing.py:75)           */
    float hfov, vfov;
} sdfnmpc_encoder_desc;

/* Weights: host pointers, torch layouts.  conv weights [Cout][Cin][kh][kw]; BatchNorm as
 * (weight, bias, running_mean, running_var) with eps; folded into the convolutions at creation. */
typedef struct {
    const float *conv0_w, *conv0_b;                 /* 64x1x7x7 (+bias)                     */
    const float* conv1_w[4]; const float* bn1[4][4];
    const float* conv2_w[4]; const float* bn2[4][4];
    const float* sc_w[4];    const float* sc_bn[4][4];   /* NULL for the stride-1 block     */
    const float *mean_w, *mean_b;                   /* [L][2048], [L]                       */
    float bn_eps;
} sdfnmpc_encoder_weights;

int sdfnmpc_encoder_create(sdfnmpc_encoder_t* out, const sdfnmpc_encoder_desc* desc, const sdfnmpc_encoder_weights* w);
int sdfnmpc_encoder_destroy(sdfnmpc_encoder_t h);
/* imgs [B][H][W] float32 raw sensor values -> latents [B][L] (preprocessing fused). */
int sdfnmpc_encoder_encode(sdfnmpc_encoder_t h, const float* imgs_dev, int32_t B, float* latents_dev, void* stream);
int sdfnmpc_encoder_encode_host(sdfnmpc_encoder_t h, const float* imgs, int32_t B, float* latents);

/* ------------------------------------------------------------------ VAE decoder --------- */
/* VaeWrapper.decode (sdf_nmpc/vae.py:42-45): latent -> reconstructed range image in [0, 1]
 * (network/vae.py:63-90: Linear + ELU, 4 x ResBlockDeconv of network/resnet.py:59-111, ConvTranspose2d 5x5, bilinear
 * Upsample to (height, width), Sigmoid).  The feature map after the Linear layer is the reference's fixed [512][8][15]. */
typedef struct sdfnmpc_decoder* sdfnmpc_decoder_t;

typedef struct {
    int32_t height, width;   /* Upsample size = shape_imgs                                  */
    int32_t size_latent;
    int32_t max_batch;
} sdfnmpc_decoder_desc;

/* Weights: host pointers, torch layouts.  ConvTranspose2d weights [Cin][Cout][kh][kw]; block i maps C_i = 512 >> i
 * channels to C_i / 2: ct1 (3x3, stride 2, padding 1, output_padding 1), ct2 (3x3, stride 1, padding 1), shortcut
 * sc (1x1, stride 2, output_padding 1); BatchNorm as (weight, bias, running_mean, running_var), folded at creation. */
typedef struct {
    const float *lin_w, *lin_b;                     /* [512*8*15][L], [512*8*15]            */
    const float* ct1_w[4]; const float* bn1[4][4];
    const float* ct2_w[4]; const float* bn2[4][4];
    const float* sc_w[4];  const float* sc_bn[4][4];
    const float *out_w, *out_b;                     /* [32][1][5][5], [1]                   */
    float bn_eps;
} sdfnmpc_decoder_weights;

int sdfnmpc_decoder_create(sdfnmpc_decoder_t* out, const sdfnmpc_decoder_desc* desc, const sdfnmpc_decoder_weights* w);
int sdfnmpc_decoder_destroy(sdfnmpc_decoder_t h);
/* latents [B][L] float32 -> imgs [B][height][width] float32 in [0, 1] (normalised by dmax, like the encoder's input). */
int sdfnmpc_decoder_decode(sdfnmpc_decoder_t h, const float* latents_dev, int32_t B, float* imgs_dev, void* stream);
int sdfnmpc_decoder_decode_host(sdfnmpc_decoder_t h, const float* latents, int32_t B, float* imgs);

/* ------------------------------------------------------------------ CasADi external ABI - */
/* Bind the process-global network the `sdf_l4c` symbols evaluate (L4CasADi bakes the model path
 * into the generated library; here it is bound at run time, or loaded from the weight blob named
 * by the environment variable SDFNMPC_L4C_BLOB on first use). */
int sdfnmpc_l4c_bind(sdfnmpc_sdf_t h);
int sdfnmpc_l4c_load_blob(const char* path);

typedef long long int casadi_int;
#define SDFNMPC_DECLARE_CASADI(F)                                                              \
    int F(const double** arg, double** res, casadi_int* iw, double* w, int mem);               \
    int F##_alloc_mem(void);                                                                   \
    int F##_init_mem(int mem);                                                                 \
    void F##_free_mem(int mem);                                                                \
    int F##_checkout(void);                                                                    \
    void F##_release(int mem);                                                                 \
    void F##_incref(void);                                                                     \
    void F##_decref(void);                                                                     \
    casadi_int F##_n_in(void);                                                                 \
    casadi_int F##_n_out(void);                                                                \
    double F##_default_in(casadi_int i);                                                       \
    const char* F##_name_in(casadi_int i);                                                     \
    const char* F##_name_out(casadi_int i);                                                    \
    const casadi_int* F##_sparsity_in(casadi_int i);                                           \
    const casadi_int* F##_sparsity_out(casadi_int i);                                          \
    int F##_work(casadi_int* sz_arg, casadi_int* sz_res, casadi_int* sz_iw, casadi_int* sz_w);
/* sdf_l4c: i0 (3+L)x1 dense -> o0 1x1.  jac_sdf_l4c: (i0, out_o0 [may be NULL]) -> jac_o0_i0 1x(3+L) dense. */
SDFNMPC_DECLARE_CASADI(sdf_l4c)
SDFNMPC_DECLARE_CASADI(jac_sdf_l4c)

#ifdef __cplusplus
}
#endif
#endif /* SDFNMPC_B200_H */
