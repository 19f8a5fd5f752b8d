// Batched OCP-QP solver, tensor-core form (device only): the interior-point method of qp_core.cuh - same formulation, same
// iterates up to rounding - with every Riccati sweep running on the fp64 tensor pipe (mma.sync m8n8k4 f64, SASS DMMA.8x8x4)
// and the recursion's matrices living in the register file.  qp_core.cuh stays the scalar statement of the algorithm (host
// emulation, -DSDFNMPC_QP_SCALAR builds); this file is what libsdfnmpc_b200.so runs for the reference's stage shape
// (nx 10, nu 4, ny 11, nh 3: models att / att_tau / acc).
//
// Role in the reference: the QP acados hands to HPIPM at every SQP-RTI iteration (sdf_nmpc/ocp.py:109-120).
//
// Layout idea.  The augmented stage variable w~ = [u; x; 1] (the affine terms ride along as one more column) lives on 16
// POSITIONS: 0..7 = x0..x7, 8 / 10 / 12 / 14 = x8, x9, the affine entry, a zero pad, 9 / 11 / 13 / 15 = u0..u3; the contraction
// index of the dynamics lives on 12 SLOTS (slot s < 10 = x_s, 10 = affine, 11 = pad; slot s sits at position s for s < 8 and at
// 8 + 2 (s - 8) above).  With the m8n8k4 fragment layout (lane 4 g + q holds A[g][q], B[q][g], D[g][2q], D[g][2q+1]) this
// ordering makes the accumulator fragment of one product, register for register, an operand fragment of the next:
//   X' = T' P~        T = [A B | phi rd; 0 1]~ (12 slots x 16 positions),  P~ = [P p; p' .] read by columns (symmetric)
//   M~ = H~ + X' T    H~ = [H rv; rv' .] + this iteration's row-elimination terms
//   K~ = -Muu^-1 Mu.  LDL' of the 4 x 4 input block in every lane, one row per quad
//   P~ <- M~ + M.u K~
// = 12 + 12 + 4 DMMA per stage, no shared-memory round trip and no barrier inside a sweep.  Roll-outs and the corrector's
// right-hand-side sweep are chains of matrix-vector DMMAs (vector in column 0 of the B operand, result in column 0 of the
// accumulators, a few shuffles put it back into operand position).  Stage operands come from L2 through a register pipeline
// several stages deep (`pipelined_stages`).
//
// Measured on B200 (tools/dmma_rate.cu): DMMA.8x8x4 26 cycles dependent, one per 16 cycles and SM sub-partition - the same
// 16 MAC / cycle as DFMA, in 1 / 8 of the issue slots and with register operands.
#pragma once
#include "qp_core.cuh"

#if defined(__CUDACC__)
namespace sdfnmpc {

struct QpMmaShape {
    static constexpr int NX = 10, NU = 4, NW = 14, NY = 11, NH = 3;
    static constexpr int NCORR = NU + 6 + NU + 3;   // du, Dp (sym 3x3), gu, gp
    static constexpr int NPROD = 2 * NU + 4 * NH;
    static constexpr int NROW = 4 * NU + 8 * NH;    // parked doubles per node: (t, lam) of 8 hard rows, (t, lam, s, mu) of 6 soft rows
    static constexpr int HP = 256;                  // H~_k, 16 x 16 by position (row-major)
    static constexpr int ABT = 192;                 // T_k transposed: [16 positions][12 slots] - written by the linearisation kernel (LinOut::ABt)
    static constexpr int KF = 64;                   // K~_k as fragments: [tile][lane] = K~[q][position 8 tile + g]
    static constexpr int LF = 16;                   // LDL' of Muu_k: inverse pivots [4] | unit lower factor by columns [6]
    static constexpr int GP = 16;                   // cost gradient by position
#ifndef QM_STATE_SMEM
#define QM_STATE_SMEM 1
#endif
#if QM_STATE_SMEM && !defined(QM_KF_GLOBAL)
#define QM_KF_GLOBAL
#endif
    // row state (t, lam, s, mu of every row, NROW doubles per node) in shared memory, gains K~ in the slot's L2 scratch (1), or the
    // round-2a arrangement: gains in shared memory, row state parked in L2 between the row phases (0)
    static constexpr bool STATE_SHARED = QM_STATE_SMEM != 0;
#if defined(QM_KF_GLOBAL)
    static constexpr bool KF_SHARED = false;        // gains in the slot's global scratch (L2): 12.8 KB of shared memory per warp
#else
    static constexpr bool KF_SHARED = true;         // gains in shared memory: 22.8 KB per warp
#endif
    SDF_HD static int kp(int N) { return N + 2; }
    SDF_HD static constexpr int wpos(int i) { return i < NU ? 9 + 2 * i : (i < NU + 8 ? i - NU : 8 + 2 * (i - NU - 8)); }
    // doubles of global scratch per warp slot
    SDF_HD static size_t work_doubles(int N) {
        const size_t n1 = (size_t)N + 1;
        return n1 * HP + (size_t)N * LF + n1 * GP + ((n1 * NH * 3 + 1) & ~(size_t)1) + (STATE_SHARED ? 1 : 2) * (size_t)NROW * kp(N) +
               (KF_SHARED ? 0 : (size_t)N * KF);
    }
    // doubles of shared memory per warp: iterate, step, elimination terms (+ a zero row), second-order products, corrector
    // feed-forward, costate ping-pong of the verification, and the gains K~_k of every stage (written by the factorising sweep,
    // read by three more sweeps per iteration: 10 KB that would otherwise cycle through L2)
    SDF_HD static int smem_doubles(int N) {
        return (2 * NW + NCORR + 1 + NPROD + NU) * kp(N) + 2 * 16 + (KF_SHARED ? N * KF : 0) + (STATE_SHARED ? NROW * kp(N) : 0);
    }
};

// Lane-strided loops in this file run a WARP-UNIFORM number of trips with the tail predicated (`QM_LANES(e, n)`): a loop whose
// lanes leave at different trips is a divergent exit, and on this toolchain (nvcc 12.9, sm_100a) the lanes of such a loop -
// the partially unrolled scatter of [A B], 2800 entries = 87.5 trips - did NOT re-join behind it: the warp went on as two
// halves, the compiler (which assumes they re-join) had removed the __syncwarp() that followed, and every mma.sync behind it
// multiplied half-warp fragments only (found with __activemask() probes; -DQM_ASSERT_CONVERGED keeps them as traps).
// read-once inputs (J, residuals, constraint rows) are loaded with the evict-first policy: the slots' scratch and operand
// blocks (97 KB x 1184 slots) fill the 126 MB L2 by themselves, and 36 KB per instance of streaming input evicts them otherwise
#define QM_LDS(p_) __ldcs(p_)
#define QM_LANES(v, n) for (int v##_0 = 0; v##_0 < (n); v##_0 += 32) if (const int v = v##_0 + lane; v < (n))
// register-pipeline depths of the corrector's sweep and the roll-outs (A/B knobs; tools/qp_ab.py)
#ifndef QM_D_BACK
#define QM_D_BACK 3
#endif
#ifndef QM_D_ROLL
#define QM_D_ROLL 2
#endif
#ifndef QM_D_FAC
#define QM_D_FAC 2
#endif
#ifndef QM_D_HES
#define QM_D_HES 2
#endif
// row phases of the iteration as rolled loops over the rows (1) or with all rows unrolled and the state in registers (0)
#ifndef QM_ROLLED_ROWS
#define QM_ROLLED_ROWS 1
#endif
// (rolling the two sides of a row as well: 2.57 vs 2.50 ms - the trip gets too short to cover its own latencies)
#if defined(QM_ASSERT_CONVERGED)
#define QM_CONVERGED() { if (__activemask() != 0xffffffffu) __trap(); }
#else
#define QM_CONVERGED()
#endif

namespace qpdetail {

struct Acc2 { double x, y; };
__device__ __forceinline__ void dmma(Acc2& d, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d.x), "+d"(d.y) : "d"(a), "d"(b));
}

// Stage loop with the operands of the next D - 1 stages in flight IN REGISTERS: D operand sets, the loop unrolled by D so
// that every set keeps its own registers from the load to its use D stages later (a rotating copy o0 = o1 waits for the load
// at the copy: measured, a stage then takes one full L2 latency, ~1300 cycles, whatever its arithmetic).
template <int D, class Ops, class Ld, class St>
__device__ __forceinline__ void pipelined_stages(int n, Ld&& ld, St&& stage) {
    Ops o[D];
#pragma unroll
    for (int d = 0; d < D; ++d) o[d] = ld(d < n ? d : n - 1);
    int i = 0;
#pragma unroll 1
    for (; i + D <= n; i += D) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            stage(i + d, o[d]);
            const int nx = i + d + D;
            o[d] = ld(nx < n ? nx : n - 1);
        }
    }
#pragma unroll
    for (int d = 0; d < D - 1; ++d)
        if (i + d < n) stage(i + d, o[d]);
}

}  // namespace qpdetail

// LIFT: terminal rows lifted onto auxiliary nodes behind node P.n_real (term_lift.cuh): per-node bounds / slack weights P.nlh ...
// P.nZs, cost rows and state regularisation end at node n_real.  false compiles to the plain OCP with no trace of it.
template <int NFIX = 0, bool LIFT = false>
__device__ void qp_solve_mma(const Lanes& L, const QpProb<10, 4, 11, 3>& P, double* sm) {
    using namespace qpdetail;
    using S = QpMmaShape;
    constexpr int NX = S::NX, NU = S::NU, NW = S::NW, NY = S::NY, NH = S::NH, NCORR = S::NCORR, NPROD = S::NPROD;
    constexpr int HC0 = NU;                       // position block of the constraint rows = first three states
    const int N = NFIX > 0 ? NFIX : P.N;
    const int k0 = P.k0, KP = S::kp(N);
    const int NR = LIFT ? P.n_real : N;          // terminal node of the OCP
    const int lane = L.lane, g_ = lane >> 2, q_ = lane & 3;
    const int ncomp = N * 2 * NU + 2 * (N + 1 - k0) * 2 * NH;
    constexpr unsigned FULL = 0xffffffffu;

    // ---- global scratch (one block per resident warp) ----
    double* p = P.work;
    auto take = [&](size_t n) { double* r = p; p += n; return r; };
    const size_t n1 = (size_t)N + 1;
    double* Hp = take(n1 * S::HP);
    const double* ABt = P.ABt;                    // [A B | b]~ by (position, slot), straight from the linearisation kernel
    double* Lf = take((size_t)N * S::LF);
    double* gpos = take(n1 * S::GP);
    double* cpos = take((n1 * NH * 3 + 1) & ~(size_t)1);
    double* TB = S::STATE_SHARED ? nullptr : take((size_t)S::NROW * KP);      // row state (shared memory when STATE_SHARED: set below)
    double* TD = take((size_t)S::NROW * KP);                                   // row steps

    // ---- shared workspace ----
    double* Wsm = sm;                             // [NW][KP]   iterate w
    double* dW = Wsm + NW * KP;                   // [NW][KP]   step
    double* corr = dW + NW * KP;                  // [NCORR + 1][KP] du(NU) Dp(6) gu(NU) gp(3) per node, then a zero row
    double* prod = corr + (NCORR + 1) * KP;       // [NPROD][KP] affine second-order products
    double* ffd = prod + NPROD * KP;              // [NU][KP]   input residual / feed-forward of the corrector's sweep
    double* pp = ffd + NU * KP;                   // [2][16]    costate ping-pong of the verification
    double* Kf;                                   // [N][2][32] gains as fragments
    if constexpr (S::KF_SHARED) Kf = pp + 2 * 16; else Kf = take((size_t)N * S::KF);
    if constexpr (S::STATE_SHARED) TB = pp + 2 * 16 + (S::KF_SHARED ? N * S::KF : 0);
#if defined(QM_KF_GLOBAL) && !defined(QM_KF_L1)
#define QM_KLD(p_) __ldcg(p_)
#define QM_KST(p_, v_) __stcg(p_, v_)
#else
#define QM_KLD(p_) (*(p_))
#define QM_KST(p_, v_) (*(p_) = (v_))
#endif

    auto scale = [&](int k) { return k < N ? P.dt[k] : 1.0; };
    auto wrow = [&](int k, int r) { return k < NR ? P.W[k * NY + r] : ((k == NR && r < P.nyN) ? P.WN[r] : 0.0); };
    auto lh_ = [&](int kk, int r) { return LIFT ? P.nlh[kk * NH + r] : P.lh[r]; };
    auto uh_ = [&](int kk, int r) { return LIFT ? P.nuh[kk * NH + r] : P.uh[r]; };
    auto zs_ = [&](int kk, int r) { return LIFT ? P.nzs[kk * NH + r] : P.zs[r]; };
    auto Zs_ = [&](int kk, int r) { return LIFT ? P.nZs[kk * NH + r] : P.Zs[r]; };
    auto c_pos = [&](int kk, int r, int c) { return cpos[(kk * NH + r) * 3 + c]; };
    // position -> index into w = [u; x] (14 = affine, 15 = pad)
    auto widx = [](int pos) { return pos < 8 ? NU + pos : ((pos & 1) ? (pos - 9) >> 1 : (pos == 8 ? NU + 8 : (pos == 10 ? NU + 9 : (pos == 12 ? 14 : 15)))); };

#if defined(SDFNMPC_QP_TRACE)
    // development: cycles per phase of IPM iteration 2 ([pass][8]: elimination, backward sweep, roll-out, row steps) and of
    // the once-per-solve parts ([16..]: prologue operands, Hessian build, initial point, verification + outputs)
    long long qpt_last = clock64();
#define QPT(slot, cond)                                                        \
    {                                                                          \
        if (P.trace && lane == 0) {                                            \
            const long long now_ = clock64();                                  \
            if (cond) P.trace[slot] += now_ - qpt_last;                        \
            qpt_last = now_;                                                   \
        }                                                                      \
    }
#else
#define QPT(slot, cond)
#endif
    // =====================================================================================================
    // prologue: SDF row of C / h, operands of the sweeps
    // =====================================================================================================
    const bool sdf_cost = P.sdf_cost_w > 0.0;     // the cost row also needs the SDF at node 0
    QM_LANES(k, N + 1) {
        if (k >= (sdf_cost ? 0 : k0) && k <= NR) {
            const double* R = P.pose + 3;
            const double gx = P.sdf_g[k * 3 + 0], gy = P.sdf_g[k * 3 + 1], gz = P.sdf_g[k * 3 + 2];
            double* Cr = P.C + ((size_t)k * NH + (NH - 1)) * NW;
            for (int jj = 0; jj < 3; ++jj) Cr[HC0 + jj] = P.flag * (R[jj * 3 + 0] * gx + R[jj * 3 + 1] * gy + R[jj * 3 + 2] * gz);
            P.hval[k * NH + (NH - 1)] = P.flag * (double)P.sdf_d[k] + (1.0 - P.flag) * P.max_df;
        }
        for (int r = 0; r < NH; ++r)
            for (int c = 0; c < 3; ++c) cpos[(k * NH + r) * 3 + c] = QM_LDS(P.C + ((size_t)k * NH + r) * NW + HC0 + c);
    }
    // T_k = [A B | b; 0 1]~ by (position, slot) comes ready-made from the linearisation kernel.  Node 0 is linearised at the measured
    // state (ocp.py:161), so dx_0 = x0 - xbar_0 = 0: the iterate starts at w = 0 and the dynamics defect IS the affine column.
    QM_LANES(e, (N + 1) * S::GP) __stcg(gpos + e, 0.0);
    QM_LANES(e, NW * KP) { Wsm[e] = 0.0; dW[e] = 0.0; }
    QM_LANES(e, KP) corr[NCORR * KP + e] = 0.0;
    QPT(16, true)
    // H~_k = J~' (W dt) J~ + lm dt I with J~ = [J | y - yref]: 12 DMMA per node straight from the linearisation's J; the affine
    // column comes out as the cost gradient.
    // mma.sync.aligned needs the WHOLE warp at the instruction: see QM_LANES above for why every loop before this point has a
    // warp-uniform trip count.
    L.sync();
    {
        struct HOps { double v[3][2], w[3]; };             // J~ entries and their row weights
        // the lane's operand entries: J~[row 4 s + q][position 8 t + g]
        const int wi0 = widx(g_), wi1 = widx(8 + g_);
        auto ld_ops = [&](int k) {
            // every load is unconditional (rows past NY are clamped and carry weight 0, the pad position is masked): nothing
            // here is predicated
            HOps o;
            const double* Jk = P.J + (size_t)k * NY * NW;
            const double* yk = P.yres + (size_t)k * NY;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int r = 4 * s + q_;
                const int rc = r < NY ? r : NY - 1;
                // loads only: any arithmetic on the loaded values here would wait for them and undo the pipelining
                o.w[s] = wrow(k, rc);
                o.v[s][0] = QM_LDS(Jk + rc * NW + wi0);
                const double* src = wi1 < NW ? Jk + rc * NW + wi1 : yk + rc;
                o.v[s][1] = QM_LDS(src);
            }
            return o;
        };
        auto stage = [&](int k, const HOps& o) {
            L.sync();
            QM_CONVERGED()
            // accumulators start from the Levenberg-Marquardt diagonal (and the optional SDF cost row), then J~' (W dt) J~
            Acc2 H[2][2];
            const double lmk = P.lm * scale(k);
            double cr = 0.0, cJ[3] = {0.0, 0.0, 0.0}, wk = 0.0;
            if (sdf_cost && k < NR) {       // optional cost row r = (1 - 0.5 d)^4, position block only (gen_model.py:65-66)
                const double qq = 1.0 - 0.5 * P.hval[k * NH + (NH - 1)];
                const double* Cr = P.C + ((size_t)k * NH + (NH - 1)) * NW;
                cr = qq * qq * qq * qq;
                for (int c = 0; c < 3; ++c) cJ[c] = -2.0 * qq * qq * qq * Cr[HC0 + c];
                wk = P.sdf_cost_w * scale(k);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int row = 8 * (e >> 2) + g_, col = 8 * ((e >> 1) & 1) + 2 * q_ + (e & 1);
                // auxiliary nodes (LIFT): inputs only (positions 9, 11, 13, 15), their state is not regularised
                double v = (row == col && row != 12 && row != 14 && (!LIFT || k <= NR || (row >= 9 && (row & 1)))) ? lmk : 0.0;
                if (sdf_cost) {
                    const double jr = row == 0 ? cJ[0] : (row == 1 ? cJ[1] : cJ[2]), jc = col == 0 ? cJ[0] : (col == 1 ? cJ[1] : cJ[2]);
                    if (row < 3 && col < 3) v += wk * jr * jc;
                    if (row < 3 && col == 12) v += wk * cr * jr;
                    if (col < 3 && row == 12) v += wk * cr * jc;
                }
                if (e & 1) H[e >> 2][(e >> 1) & 1].y = v; else H[e >> 2][(e >> 1) & 1].x = v;
            }
            // rows past NY carry weight 0, the pad position (15 in w~) is masked
            double a[3][2], v1[3][2];
            const double sck = scale(k), m1 = wi1 == 15 ? 0.0 : 1.0;
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const double w = (4 * s + q_ < NY) ? o.w[s] * sck : 0.0;
                v1[s][0] = o.v[s][0]; v1[s][1] = o.v[s][1] * m1;
                a[s][0] = v1[s][0] * w; a[s][1] = v1[s][1] * w;
            }
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                    for (int s = 0; s < 3; ++s) dmma(H[mt][nt], a[s][mt], v1[s][nt]);
            double* hb = Hp + (size_t)k * S::HP + g_ * 16 + 2 * q_;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 2; ++nt)
                    __stcg(reinterpret_cast<double2*>(hb + mt * 128 + nt * 8), make_double2(H[mt][nt].x, H[mt][nt].y));
            // cost gradient (affine column, position 12 = tile 1, q 2, register x)
            if (q_ == 2) { __stcg(gpos + k * S::GP + g_, H[0][1].x); __stcg(gpos + k * S::GP + 8 + g_, H[1][1].x); }
        };
        pipelined_stages<QM_D_HES, HOps>(N + 1, ld_ops, stage);
    }
    L.sync();

    QPT(17, true)
    // =====================================================================================================
    // per-lane node state (lane k <-> node k), parked in L2 between the row phases ([item][node]: coalesced)
    // =====================================================================================================
    double tb[2 * NU], lb[2 * NU], ts[2 * NH], ls[2 * NH], ss[2 * NH], ms[2 * NH];
    double dtb[2 * NU], dlb[2 * NU], dts[2 * NH], dls[2 * NH], dss[2 * NH], dms[2 * NH];
    auto hard_d = [&](int kk, int a) {
        const int j = a < NU ? a : a - NU;
        const double ub = P.ubar[kk * NU + j];
        return a < NU ? P.lbu[j] - ub : ub - P.ubu[j];
    };
    auto soft_d = [&](int kk, int a) {
        const int r = a < NH ? a : a - NH;
        const double hv = P.hval[kk * NH + r];
        return a < NH ? lh_(kk, r) - hv : hv - uh_(kk, r);
    };
#define QM_LD(p_) __ldcg(p_)
#define QM_ST(p_, v_) __stcg(p_, v_)
#if QM_STATE_SMEM
#define QM_TB_LD(p_) (*(p_))
#define QM_TB_ST(p_, v_) (*(p_) = (v_))
#else
#define QM_TB_LD(p_) __ldcg(p_)
#define QM_TB_ST(p_, v_) __stcg(p_, v_)
#endif
#define QM_LOAD(LD_, BASE, A, B, C4, D4, E4, F4, kk)                                                                      \
    {                                                                                                                \
        const double* q__ = BASE + (kk);                                                                             \
        _Pragma("unroll") for (int a = 0; a < 2 * NU; ++a) { A[a] = LD_(q__ + a * KP); B[a] = LD_(q__ + (2 * NU + a) * KP); } \
        _Pragma("unroll") for (int a = 0; a < 2 * NH; ++a) {                                                         \
            C4[a] = LD_(q__ + (4 * NU + a) * KP); D4[a] = LD_(q__ + (4 * NU + 2 * NH + a) * KP);                   \
            E4[a] = LD_(q__ + (4 * NU + 4 * NH + a) * KP); F4[a] = LD_(q__ + (4 * NU + 6 * NH + a) * KP);          \
        }                                                                                                            \
    }
#define QM_STORE(ST_, BASE, A, B, C4, D4, E4, F4, kk)                                                                     \
    {                                                                                                                \
        double* q__ = BASE + (kk);                                                                                   \
        _Pragma("unroll") for (int a = 0; a < 2 * NU; ++a) { ST_(q__ + a * KP, A[a]); ST_(q__ + (2 * NU + a) * KP, B[a]); } \
        _Pragma("unroll") for (int a = 0; a < 2 * NH; ++a) {                                                         \
            ST_(q__ + (4 * NU + a) * KP, C4[a]); ST_(q__ + (4 * NU + 2 * NH + a) * KP, D4[a]);                    \
            ST_(q__ + (4 * NU + 4 * NH + a) * KP, E4[a]); ST_(q__ + (4 * NU + 6 * NH + a) * KP, F4[a]);            \
        }                                                                                                            \
    }
#define QM_LOAD_STATE(kk) QM_LOAD(QM_TB_LD, TB, tb, lb, ts, ls, ss, ms, kk)
#define QM_STORE_STATE(kk) QM_STORE(QM_TB_ST, TB, tb, lb, ts, ls, ss, ms, kk)
#define QM_LOAD_DELTA(kk) QM_LOAD(QM_LD, TD, dtb, dlb, dts, dls, dss, dms, kk)
#define QM_STORE_DELTA(kk) QM_STORE(QM_ST, TD, dtb, dlb, dts, dls, dss, dms, kk)
#define QM_NODE_LOOP(kk) QM_LANES(kk, N + 1)

    // ---- initial point and initial residuals (w = 0 except dx_0, costates 0) ----
    double res_v0 = 0.0;
    QM_NODE_LOOP(kk) {
#pragma unroll
        for (int a = 0; a < 2 * NU; ++a) {
            double t = 1.0, l = 1.0;
            if (kk < N) {
                t = fmax(-hard_d(kk, a), SDFNMPC_QP_T0);
                l = fmax(SDFNMPC_QP_MU0 / t, 1e-3 * SDFNMPC_QP_MU0);
            }
            tb[a] = t; lb[a] = l;
        }
#pragma unroll
        for (int a = 0; a < 2 * NH; ++a) {
            const int r = a < NH ? a : a - NH;
            double t = 1.0, l = 1.0, m = 1.0, s0 = SDFNMPC_QP_S0;
            if (kk >= k0) {
                const double z = zs_(kk, r) * scale(kk), Z = Zs_(kk, r) * scale(kk);
                double cw = 0.0;
                for (int c = 0; c < 3; ++c) cw += c_pos(kk, r, c) * Wsm[(HC0 + c) * KP + kk];
                const double gw = a < NH ? cw : -cw;
                t = fmax(gw - soft_d(kk, a) + s0, SDFNMPC_QP_T0);
                l = fmax(0.5 * (z + Z * s0), 1e-3);
                m = l;
                // lifted rows (exact-penalty weight z_hard): row multiplier capped, the rest to the slack's multiplier (qp_core.cuh)
                constexpr double lcap = LIFT ? SDFNMPC_QP_LCAP_LIFT : SDFNMPC_QP_LCAP;
                if ((LIFT || SDFNMPC_QP_LCAP_ALL) && l > lcap) { l = lcap; m = z + Z * s0 - l; }
            }
            ts[a] = t; ss[a] = s0; ls[a] = l; ms[a] = m;
        }
        QM_STORE_STATE(kk);
        // rv0_k = H_k w_k + g_k - sum lam_i gvec_i -> affine row / column of H~_k (inputs, then states: no dynamically
        // indexed register arrays)
        double* hk = Hp + (size_t)kk * S::HP;
#pragma unroll
        for (int i = 0; i < NU; ++i) {
            const int pi = S::wpos(i);
            double s = __ldcg(gpos + kk * S::GP + pi);
            if (kk < N) s -= lb[i] - lb[NU + i];
            if (kk == N) s = 0.0;
            __stcg(hk + pi * 16 + 12, s); __stcg(hk + 12 * 16 + pi, s);
            res_v0 = fmax(res_v0, fabs(s));
        }
        const double lsd[NH] = {ls[0] - ls[NH], ls[1] - ls[NH + 1], ls[2] - ls[NH + 2]};
#pragma unroll 1
        for (int i = NU; i < NW; ++i) {
            const int pi = S::wpos(i);
            double s = __ldcg(gpos + kk * S::GP + pi);
            if (kk >= k0 && i < HC0 + 3) s -= lsd[0] * c_pos(kk, 0, i - HC0) + lsd[1] * c_pos(kk, 1, i - HC0) + lsd[2] * c_pos(kk, 2, i - HC0);
            if (kk == 0) s = 0.0;
            __stcg(hk + pi * 16 + 12, s); __stcg(hk + 12 * 16 + pi, s);
            res_v0 = fmax(res_v0, fabs(s));
        }
        // the dynamics defect b_{k-1} (affine column of T_{k-1}) only enters the residual norm here
        if (kk >= 1) {
#pragma unroll 1
            for (int i = 0; i < NX; ++i) res_v0 = fmax(res_v0, fabs(P.bdyn[(kk - 1) * NX + i]));
        }
    }
    res_v0 = L.max(res_v0);
    L.sync();

    // per-lane metadata of the factorising sweep: which row of `corr` each of the lane's 8 accumulator registers receives
    // (NCORR = the zero row) and whether it belongs to the affine row / column (scaled by phi); register e = (ri, rj, c)
    // holds the entry at position (8 ri + g, 8 rj + 2 q + c)
    unsigned long long mm_ci = 0;
    unsigned mm_aff = 0;
    {
        auto upos = [](int p_) { return (p_ >= 9 && (p_ & 1)) ? (p_ - 9) >> 1 : -1; };
        for (int e = 0; e < 8; ++e) {
            const int row = 8 * (e >> 2) + g_, col = 8 * ((e >> 1) & 1) + 2 * q_ + (e & 1);
            int ci = NCORR;
            if (row == col && upos(row) >= 0) ci = upos(row);
            else if (row < 3 && col < 3) ci = NU + sym3(row, col);
            else if ((row == 12) != (col == 12)) {
                const int o = row == 12 ? col : row;
                if (upos(o) >= 0) ci = NU + 6 + upos(o);
                else if (o < 3) ci = 2 * NU + 6 + o;
            }
            mm_ci |= (unsigned long long)ci << (5 * e);
            if (row == 12 || col == 12) mm_aff |= 1u << e;
        }
    }

    // stage operands in the lane layouts of the three kinds of sweeps
    auto ld_h = [&](int kk, Acc2 (&h)[2][2]) {                       // H~ tiles as accumulators
        const double* b_ = Hp + (size_t)kk * S::HP + g_ * 16 + 2 * q_;
#pragma unroll
        for (int ri = 0; ri < 2; ++ri)
#pragma unroll
            for (int rj = 0; rj < 2; ++rj) {
                const double2 v = __ldcg(reinterpret_cast<const double2*>(b_ + ri * 128 + rj * 8));
                h[ri][rj].x = v.x; h[ri][rj].y = v.y;
            }
    };
    struct TOps { double f0[2], f1[2], f2[2]; };                      // T'[position 8 r + g][slots 2q, 2q+1, 8+q]
    auto ld_t = [&](int kk) {
        TOps o;
        const double* b_ = ABt + (size_t)kk * S::ABT + g_ * 12;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const double2 v = __ldcg(reinterpret_cast<const double2*>(b_ + r * 96 + 2 * q_));
            o.f0[r] = v.x; o.f1[r] = v.y;
            o.f2[r] = __ldcg(b_ + r * 96 + 8 + q_);
        }
        return o;
    };
    auto inject = [&](Acc2 (&h)[2][2], int kk, double phi_) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int ci = (int)((mm_ci >> (5 * e)) & 31u);
            const double f = ((mm_aff >> e) & 1u) ? phi_ : 1.0;
            double& v = (e & 1) ? h[e >> 2][(e >> 1) & 1].y : h[e >> 2][(e >> 1) & 1].x;
            v = fma(v, f, corr[ci * KP + kk]);
        }
    };

    QPT(18, true)
    int status = 2, it = 0;
    bool safe = false;              // conservative mode after a stall: more centring, no second-order term
    double cm_prev1 = 1e300, cm_prev2 = 1e300;
    double phi = 1.0;               // product of (1 - alpha): scale of the stationarity / dynamics residuals
    double res_track = 0.0;
    for (;; ++it) {
        double tau = 0.0, mubar = 0.0;
        bool stop = false;
        for (int pass = 0; pass < 2 && !stop; ++pass) {
            // =========== row elimination (lane = node): corrections of the stage Hessian / rhs ===========
            double res_rows = 0.0, cmax = 0.0, csum = 0.0;
#if QM_ROLLED_ROWS
            // Rolled form: one trip per input (both bounds) / per constraint row (both sides), the row state read from its parked
            // place inside the trip.  A quarter of the unrolled form's code (which held all 40 state doubles in registers and was
            // 2000 straight-line instructions per phase): these phases were half instruction-fetch stalls.
            QM_NODE_LOOP(kk) {
                const double* sb = TB + kk;
#pragma unroll 1
                for (int j = 0; j < NU; ++j) {
                    double thn = 0.0, cn = 0.0;
                    if (kk < N) {
                        const double wu = Wsm[j * KP + kk];
                        const double ub = P.ubar[kk * NU + j];
#pragma unroll
                        for (int side = 0; side < 2; ++side) {
                            const int a = side * NU + j;
                            const double sgn = side == 0 ? 1.0 : -1.0;
                            const double t_ = QM_TB_LD(sb + a * KP), l_ = QM_TB_LD(sb + (2 * NU + a) * KP);
                            const double cr = pass == 0 ? 0.0 : tau - prod[a * KP + kk];
                            const double d_ = side == 0 ? P.lbu[j] - ub : ub - P.ubu[j];
                            const HardEval ev = hard_eval(t_, l_, sgn * wu, d_, cr);
                            thn += ev.Th;
                            cn += sgn * ev.coef;
                            if (pass == 0) {
                                res_rows = fmax(res_rows, fabs(ev.r_p));
                                const double c = l_ * t_;
                                cmax = fmax(cmax, c); csum += c;
                            }
                        }
                    }
                    if (pass == 0) { corr[j * KP + kk] = thn; corr[(NU + 6 + j) * KP + kk] = cn; }
                    else corr[(NU + 6 + j) * KP + kk] = cn - corr[(NU + 6 + j) * KP + kk];
                }
                double Dp[6] = {0, 0, 0, 0, 0, 0}, gp[3] = {0, 0, 0};
                if (kk >= k0) {
#pragma unroll 1
                    for (int r = 0; r < NH; ++r) {
                        const double c0 = c_pos(kk, r, 0), c1 = c_pos(kk, r, 1), c2 = c_pos(kk, r, 2);
                        const double cw = c0 * Wsm[(HC0 + 0) * KP + kk] + c1 * Wsm[(HC0 + 1) * KP + kk] + c2 * Wsm[(HC0 + 2) * KP + kk];
                        const double z = zs_(kk, r) * scale(kk), Z = Zs_(kk, r) * scale(kk);
                        const double hv = P.hval[kk * NH + r];
                        double thn = 0.0, cn = 0.0;
#pragma unroll
                        for (int side = 0; side < 2; ++side) {
                            const int a = side * NH + r;
                            const double sgn = side == 0 ? 1.0 : -1.0;
                            const double t_ = QM_TB_LD(sb + (4 * NU + a) * KP), l_ = QM_TB_LD(sb + (4 * NU + 2 * NH + a) * KP);
                            const double s_ = QM_TB_LD(sb + (4 * NU + 4 * NH + a) * KP), m_ = QM_TB_LD(sb + (4 * NU + 6 * NH + a) * KP);
                            const double ct = pass == 0 ? 0.0 : tau - prod[(2 * NU + a) * KP + kk];
                            const double cs = pass == 0 ? 0.0 : tau - prod[(2 * NU + 2 * NH + a) * KP + kk];
                            const double d_ = side == 0 ? lh_(kk, r) - hv : hv - uh_(kk, r);
                            const SoftEval ev = soft_eval(t_, l_, s_, m_, sgn * cw, d_, z, Z, ct, cs);
                            thn += ev.Tht;
                            cn += sgn * ev.coef;
                            if (pass == 0) {
                                res_rows = fmax(res_rows, fmax(fabs(ev.r_p), fabs(ev.r_s)));
                                const double ca = l_ * t_, cb = m_ * s_;
                                cmax = fmax(cmax, fmax(ca, cb)); csum += ca + cb;
                            }
                        }
                        Dp[0] += thn * c0 * c0; Dp[1] += thn * c1 * c0; Dp[2] += thn * c1 * c1;
                        Dp[3] += thn * c2 * c0; Dp[4] += thn * c2 * c1; Dp[5] += thn * c2 * c2;
                        gp[0] += cn * c0; gp[1] += cn * c1; gp[2] += cn * c2;
                    }
                }
                if (pass == 0) {
#pragma unroll
                    for (int c = 0; c < 6; ++c) corr[(NU + c) * KP + kk] = Dp[c];
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) corr[(2 * NU + 6 + c) * KP + kk] = pass == 0 ? gp[c] : gp[c] - corr[(2 * NU + 6 + c) * KP + kk];
            }
#else
            QM_NODE_LOOP(kk) {
                QM_LOAD_STATE(kk);
                double du[NU], gu[NU], Dp[6] = {0, 0, 0, 0, 0, 0}, gp[3] = {0, 0, 0};
#pragma unroll
                for (int j = 0; j < NU; ++j) {
                    double thn = 0.0, cn = 0.0;
                    if (kk < N) {
                        const double wu = Wsm[j * KP + kk];
#pragma unroll
                        for (int side = 0; side < 2; ++side) {
                            const int a = side * NU + j;
                            const double sgn = side == 0 ? 1.0 : -1.0;
                            const double cr = pass == 0 ? 0.0 : tau - prod[a * KP + kk];
                            const HardEval ev = hard_eval(tb[a], lb[a], sgn * wu, hard_d(kk, a), cr);
                            thn += ev.Th;
                            cn += sgn * ev.coef;
                            if (pass == 0) {
                                res_rows = fmax(res_rows, fabs(ev.r_p));
                                const double c = lb[a] * tb[a];
                                cmax = fmax(cmax, c); csum += c;
                            }
                        }
                    }
                    du[j] = thn; gu[j] = cn;
                }
                if (kk >= k0) {
#pragma unroll
                    for (int r = 0; r < NH; ++r) {
                        const double c0 = c_pos(kk, r, 0), c1 = c_pos(kk, r, 1), c2 = c_pos(kk, r, 2);
                        const double cw = c0 * Wsm[(HC0 + 0) * KP + kk] + c1 * Wsm[(HC0 + 1) * KP + kk] + c2 * Wsm[(HC0 + 2) * KP + kk];
                        const double z = zs_(kk, r) * scale(kk), Z = Zs_(kk, r) * scale(kk);
                        double thn = 0.0, cn = 0.0;
#pragma unroll
                        for (int side = 0; side < 2; ++side) {
                            const int a = side * NH + r;
                            const double sgn = side == 0 ? 1.0 : -1.0;
                            const double ct = pass == 0 ? 0.0 : tau - prod[(2 * NU + a) * KP + kk];
                            const double cs = pass == 0 ? 0.0 : tau - prod[(2 * NU + 2 * NH + a) * KP + kk];
                            const SoftEval ev = soft_eval(ts[a], ls[a], ss[a], ms[a], sgn * cw, soft_d(kk, a), z, Z, ct, cs);
                            thn += ev.Tht;
                            cn += sgn * ev.coef;
                            if (pass == 0) {
                                res_rows = fmax(res_rows, fmax(fabs(ev.r_p), fabs(ev.r_s)));
                                const double ca = ls[a] * ts[a], cb = ms[a] * ss[a];
                                cmax = fmax(cmax, fmax(ca, cb)); csum += ca + cb;
                            }
                        }
                        Dp[0] += thn * c0 * c0; Dp[1] += thn * c1 * c0; Dp[2] += thn * c1 * c1;
                        Dp[3] += thn * c2 * c0; Dp[4] += thn * c2 * c1; Dp[5] += thn * c2 * c2;
                        gp[0] += cn * c0; gp[1] += cn * c1; gp[2] += cn * c2;
                    }
                }
                // the corrector solves for the CHANGE of the step, so its right-hand side is the change of the elimination
                // terms against the predictor's (same lane, same slots)
#pragma unroll
                for (int j = 0; j < NU; ++j) {
                    if (pass == 0) { corr[j * KP + kk] = du[j]; corr[(NU + 6 + j) * KP + kk] = gu[j]; }
                    else corr[(NU + 6 + j) * KP + kk] = gu[j] - corr[(NU + 6 + j) * KP + kk];
                }
                if (pass == 0) {
#pragma unroll
                    for (int c = 0; c < 6; ++c) corr[(NU + c) * KP + kk] = Dp[c];
                }
#pragma unroll
                for (int c = 0; c < 3; ++c) corr[(2 * NU + 6 + c) * KP + kk] = pass == 0 ? gp[c] : gp[c] - corr[(2 * NU + 6 + c) * KP + kk];
            }
#endif
            if (pass == 0) {
                // ---- convergence test on the state BEFORE this iteration's step ----
                res_rows = L.max(res_rows);
                cmax = L.max(cmax);
                mubar = L.sum(csum) / ncomp;
                res_track = fmax(res_rows, phi * res_v0);
#if defined(SDFNMPC_QP_DEBUG_INST)
                if (P.debug && lane == 0) printf("it %d res %.3e (rows %.3e) cmax %.3e phi %.3e mubar %.3e\n", it, res_track, res_rows, cmax, phi, mubar);
#endif
                if (!(res_track == res_track) || !(mubar == mubar) || res_track > 1e300) {
                    if (lane == 0 && P.res_out) { P.res_out[0] = res_v0; P.res_out[1] = res_rows + 1e3 * mubar; }   // breakdown diagnostics
                    status = 4; stop = true; break;
                }
                if (res_track <= P.tol_res && cmax <= P.tol) { status = 0; stop = true; break; }
                if (it >= P.max_iter) { status = 2; stop = true; break; }
                // Mehrotra's corrector can cycle on poorly centred iterates; if the largest complementarity product
                // has not dropped by 30 % over two iterations switch (for good) to plain damped path following.
                if (it >= 2 && cmax > 0.7 * cm_prev2) safe = true;
                cm_prev2 = cm_prev1; cm_prev1 = cmax;
            }
            L.sync();
            QPT(pass * 8 + 0, it == 2)
            if (pass == 0) {
                // =========== factorising backward sweep ===========
                struct FacOps { Acc2 h[2][2]; TOps t; };
                auto ld_ops = [&](int i_) {
                    FacOps o;
                    ld_h(N - 1 - i_, o.h);
                    o.t = ld_t(N - 1 - i_);
                    return o;
                };
                Acc2 Pf[2][2];
                ld_h(N, Pf);
                inject(Pf, N, phi);
                bool ok = true;
                auto stage = [&](int i_, const FacOps& o) {
                    L.sync();
                    QM_CONVERGED()
                    const int kk = N - 1 - i_;
                    Acc2 M[2][2];
                    double f0[2], f1[2], f2[2];
#pragma unroll
                    for (int r = 0; r < 2; ++r) { f0[r] = o.t.f0[r]; f1[r] = o.t.f1[r]; f2[r] = o.t.f2[r]; M[r][0] = o.h[r][0]; M[r][1] = o.h[r][1]; }
                    inject(M, kk, phi);
                    // the affine column of T (position 12: tile 1, g = 4) carries phi rd; its own entry (slot 10) stays 1
                    if (g_ == 4) { f0[1] *= phi; f1[1] *= phi; if (q_ != 2) f2[1] *= phi; }
                    Acc2 X[2][2];
                    // tile row 1 first: it holds the input block, whose (scalar) factorisation then runs under the tensor
                    // work of tile row 0
#pragma unroll
                    for (int ri = 1; ri >= 0; --ri)
#pragma unroll
                        for (int rm = 0; rm < 2; ++rm) {
                            X[ri][rm].x = 0.0; X[ri][rm].y = 0.0;
                            dmma(X[ri][rm], f0[ri], Pf[rm][0].x);
                            dmma(X[ri][rm], f1[ri], Pf[rm][0].y);
                            dmma(X[ri][rm], f2[ri], Pf[rm][1].x);
                        }
#pragma unroll
                    for (int ri = 1; ri >= 0; --ri)
#pragma unroll
                        for (int rj = 1; rj >= 0; --rj) {
                            dmma(M[ri][rj], X[ri][0].x, f0[rj]);
                            dmma(M[ri][rj], X[ri][0].y, f1[rj]);
                            dmma(M[ri][rj], X[ri][1].x, f2[rj]);
                        }
                    // Muu (positions 9, 11, 13, 15): entry (a, b) sits in lane 4 (1 + 2 a) + b, register y of tile (1, 1)
                    const double mu_ = M[1][1].y;
                    const double m00 = __shfl_sync(FULL, mu_, 4), m10 = __shfl_sync(FULL, mu_, 12), m11 = __shfl_sync(FULL, mu_, 13),
                                 m20 = __shfl_sync(FULL, mu_, 20), m21 = __shfl_sync(FULL, mu_, 21), m22 = __shfl_sync(FULL, mu_, 22),
                                 m30 = __shfl_sync(FULL, mu_, 28), m31 = __shfl_sync(FULL, mu_, 29), m32 = __shfl_sync(FULL, mu_, 30),
                                 m33 = __shfl_sync(FULL, mu_, 31);
                    // LDL' of Muu, redundantly in every lane (a closed-form inverse by 2 x 2 blocks needs only two dependent
                    // reciprocals, but is not backward stable: late iterations carry ~1e10 rank-one barrier terms that cancel in
                    // the Schur complement below, and the recursion lost definiteness on 1.4 % of the instances)
                    const double r0 = rcp(m00);
                    const double l10 = m10 * r0, l20 = m20 * r0, l30 = m30 * r0;
                    const double d1 = m11 - l10 * m10;
                    const double r1 = rcp(d1);
                    const double t21 = m21 - l20 * m10, t31 = m31 - l30 * m10;
                    const double l21 = t21 * r1, l31 = t31 * r1;
                    const double d2 = m22 - l20 * m20 - l21 * t21;
                    const double r2 = rcp(d2);
                    const double t32 = m32 - l30 * m20 - l31 * t21;
                    const double l32 = t32 * r2;
                    const double d3 = m33 - l30 * m30 - l31 * t31 - l32 * t32;
                    const double r3 = rcp(d3);
                    ok = ok && (m00 > 0.0) && (d1 > 0.0) && (d2 > 0.0) && (d3 > 0.0);
                    double kq[2];
                    const double mxu[2] = {M[0][1].y, M[1][1].y};        // M~[8 ri + g][u_q]
                    {
                        // the two rows of the lane's quads are solved side by side (two independent dependency chains); the lane
                        // keeps component q of each, picked with selects on two predicates (no branches)
                        const int qb = lane & ~3;
                        double v0[2], v1[2], v2[2], v3[2], z0[2], z1[2], z2[2], z3[2];
#pragma unroll
                        for (int ri = 0; ri < 2; ++ri) {
                            v0[ri] = __shfl_sync(FULL, mxu[ri], qb); v1[ri] = __shfl_sync(FULL, mxu[ri], qb + 1);
                            v2[ri] = __shfl_sync(FULL, mxu[ri], qb + 2); v3[ri] = __shfl_sync(FULL, mxu[ri], qb + 3);
                        }
#pragma unroll
                        for (int ri = 0; ri < 2; ++ri) {
                            const double y1 = v1[ri] - l10 * v0[ri], y2 = v2[ri] - l20 * v0[ri] - l21 * y1,
                                         y3 = v3[ri] - l30 * v0[ri] - l31 * y1 - l32 * y2;
                            z3[ri] = y3 * r3; z2[ri] = y2 * r2 - l32 * z3[ri]; z1[ri] = y1 * r1 - l21 * z2[ri] - l31 * z3[ri];
                            z0[ri] = v0[ri] * r0 - l10 * z1[ri] - l20 * z2[ri] - l30 * z3[ri];
                        }
                        const bool qlo = (q_ & 1) == 0, qhi = q_ < 2;
#pragma unroll
                        for (int ri = 0; ri < 2; ++ri) {
                            const double a_ = qlo ? z0[ri] : z1[ri], b_ = qlo ? z2[ri] : z3[ri];
                            kq[ri] = -(qhi ? a_ : b_);
                        }
                    }
                    // P~ <- M~ + M.u K~
#pragma unroll
                    for (int ri = 0; ri < 2; ++ri)
#pragma unroll
                        for (int rj = 0; rj < 2; ++rj) {
                            Pf[ri][rj] = M[ri][rj];
                            dmma(Pf[ri][rj], mxu[ri], kq[rj]);
                        }
                    // for the roll-outs and the corrector: gains (with the feed-forward in the affine position) and the factor
                    QM_KST(Kf + kk * S::KF + lane, kq[0]);
                    QM_KST(Kf + kk * S::KF + 32 + lane, kq[1]);
                    if (lane == 0) {
                        double2* lo = reinterpret_cast<double2*>(Lf + (size_t)kk * S::LF);
                        __stcg(lo + 0, make_double2(r0, r1)); __stcg(lo + 1, make_double2(r2, r3));
                        __stcg(lo + 2, make_double2(l10, l20)); __stcg(lo + 3, make_double2(l30, l21)); __stcg(lo + 4, make_double2(l31, l32));
                    }
                };
                pipelined_stages<QM_D_FAC, FacOps>(N, ld_ops, stage);
                if (!ok) status = 4;
            } else {
                // =========== corrector: same matrix, right-hand side = change of the elimination terms, no dynamics residual ===========
                //   r~_k = delta r_k + T_k' p_{k+1},   p_k = r~_k + K_k' r~_k,u     (column 0 of the accumulators, lanes 4 g)
                // 8 DMMA per stage; the input residuals r~_k,u are left in `ffd` and turned into feed-forwards for all stages
                // at once afterwards (lane = stage), which keeps the 4 x 4 solves out of the recursion.
                const int cr0 = (q_ == 0 && g_ < 3) ? 2 * NU + 6 + g_ : NCORR;
                const int cr1 = (q_ == 0 && (g_ & 1)) ? NU + 6 + (g_ >> 1) : NCORR;
                struct BackOps { TOps t; double kf[2]; };
                auto ld_ops = [&](int i_) {
                    BackOps o;
                    const int kk = N - 1 - i_;
                    o.t = ld_t(kk);
                    o.kf[0] = QM_KLD(Kf + kk * S::KF + lane);
                    o.kf[1] = QM_KLD(Kf + kk * S::KF + 32 + lane);
                    return o;
                };
                double bA = 2 * q_ < 3 ? corr[(2 * NU + 6 + 2 * q_) * KP + N] : 0.0;
                double bB = q_ == 0 ? corr[(2 * NU + 6 + 1) * KP + N] : 0.0;
                double bC = 0.0;
                auto stage = [&](int i_, const BackOps& o) {
                    L.sync();
                    QM_CONVERGED()
                    const int kk = N - 1 - i_;
                    Acc2 R[2];
                    R[0].x = corr[cr0 * KP + kk]; R[0].y = 0.0;
                    R[1].x = corr[cr1 * KP + kk]; R[1].y = 0.0;
                    dmma(R[0], o.t.f0[0], bA); dmma(R[1], o.t.f0[1], bA);
                    dmma(R[0], o.t.f1[0], bB); dmma(R[1], o.t.f1[1], bB);
                    dmma(R[0], o.t.f2[0], bC); dmma(R[1], o.t.f2[1], bC);
                    if (q_ == 0 && (g_ & 1)) ffd[(g_ >> 1) * KP + kk] = R[1].x;
                    const double bR = __shfl_sync(FULL, R[1].x, 4 * (1 + 2 * q_));
                    dmma(R[0], o.kf[0], bR); dmma(R[1], o.kf[1], bR);
                    bA = __shfl_sync(FULL, R[0].x, 8 * q_);
                    bB = __shfl_sync(FULL, R[0].x, 8 * q_ + 4);
                    bC = __shfl_sync(FULL, R[1].x, 8 * q_);
                    if (q_ >= 2) bC = 0.0;
                };
                pipelined_stages<QM_D_BACK, BackOps>(N, ld_ops, stage);
                L.sync();
                // feed-forward of every stage: -Muu^-1 r~_u with the factor the predictor's sweep left (lane = stage)
                QM_LANES(kk, N) {
                    double Fk[10];
#pragma unroll
                    for (int e = 0; e < 10; e += 2) {
                        const double2 v = __ldcg(reinterpret_cast<const double2*>(Lf + (size_t)kk * S::LF + e));
                        Fk[e] = v.x; Fk[e + 1] = v.y;
                    }
                    double x[NU];
#pragma unroll
                    for (int i = 0; i < NU; ++i) {
                        double s_ = ffd[i * KP + kk];
#pragma unroll
                        for (int l = 0; l < i; ++l) s_ -= Fk[NU + l * NU - l * (l + 1) / 2 + (i - l - 1)] * x[l];
                        x[i] = s_;
                    }
#pragma unroll
                    for (int i = NU - 1; i >= 0; --i) {
                        double s_ = x[i] * Fk[i];
#pragma unroll
                        for (int l = i + 1; l < NU; ++l) s_ -= Fk[NU + i * NU - i * (i + 1) / 2 + (l - i - 1)] * x[l];
                        x[i] = s_;
                    }
#pragma unroll
                    for (int i = 0; i < NU; ++i) ffd[i * KP + kk] = -x[i];
                }
            }
            L.sync();
            QPT(pass * 8 + 1, it == 2)
            // =========== forward roll-out ===========
            // x~_k = [dx_k; 1] sits in column 0 of the B operand (lanes 0..3); u_k = K~_k x~_k and x~_{k+1} = T_k [u_k; x~_k]
            // come out in column 0 of the accumulators (lanes 4 g); four shuffles put them back into operand position.
            // 11 DMMA per stage.  Predictor: the feed-forward is the affine column of K~.  Corrector: the CHANGE of the step
            // (affine entry 0, feed-forward from `ffd`), accumulated onto the predictor's step.
            {
                struct RollOps { double tA[2], tB[2], tC[2], tU[2], kA, kB, kC; };
                auto ld_ops = [&](int kk) {
                    RollOps o;
                    const double* a_ = ABt + (size_t)kk * S::ABT + g_;
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        const bool in = mt == 0 || g_ < 4;        // output slots 12..15 do not exist
                        o.tA[mt] = in ? __ldcg(a_ + (2 * q_) * 12 + 8 * mt) : 0.0;
                        o.tB[mt] = in ? __ldcg(a_ + (2 * q_ + 1) * 12 + 8 * mt) : 0.0;
                        o.tC[mt] = in ? __ldcg(a_ + (8 + 2 * q_) * 12 + 8 * mt) : 0.0;
                        o.tU[mt] = in ? __ldcg(a_ + (9 + 2 * q_) * 12 + 8 * mt) : 0.0;
                    }
                    // K~[u = g][slot]: slot s sits at position s (s < 8) or 8 + 2 (s - 8), fragment entry [tile][4 (pos % 8) + u]
                    const double* k_ = Kf + kk * S::KF + g_;
                    const bool ku = g_ < NU;
                    o.kA = ku ? QM_KLD(k_ + 4 * (2 * q_)) : 0.0;
                    o.kB = ku ? QM_KLD(k_ + 4 * (2 * q_ + 1)) : 0.0;
                    o.kC = ku ? QM_KLD(k_ + 32 + 4 * (2 * q_)) : 0.0;
                    return o;
                };
                const bool delta = pass == 1;
                double bA = 0.0, bB = 0.0, bC = (q_ == 2 && !delta) ? 1.0 : 0.0;       // x~_0 = [0; 1] (dx_0 is part of w already)
                if (lane < NX) dW[(NU + lane) * KP] = 0.0;
                L.sync();
                auto stage = [&](int kk, const RollOps& o) {
                    L.sync();
                    QM_CONVERGED()
                    // affine column of T (position 12: step C, q = 2) carries phi rd; its own entry (slot 10) stays 1
                    const double tC0 = q_ == 2 ? o.tC[0] * phi : o.tC[0];
                    const double tC1 = (q_ == 2 && g_ != 2) ? o.tC[1] * phi : o.tC[1];
                    Acc2 U, Xn[2];
                    U.x = (delta && q_ == 0 && g_ < NU) ? ffd[g_ * KP + kk] : 0.0;
                    U.y = 0.0;
                    Xn[0].x = 0.0; Xn[0].y = 0.0; Xn[1].x = 0.0; Xn[1].y = 0.0;
                    dmma(U, o.kA, bA); dmma(Xn[0], o.tA[0], bA); dmma(Xn[1], o.tA[1], bA);
                    dmma(U, o.kB, bB); dmma(Xn[0], o.tB[0], bB); dmma(Xn[1], o.tB[1], bB);
                    dmma(U, o.kC, bC); dmma(Xn[0], tC0, bC); dmma(Xn[1], tC1, bC);
                    const double bU = __shfl_sync(FULL, U.x, 4 * q_);
                    dmma(Xn[0], o.tU[0], bU); dmma(Xn[1], o.tU[1], bU);
                    if (q_ == 0) {
                        if (delta) {
                            if (g_ < NU) dW[g_ * KP + kk] += U.x;
                            dW[(NU + g_) * KP + kk + 1] += Xn[0].x;
                            if (g_ < 2) dW[(NU + 8 + g_) * KP + kk + 1] += Xn[1].x;
                        } else {
                            if (g_ < NU) dW[g_ * KP + kk] = U.x;
                            dW[(NU + g_) * KP + kk + 1] = Xn[0].x;
                            if (g_ < 2) dW[(NU + 8 + g_) * KP + kk + 1] = Xn[1].x;
                        }
                    }
                    bA = __shfl_sync(FULL, Xn[0].x, 8 * q_);
                    bB = __shfl_sync(FULL, Xn[0].x, 8 * q_ + 4);
                    bC = __shfl_sync(FULL, Xn[1].x, 4 * q_);
                };
                pipelined_stages<QM_D_ROLL, RollOps>(N, ld_ops, stage);
            }
            L.sync();
            QPT(pass * 8 + 2, it == 2)
            // =========== row steps (lane = node) ===========
            double rmax = 0.0;              // largest -dx/x over the lane's rows
#if QM_ROLLED_ROWS
            QM_NODE_LOOP(kk) {
                const double* sb = TB + kk;
                double* db = TD + kk;
#pragma unroll 1
                for (int j = 0; j < NU; ++j) {
                    double dt_[2] = {0.0, 0.0}, dl_[2] = {0.0, 0.0};
                    if (kk < N) {
                        const double wu = Wsm[j * KP + kk], dwu = dW[j * KP + kk];
                        const double ub = P.ubar[kk * NU + j];
#pragma unroll
                        for (int side = 0; side < 2; ++side) {
                            const int a = side * NU + j;
                            const double sgn = side == 0 ? 1.0 : -1.0;
                            const double t_ = QM_TB_LD(sb + a * KP), l_ = QM_TB_LD(sb + (2 * NU + a) * KP);
                            const double cr = pass == 0 ? 0.0 : tau - prod[a * KP + kk];
                            const double d_ = side == 0 ? P.lbu[j] - ub : ub - P.ubu[j];
                            const HardEval ev = hard_eval(t_, l_, sgn * wu, d_, cr);
                            const double kap = cr - l_ * t_;
                            dt_[side] = sgn * dwu - ev.r_p;
                            dl_[side] = (kap - l_ * dt_[side]) * ev.rt;
                            rmax = fmax(rmax, fmax(boundary_ratio(dt_[side], ev.rt), boundary_ratio(dl_[side], ev.rl)));
                        }
                    }
#pragma unroll
                    for (int side = 0; side < 2; ++side) {
                        QM_ST(db + (side * NU + j) * KP, dt_[side]);
                        QM_ST(db + (2 * NU + side * NU + j) * KP, dl_[side]);
                    }
                }
#pragma unroll 1
                for (int r = 0; r < NH; ++r) {
                    double dt_[2] = {0.0, 0.0}, dl_[2] = {0.0, 0.0}, ds_[2] = {0.0, 0.0}, dm_[2] = {0.0, 0.0};
                    if (kk >= k0) {
                        const double c0 = c_pos(kk, r, 0), c1 = c_pos(kk, r, 1), c2 = c_pos(kk, r, 2);
                        const double cw = c0 * Wsm[(HC0 + 0) * KP + kk] + c1 * Wsm[(HC0 + 1) * KP + kk] + c2 * Wsm[(HC0 + 2) * KP + kk];
                        const double cd = c0 * dW[(HC0 + 0) * KP + kk] + c1 * dW[(HC0 + 1) * KP + kk] + c2 * dW[(HC0 + 2) * KP + kk];
                        const double z = zs_(kk, r) * scale(kk), Z = Zs_(kk, r) * scale(kk);
                        const double hv = P.hval[kk * NH + r];
#pragma unroll
                        for (int side = 0; side < 2; ++side) {
                            const int a = side * NH + r;
                            const double sgn = side == 0 ? 1.0 : -1.0;
                            const double t_ = QM_TB_LD(sb + (4 * NU + a) * KP), l_ = QM_TB_LD(sb + (4 * NU + 2 * NH + a) * KP);
                            const double s_ = QM_TB_LD(sb + (4 * NU + 4 * NH + a) * KP), m_ = QM_TB_LD(sb + (4 * NU + 6 * NH + a) * KP);
                            const double ct = pass == 0 ? 0.0 : tau - prod[(2 * NU + a) * KP + kk];
                            const double cs = pass == 0 ? 0.0 : tau - prod[(2 * NU + 2 * NH + a) * KP + kk];
                            const double d_ = side == 0 ? lh_(kk, r) - hv : hv - uh_(kk, r);
                            const SoftEval ev = soft_eval(t_, l_, s_, m_, sgn * cw, d_, z, Z, ct, cs);
                            const double kap = ct - l_ * t_, nu = cs - m_ * s_;
                            const double gd = sgn * cd;
                            ds_[side] = (ev.rho - ev.Th * gd) * ev.rD;
                            dt_[side] = gd + ds_[side] - ev.r_p;
                            dl_[side] = (kap - l_ * dt_[side]) * ev.rt;
                            dm_[side] = (nu - m_ * ds_[side]) * ev.rs;
                            rmax = fmax(rmax, fmax(fmax(boundary_ratio(dt_[side], ev.rt), boundary_ratio(dl_[side], ev.rl)),
                                                   fmax(boundary_ratio(ds_[side], ev.rs), boundary_ratio(dm_[side], ev.rm))));
                        }
                    }
#pragma unroll
                    for (int side = 0; side < 2; ++side) {
                        const int a = side * NH + r;
                        QM_ST(db + (4 * NU + a) * KP, dt_[side]);
                        QM_ST(db + (4 * NU + 2 * NH + a) * KP, dl_[side]);
                        QM_ST(db + (4 * NU + 4 * NH + a) * KP, ds_[side]);
                        QM_ST(db + (4 * NU + 6 * NH + a) * KP, dm_[side]);
                    }
                }
            }
#else
            QM_NODE_LOOP(kk) {
                QM_LOAD_STATE(kk);
#pragma unroll
                for (int a = 0; a < 2 * NU; ++a) { dtb[a] = 0.0; dlb[a] = 0.0; }
#pragma unroll
                for (int a = 0; a < 2 * NH; ++a) { dts[a] = 0.0; dls[a] = 0.0; dss[a] = 0.0; dms[a] = 0.0; }
                if (kk < N) {
#pragma unroll
                    for (int j = 0; j < NU; ++j) {
                        const double wu = Wsm[j * KP + kk], dwu = dW[j * KP + kk];
#pragma unroll
                        for (int side = 0; side < 2; ++side) {
                            const int a = side * NU + j;
                            const double sgn = side == 0 ? 1.0 : -1.0;
                            const double cr = pass == 0 ? 0.0 : tau - prod[a * KP + kk];
                            const HardEval ev = hard_eval(tb[a], lb[a], sgn * wu, hard_d(kk, a), cr);
                            const double kap = cr - lb[a] * tb[a];
                            dtb[a] = sgn * dwu - ev.r_p;
                            dlb[a] = (kap - lb[a] * dtb[a]) * ev.rt;
                            rmax = fmax(rmax, fmax(boundary_ratio(dtb[a], ev.rt), boundary_ratio(dlb[a], ev.rl)));
                        }
                    }
                }
                if (kk >= k0) {
#pragma unroll
                    for (int r = 0; r < NH; ++r) {
                        const double c0 = c_pos(kk, r, 0), c1 = c_pos(kk, r, 1), c2 = c_pos(kk, r, 2);
                        const double cw = c0 * Wsm[(HC0 + 0) * KP + kk] + c1 * Wsm[(HC0 + 1) * KP + kk] + c2 * Wsm[(HC0 + 2) * KP + kk];
                        const double cd = c0 * dW[(HC0 + 0) * KP + kk] + c1 * dW[(HC0 + 1) * KP + kk] + c2 * dW[(HC0 + 2) * KP + kk];
                        const double z = zs_(kk, r) * scale(kk), Z = Zs_(kk, r) * scale(kk);
#pragma unroll
                        for (int side = 0; side < 2; ++side) {
                            const int a = side * NH + r;
                            const double sgn = side == 0 ? 1.0 : -1.0;
                            const double ct = pass == 0 ? 0.0 : tau - prod[(2 * NU + a) * KP + kk];
                            const double cs = pass == 0 ? 0.0 : tau - prod[(2 * NU + 2 * NH + a) * KP + kk];
                            const SoftEval ev = soft_eval(ts[a], ls[a], ss[a], ms[a], sgn * cw, soft_d(kk, a), z, Z, ct, cs);
                            const double kap = ct - ls[a] * ts[a], nu = cs - ms[a] * ss[a];
                            const double gd = sgn * cd;
                            dss[a] = (ev.rho - ev.Th * gd) * ev.rD;
                            dts[a] = gd + dss[a] - ev.r_p;
                            dls[a] = (kap - ls[a] * dts[a]) * ev.rt;
                            dms[a] = (nu - ms[a] * dss[a]) * ev.rs;
                            rmax = fmax(rmax, fmax(fmax(boundary_ratio(dts[a], ev.rt), boundary_ratio(dls[a], ev.rl)),
                                                   fmax(boundary_ratio(dss[a], ev.rs), boundary_ratio(dms[a], ev.rm))));
                        }
                    }
                }
                QM_STORE_DELTA(kk);
            }
#endif
            rmax = L.max(rmax);
            const double amax = rmax > 1e-30 ? 1.0 / rmax : 1e30;
            if (pass == 0) {
                // affine step length -> centring parameter; second-order products for the corrector
                const double aa = fmin(1.0, amax);
                const double so = safe ? 0.0 : 1.0;
                double csum_aff = 0.0;
#if QM_ROLLED_ROWS
                QM_NODE_LOOP(kk) {
                    const double* sb = TB + kk;
                    const double* db = TD + kk;
                    if (kk < N) {
#pragma unroll
                        for (int a = 0; a < 2 * NU; ++a) {
                            const double t_ = QM_TB_LD(sb + a * KP), l_ = QM_TB_LD(sb + (2 * NU + a) * KP);
                            const double dt_ = QM_LD(db + a * KP), dl_ = QM_LD(db + (2 * NU + a) * KP);
                            csum_aff += (l_ + aa * dl_) * (t_ + aa * dt_);
                            prod[a * KP + kk] = so * dl_ * dt_;
                        }
                    }
                    if (kk >= k0) {
#pragma unroll
                        for (int a = 0; a < 2 * NH; ++a) {
                            const double t_ = QM_TB_LD(sb + (4 * NU + a) * KP), l_ = QM_TB_LD(sb + (4 * NU + 2 * NH + a) * KP);
                            const double s_ = QM_TB_LD(sb + (4 * NU + 4 * NH + a) * KP), m_ = QM_TB_LD(sb + (4 * NU + 6 * NH + a) * KP);
                            const double dt_ = QM_LD(db + (4 * NU + a) * KP), dl_ = QM_LD(db + (4 * NU + 2 * NH + a) * KP);
                            const double ds_ = QM_LD(db + (4 * NU + 4 * NH + a) * KP), dm_ = QM_LD(db + (4 * NU + 6 * NH + a) * KP);
                            csum_aff += (l_ + aa * dl_) * (t_ + aa * dt_) + (m_ + aa * dm_) * (s_ + aa * ds_);
                            prod[(2 * NU + a) * KP + kk] = so * dl_ * dt_;
                            prod[(2 * NU + 2 * NH + a) * KP + kk] = so * dm_ * ds_;
                        }
                    }
                }
#else
                QM_NODE_LOOP(kk) {
                    QM_LOAD_STATE(kk);
                    QM_LOAD_DELTA(kk);
                    if (kk < N) {
#pragma unroll
                        for (int a = 0; a < 2 * NU; ++a) {
                            csum_aff += (lb[a] + aa * dlb[a]) * (tb[a] + aa * dtb[a]);
                            prod[a * KP + kk] = so * dlb[a] * dtb[a];
                        }
                    }
                    if (kk >= k0) {
#pragma unroll
                        for (int a = 0; a < 2 * NH; ++a) {
                            csum_aff += (ls[a] + aa * dls[a]) * (ts[a] + aa * dts[a]) + (ms[a] + aa * dms[a]) * (ss[a] + aa * dss[a]);
                            prod[(2 * NU + a) * KP + kk] = so * dls[a] * dts[a];
                            prod[(2 * NU + 2 * NH + a) * KP + kk] = so * dms[a] * dss[a];
                        }
                    }
                }
#endif
                const double mu_aff = L.sum(csum_aff) / ncomp;
                double sig = mu_aff / fmax(mubar, 1e-300);
                sig = fmin(fmax(sig * sig * sig, safe ? 0.2 : 0.0), 1.0);
                tau = fmax(sig * mubar, 0.1 * P.tol);
                L.sync();
                QPT(3, it == 2)
            } else {
                // =========== damped step ===========
                const double al = fmin(1.0, 0.995 * amax);
#if QM_ROLLED_ROWS
                QM_NODE_LOOP(kk) {
#pragma unroll
                    for (int e = 0; e < S::NROW; ++e) {
                        double* q_ = TB + e * KP + kk;
                        QM_TB_ST(q_, QM_TB_LD(q_) + al * QM_LD(TD + e * KP + kk));
                    }
                }
#else
                QM_NODE_LOOP(kk) {
                    QM_LOAD_STATE(kk);
                    QM_LOAD_DELTA(kk);
#pragma unroll
                    for (int a = 0; a < 2 * NU; ++a) { tb[a] += al * dtb[a]; lb[a] += al * dlb[a]; }
#pragma unroll
                    for (int a = 0; a < 2 * NH; ++a) { ts[a] += al * dts[a]; ls[a] += al * dls[a]; ss[a] += al * dss[a]; ms[a] += al * dms[a]; }
                    QM_STORE_STATE(kk);
                }
#endif
                QM_LANES(e, NW * KP) Wsm[e] += al * dW[e];
                phi *= (1.0 - al);
                L.sync();
                QPT(11, it == 2)
            }
        }
        if (stop || status == 4) break;
    }

    // =====================================================================================================
    // verification of the KKT residuals the iteration only tracked (costates by one backward sweep); once per solve, plain
    // loads of the operands by position
    // =====================================================================================================
    double res_true = 0.0;
    {
        QM_NODE_LOOP(kk) {
            QM_LOAD_STATE(kk);
            for (int j = 0; j < NU; ++j) corr[j * KP + kk] = kk < N ? lb[j] - lb[NU + j] : 0.0;
            for (int r = 0; r < NH; ++r) corr[(NU + r) * KP + kk] = kk >= k0 ? ls[r] - ls[NH + r] : 0.0;
        }
        L.sync();
        QPT(21, true)
        // pass 1, every node on its own (no recursion: the loads of consecutive nodes overlap): q_k = g_k + H_k w_k - (multiplier
        // terms) into dW (free after the last iteration), and the dynamics residuals
        if (status != 4) {
            // nodes are taken VC at a time with all their loads issued before the first use: one memory round trip per chunk
            // (node by node the loop paid a loaded L2 round trip, ~3.5 k cycles, 21 times: 75 k cycles per solve; 7 at a time needs 210 registers and spills)
            constexpr int VC = 3;
            const bool row_lane = lane < NW, dyn_lane = lane >= 16 && lane < 16 + NX;
            const int vi = row_lane ? lane : (dyn_lane ? lane - 16 : 0), vpi = S::wpos(row_lane ? lane : 0);
            for (int k0_ = 0; k0_ <= N; k0_ += VC) {
                double hv[VC][NW], gv[VC];
#pragma unroll
                for (int t = 0; t < VC; ++t) {
                    const int kk = k0_ + t <= N ? k0_ + t : N, kt = kk < N ? kk : (N > 0 ? N - 1 : 0);
                    // row lanes: row of H~_kk and the gradient entry; dynamics lanes: row of [A B]_kk and the defect
                    const double* src = row_lane ? Hp + (size_t)kk * S::HP + vpi * 16 : ABt + (size_t)kt * S::ABT + vi;
                    const int stride = row_lane ? 1 : 12;
#pragma unroll
                    for (int j = 0; j < NW; ++j) hv[t][j] = __ldcg(src + S::wpos(j) * stride);
                    gv[t] = row_lane ? __ldcg(gpos + kk * S::GP + vpi) : P.bdyn[kt * NX + vi];
                }
#pragma unroll
                for (int t = 0; t < VC; ++t) {
                    const int kk = k0_ + t;
                    if (kk > N) break;
                    double s = gv[t];
#pragma unroll
                    for (int j = 0; j < NW; ++j) s += hv[t][j] * Wsm[j * KP + kk];
                    if (row_lane) {
                        if (vi < NU) s -= corr[vi * KP + kk];
                        if (vi >= HC0 && vi < HC0 + 3)
                            for (int r = 0; r < NH; ++r) s -= corr[(NU + r) * KP + kk] * c_pos(kk, r, vi - HC0);
                        dW[vi * KP + kk] = s;
                    } else if (dyn_lane && kk < N) {
                        res_true = fmax(res_true, fabs(s - Wsm[(NU + vi) * KP + kk + 1]));
                    }
                }
            }
            L.sync();
            QPT(22, true)
            // pass 2, the costate recursion: stationarity of node kk = q_kk + [A B]_kk' costate_{kk+1}; its state rows DEFINE the
            // costate of node kk, its input rows are residuals.  The column of [A B] a lane needs is fetched one node ahead.
            double tcol[NX], tnext[NX];
            const int pi = lane < NW ? S::wpos(lane) : 0;
#pragma unroll
            for (int l = 0; l < NX; ++l) tnext[l] = N >= 1 ? __ldcg(ABt + (size_t)(N - 1) * S::ABT + pi * 12 + l) : 0.0;
            for (int kk = N; kk >= 0; --kk) {
                const double* pin = pp + ((N - kk) & 1) * 16;        // costate of node kk + 1
                double* pout = pp + (((N - kk) & 1) ^ 1) * 16;       // costate of node kk
                if (kk < N) {
#pragma unroll
                    for (int l = 0; l < NX; ++l) tcol[l] = tnext[l];
                    if (kk >= 1) {
#pragma unroll
                        for (int l = 0; l < NX; ++l) tnext[l] = __ldcg(ABt + (size_t)(kk - 1) * S::ABT + pi * 12 + l);
                    }
                }
                if (lane < NW && !(kk == N && lane < NU)) {
                    double s = dW[lane * KP + kk];
                    if (kk < N) {
#pragma unroll
                        for (int l = 0; l < NX; ++l) s += tcol[l] * pin[l];
                    }
                    if (lane < NU) res_true = fmax(res_true, fabs(s));      // input stationarity
                    else pout[lane - NU] = s;                               // costate of node kk (x-stationarity == 0)
                }
                L.sync();
            }
        }
        res_true = L.max(res_true);
        if (status == 0 && !(res_true <= 100.0 * P.tol_res)) status = 3;    // tracked residual was wrong: report, do not hide
    }

    double step_inf = 0.0;
    QM_LANES(e, (N + 1) * NW) { const int kk = e / NW, i = e - kk * NW; step_inf = fmax(step_inf, fabs(Wsm[i * KP + kk])); }
    step_inf = L.max(step_inf);
    // ---- full SQP step and outputs ----
    QPT(19, true)
    if (status != 4) {
        // every lane's loads of the iterate go out together (a read-modify-write per trip would pay the memory latency per trip)
        constexpr int XT = NFIX > 0 ? ((NFIX + 1) * NX + 31) / 32 : 1, UT = NFIX > 0 ? (NFIX * NU + 31) / 32 : 1;
        if (NFIX > 0) {
            double xv[XT], uv[UT];
#pragma unroll
            for (int t = 0; t < XT; ++t) { const int e = t * 32 + lane; xv[t] = e < (N + 1) * NX ? P.xbar[e] : 0.0; }
#pragma unroll
            for (int t = 0; t < UT; ++t) { const int e = t * 32 + lane; uv[t] = e < N * NU ? P.ubar[e] : 0.0; }
#pragma unroll
            for (int t = 0; t < XT; ++t) { const int e = t * 32 + lane; if (e < (N + 1) * NX) { const int kk = e / NX, i = e - kk * NX; P.xbar[e] = xv[t] + Wsm[(NU + i) * KP + kk]; } }
#pragma unroll
            for (int t = 0; t < UT; ++t) { const int e = t * 32 + lane; if (e < N * NU) { const int kk = e / NU, j = e - kk * NU; P.ubar[e] = uv[t] + Wsm[j * KP + kk]; } }
        } else {
            QM_LANES(e, (N + 1) * NX) { const int kk = e / NX, i = e - kk * NX; P.xbar[e] += Wsm[(NU + i) * KP + kk]; }
            QM_LANES(e, N * NU) { const int kk = e / NU, j = e - kk * NU; P.ubar[e] += Wsm[j * KP + kk]; }
        }
    }
    L.sync();
    QM_LANES(j, NU) P.u0[j] = P.ubar[j];
    QM_NODE_LOOP(kk) {
        QM_LOAD_STATE(kk);
        if (P.lam_hard && kk < N)
            for (int a = 0; a < 2 * NU; ++a) P.lam_hard[kk * 2 * NU + a] = lb[a];
        if (P.lam_soft)
            for (int a = 0; a < 2 * NH; ++a) P.lam_soft[kk * 2 * NH + a] = kk >= k0 ? ls[a] : 0.0;
    }
    QPT(20, true)
    if (lane == 0) {
        *P.status = status;
        *P.iters = it;
        if (P.res_out && status != 4) { P.res_out[0] = res_true; P.res_out[1] = step_inf; }
    }
}

#undef QM_LD
#undef QM_ST
#undef QM_LOAD
#undef QM_STORE
#undef QM_LOAD_STATE
#undef QM_STORE_STATE
#undef QM_LOAD_DELTA
#undef QM_STORE_DELTA
#undef QPT
#undef QM_NODE_LOOP
#undef QM_LANES
#undef QM_CONVERGED

}  // namespace sdfnmpc
#endif  // __CUDACC__
