// PoseUKF fused step kernel ("panel + stream"; DESIGN.md sections 3-6 give the algebra, the layout and the measurements).
//
// Same mathematics as pose_engine.cuh (prefix-structured unscented transform) with a different residency plan: every op of
// the fast path (predict k=15, DVL k=9, XY/Z k<=3, acceleration k=19, apply_delta k=6) reads and writes only the first KP
// columns of Sigma -- the PANEL, a prefix of the column-packed block, staged in shared memory one filter ahead by cp.async
// together with mu and the filter's inputs.  The trailing (53-KP)^2/2 entries are touched by exactly one elementwise expression,
//     Sigma_new[a][b] = f_a f_b Sigma_old[a][b] + dt^2 Q[a][b] - sum_c (K S)[a][c] K[b][c],
// so they never enter shared memory: after the panel work each warp streams them HBM -> registers -> HBM once, column by
// column (trailing_pass).  One warp per filter; a lane owns the covariance rows x0 = lane and x1 = lane + 32 in EVERY phase:
//   * chol_panel: Cholesky of the leading block carried over the whole panel, both rows of the panel factor in registers --
//     cross-covariances are dot products of registers with broadcast rows of the sigma-point deviations (no triangular
//     solves, no re-read of Sigma);
//   * process_panel / downdate / trailing_pass: row factors (f_a, (K S)_a) in registers, column factors ({K_b, f_b}, one
//     AoS table per warp) as broadcast 128-bit shared loads;
//   * lanes double as sigma points (2k+1 <= 39 distinct points, the rest is a multiplicity of the centre) for the models;
//     warp sums run on the FP64 tensor path (wsum).
#pragma once
#include "pose_engine.cuh"
#include <type_traits>

namespace uwvkf {
namespace pose {

constexpr int M_NONE = -1;

#ifndef UWVKF_ILV                 // 1: trailing pass of the predict+update kernels interleaved with apply_delta (TrailStream)
#define UWVKF_ILV 1
#endif
#ifndef UWVKF_ILV_UPDATE_ONLY     // 1: the update-only kernels interleave their trailing pass as well
#define UWVKF_ILV_UPDATE_ONLY 1
#endif
#ifndef UWVKF_TMA_PREFIX          // 1: the [mu | panel] prefix arrives by ONE bulk copy (TMA engine, cp.async.bulk + mbarrier) per filter
#define UWVKF_TMA_PREFIX 1
#endif
#ifndef UWVKF_TMA_TRAIL           // 1: the interleaved trailing pass reads from a staging buffer filled by one bulk copy per filter
#define UWVKF_TMA_TRAIL 1
#endif
#ifndef UWVKF_TMERC_PREPASS       // 1: the latitudes of the predict's sigma points come from a thread-per-filter pre-pass kernel (32 filters per warp instruction instead of one)
#define UWVKF_TMERC_PREPASS 1
#endif
#ifndef UWVKF_TMERC_EARLY         // 1: the predict evaluates the projection of its sigma points INSIDE the panel Cholesky (branch-free, one stage per column from column 1 on): parity green, chol15 grows by what the sigma-point phase loses (the other warps were already filling those slots), 2 % slower, off
#define UWVKF_TMERC_EARLY 0
#endif
#ifndef UWVKF_SHRINK              // 1: predict kernels shrink the panel to 9 columns after the predict's factorisation (FusedCfg::KT): parity green, 2 % slower, off
#define UWVKF_SHRINK 0
#endif
#ifndef UWVKF_TWO_BUF             // 1: kernels with a predict keep TWO [mu | panel] buffers per warp (incoming / working): parity green, 1.5 % slower, off
#define UWVKF_TWO_BUF 0
#endif
#ifndef UWVKF_DVL_TAIL            // 1: DVL update as prefix 6 + affine tail (velocity): parity green, 2 % slower than prefix 9
#define UWVKF_DVL_TAIL 0
#endif

// Measurement traits of the fused kernel: prefix K (tangent coordinates the model reads arbitrarily) plus an AFFINE TAIL
// of NT tangent coordinates tail(0..NT-1) (ascending, all >= K) that the model reads affinely once the prefix is fixed
// (slot(t) maps a tail coordinate back to its position in that list).
// Sigma columns j >= K move only the tail, Z_j(+-) = Z_0 +- C L[T, j], so their sums collapse to Schur complements
// (tests/structured_ref.py::pose_update_tail states the algebra):
//   S_zz = 1/2 sum_{centre, j<K} w dz dz^T + C (Sigma_TT - L_T L_T^T) C^T + R,
//   P_xz = Sigma[:, :K] L11^-T (B - L_T^T C^T) + Sigma[:, T] C^T,         L_T = Sigma[T, :K] L11^-T.
template <int KIND> struct FMeas : Meas<KIND> {
  static constexpr int NT = 0;
  static constexpr int KPMIN = 0;   // smallest panel the kernel must keep in shared memory (tail coordinates it wants resident)
  __host__ __device__ static constexpr int tail(int) { return 0; }
  __host__ __device__ static constexpr int slot(int) { return 0; }
  __device__ static void tail_jac(const double *, const double *, const Params &, double *) {}
};
// The accelerometer model q^-1 (a + g e_z) + b_a (src/PoseUKF.cpp:125-131): prefix 6 (the orientation), tail = acceleration,
// accelerometer bias, gravity.  The panel is kept 19 columns wide so that every tail row / column is resident in shared memory;
// what shrinks is the factorisation (6 instead of 19 columns) and the sigma-point set (13 instead of 39 points).
template <> struct FMeas<M_ACCELERATION> {
  static constexpr int K = 6, M = 3, NT = 7, KPMIN = 19;
  __host__ __device__ static constexpr int tail(int r) { return r < 3 ? T_ACC + r : (r < 6 ? T_BA + (r - 3) : T_G); }
  __host__ __device__ static constexpr int slot(int t) { return t < T_BG ? t - T_ACC : (t < T_G ? t - T_BA + 3 : 6); }
  template <class D> __device__ static void eval(const double *mu, D d, const double *aux, const Params &P, double *o) {
    Meas<M_ACCELERATION>::eval(mu, d, aux, P, o);
  }
  __device__ static void tail_jac(const double *mu, const double *, const Params &, double *C) {   // C[c * NT + r]
    double Rm[9];
    qmat(Quat{mu[P_ORI], mu[P_ORI + 1], mu[P_ORI + 2], mu[P_ORI + 3]}, Rm);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        C[c * 7 + r] = Rm[r * 3 + c];             // d/d acceleration: R^T
        C[c * 7 + 3 + r] = (c == r) ? 1.0 : 0.0;   // d/d bias
      }
      C[c * 7 + 6] = Rm[2 * 3 + c];               // d/d gravity: R^T e_z
    }
  }
};
#if UWVKF_DVL_TAIL
// DVL model q^-1 v (src/PoseUKF.cpp:117-123): prefix 6 (position, orientation), tail = velocity (affine once q is fixed)
template <> struct FMeas<M_VELOCITY> {
  static constexpr int K = 6, M = 3, NT = 3, KPMIN = 9;
  __host__ __device__ static constexpr int tail(int r) { return T_VEL + r; }
  __host__ __device__ static constexpr int slot(int t) { return t - T_VEL; }
  template <class D> __device__ static void eval(const double *mu, D d, const double *aux, const Params &P, double *o) {
    Meas<M_VELOCITY>::eval(mu, d, aux, P, o);
  }
  __device__ static void tail_jac(const double *mu, const double *, const Params &, double *C) {   // C[c * NT + r] = (R^T)[c][r]
    double Rm[9];
    qmat(Quat{mu[P_ORI], mu[P_ORI + 1], mu[P_ORI + 2], mu[P_ORI + 3]}, Rm);
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int r = 0; r < 3; ++r) C[c * 3 + r] = Rm[r * 3 + c];
  }
};
#endif
template <> struct FMeas<M_PRESSURE> {   // src/PoseUKF.cpp:107-115: prefix (z, q, gravity) = 19, tail = water density
  static constexpr int K = 19, M = 1, NT = 1, KPMIN = 0;
  __host__ __device__ static constexpr int tail(int) { return T_RHO; }
  __host__ __device__ static constexpr int slot(int) { return 0; }
  template <class D> __device__ static void eval(const double *mu, D d, const double *aux, const Params &P, double *o) {
    Meas<M_PRESSURE>::eval(mu, d, aux, P, o);
  }
  __device__ static void tail_jac(const double *mu, const double *aux, const Params &, double *C) {
    Vec3 rn = qrot(Quat{mu[P_ORI], mu[P_ORI + 1], mu[P_ORI + 2], mu[P_ORI + 3]}, Vec3{aux[0], aux[1], aux[2]});
    C[0] = -((mu[P_POS + 2] + rn.z) * mu[P_G]);
  }
};
template <> struct FMeas<M_WATER_VELOCITY> {   // src/PoseUKF.cpp:133-151: prefix (q, v) = 9, tail = wv, wv_below, bias_adcp
  static constexpr int K = 9, M = 2, NT = 6, KPMIN = 0;
  __host__ __device__ static constexpr int tail(int r) { return T_WV + r; }
  __host__ __device__ static constexpr int slot(int t) { return t - T_WV; }
  template <class D> __device__ static void eval(const double *mu, D d, const double *aux, const Params &P, double *o) {
    Meas<M_WATER_VELOCITY>::eval(mu, d, aux, P, o);
  }
  __device__ static void tail_jac(const double *mu, const double *aux, const Params &, double *C) {   // C[c * NT + r]
    double Rm[9];
    qmat(Quat{mu[P_ORI], mu[P_ORI + 1], mu[P_ORI + 2], mu[P_ORI + 3]}, Rm);
    const double wgt = aux[0];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {   // (R^T)[c][r] = R[r][c]
        C[c * 6 + r] = -(1 - wgt) * Rm[r * 3 + c];
        C[c * 6 + 2 + r] = -wgt * Rm[r * 3 + c];
        C[c * 6 + 4 + r] = (c == r) ? 1.0 : 0.0;
      }
    }
  }
};

// KIND2 != M_NONE: a SECOND update in the same pass over the filter (predict + DVL + accelerometer, the 100 Hz step of BASELINE
// config 3): the panel stays in shared memory between the updates, the trailing triangle is rewritten once, with both downdates.
template <int HAS_PREDICT, int KIND, int KIND2 = M_NONE>
struct FusedCfg {
  using MF = FMeas<KIND == M_NONE ? M_XY : KIND>;
  using MF2 = FMeas<KIND2 == M_NONE ? M_XY : KIND2>;
  static_assert(KIND2 == M_NONE || KIND != M_NONE, "a second update needs a first");
  static constexpr int KM1 = (KIND == M_NONE) ? 0 : MF::K;
  static constexpr int KM2 = (KIND2 == M_NONE) ? 0 : MF2::K;
  static constexpr int KM = KM1 > KM2 ? KM1 : KM2;
  static constexpr int NT1 = (KIND == M_NONE) ? 0 : MF::NT;
  static constexpr int NT2 = (KIND2 == M_NONE) ? 0 : MF2::NT;
  static constexpr int NT = NT1 > NT2 ? NT1 : NT2;
  static constexpr int KA = (KIND == M_NONE) ? 0 : 6;                       // apply_delta prefix
  static constexpr int K0 = HAS_PREDICT ? 15 : 0;
  static constexpr int KPM1 = (KIND == M_NONE) ? 0 : MF::KPMIN;
  static constexpr int KPM2 = (KIND2 == M_NONE) ? 0 : MF2::KPMIN;
  static constexpr int KPM = KPM1 > KPM2 ? KPM1 : KPM2;
  static constexpr int KP0 = (K0 > KM ? (K0 > KA ? K0 : KA) : (KM > KA ? KM : KA));
  static constexpr int KP = KP0 > KPM ? KP0 : KPM;                          // panel width
  static constexpr int PANEL = KP * N - (KP * (KP - 1)) / 2;                // doubles in the panel
  static constexpr int PANEL_E = ((PANEL + 1) / 2) * 2;
  // Kernels with a predict need 15 columns for its factorisation, but columns 9..14 hold affine x affine entries only and
  // have no coupled partner columns: past the predict they obey the elementwise expression of the trailing triangle.  When the
  // update (if any) is content with 9 columns and reads no tail coordinates, the panel therefore SHRINKS to KT = 9 columns after
  // the factorisation: the panel transform, the panel downdate and the write-back stop there, and the trailing pass starts at
  // column 9, reading the old values of columns 9..14 from HBM (L2-hot: the prefix fetch has just pulled them through).
  static constexpr int KU = (KM > KA ? KM : KA) > 9 ? (KM > KA ? KM : KA) : 9;
  static constexpr bool SHRINK = UWVKF_SHRINK && HAS_PREDICT && NT == 0 && KPM == 0 && KIND2 == M_NONE && KU < KP;
  static constexpr int KT = SHRINK ? KU : KP;                                // first column of the trailing pass
  static constexpr int PANEL_T = KT * N - (KT * (KT - 1)) / 2;
  static constexpr int KL = KP;                                             // largest leading factor
  static constexpr int LSZ = ((KL * (KL | 1) + 1) / 2) * 2;        // square, odd column stride
  static constexpr int KE = ((KL + 1) / 2) * 2;
  // per-warp shared layout (doubles)
  static constexpr int O_MU = 0;                   // 56
  static constexpr int O_S = 56;                   // panel (same indexing as the full packed array)
  // Kernels with a predict keep TWO [mu | panel] buffers per warp: the INCOMING one (O_A), which cp.async fills one filter
  // ahead, and the WORKING one (O_MU / O_S).  The predict reads the old panel from the first and writes F Sigma F^T + dt^2 Q
  // into the second, so the panel transform needs no barriers (nothing is updated in place) and the incoming buffer is free --
  // and refilled with the next filter's prefix -- as soon as the predict is done, a whole update ahead of its use.
  static constexpr bool TWO = HAS_PREDICT && UWVKF_TWO_BUF && KP <= 15;   // (a 19-column panel twice would cost two warps per SM)
  static constexpr int PBUF = 56 + PANEL_E;        // doubles per prefix buffer
  static constexpr int O_A = TWO ? PBUF : 0;       // incoming buffer (the working buffer itself without a predict)
  static constexpr int O_L = PBUF + (TWO ? PBUF : 0);
  static constexpr int O_G = O_L + LSZ;          // per sigma-point pair j: {A[0][j], A[1][j], A[2][j], -}, AoS [KE][4]
  static constexpr int O_KC = O_G + 4 * KE;        // per column b: {K[b][0], K[b][1], K[b][2], f_b}, AoS [54][4] (two broadcast
                                                   // 128-bit loads hand a whole warp everything column b contributes); with a
                                                   // second update [54][8]: + {K2[b][0], K2[b][1], K2[b][2], -}
  static constexpr int KCW = (KIND2 == M_NONE) ? 4 : 8;
  static constexpr int O_LT = O_KC + KCW * 54;         // rows T of the panel factor, NT x KE
  // per-filter inputs staged with the prefix: rotation rate [0..2], z [4..4+M), z of the second update [8..8+M2); two sets,
  // alternating per filter (the next filter's inputs arrive while the current filter's update still reads its own)
  static constexpr int INB = (KIND2 == M_NONE) ? 8 : 12;              // rotation rate and measurement slots
  static constexpr bool TMP = UWVKF_TMERC_PREPASS && HAS_PREDICT;    // + sin / cos of the latitude of sigma points 0..4 (pre-pass)
  static constexpr int INW = INB + (TMP ? 10 : 0);
  static constexpr int O_IN = O_LT + NT * KE;
  static constexpr int O_MB = O_IN + 2 * INW;      // mbarriers of the bulk (TMA) copies: [0] prefix fetch, [1] trailing triangle
  // The interleaved trailing pass reads the old trailing triangle from a shared-memory STAGING buffer that ONE bulk copy fills at
  // the top of the filter's iteration (UWVKF_TMA_TRAIL): no loads in flight in registers across the phases of apply_delta, no
  // L2 prefetch.  The copy starts at the even entry at or below the panel's end (16-byte alignment) and runs to the block's end.
  static constexpr bool TRS = UWVKF_TMA_TRAIL && UWVKF_TMA_PREFIX;
  static constexpr int TR_START = PANEL_T & ~1;                     // first staged entry (packed index): where the trailing pass starts
  static constexpr int TR_LEN = TRS ? (BLK - OFF_S) - TR_START : 0;  // doubles staged (through the pad at the end of the block)
  static constexpr int O_TR = O_MB + 2;
  static constexpr int TOTAL = O_TR + TR_LEN;
  // per-CTA shared region (doubles): dt^2 Q packed (predict only), parameter tables
  static constexpr int C_Q = 0;
  static constexpr int C_FTAB = HAS_PREDICT ? 1432 : 0;       // affine factor per tangent index
  static constexpr int C_MTAU = C_FTAB + 56;       // -1/tau per stored index (0: not a Gauss-Markov state)
  static constexpr int C_OTAB = C_MTAU + 56;       // shared Markov offsets per stored index (bias offsets)
  static constexpr int CTA = C_OTAB + 56;
};

// Tuning switches (tools/build_variant.sh builds variants; DESIGN.md section 6 lists what was measured)
#ifndef UWVKF_TUB      // trailing-pass pipeline of the kernels with a predict: columns per batch, batches in flight
#define UWVKF_TUB 6
#define UWVKF_TDEPTH 2
#endif
#ifndef UWVKF_TRAIL_PREFETCH_L1   // 1: pull the trailing triangle into L1 instead of L2 at the top of a filter (measured: no gain)
#define UWVKF_TRAIL_PREFETCH_L1 0
#endif
#ifndef UWVKF_OPAQUE_LANE         // 1: lane index opaque per filter (no loop-invariant hoisting of lane-derived values)
#define UWVKF_OPAQUE_LANE 0
#endif
#ifndef UWVKF_TRAIL_OPAQUE        // 1: lane index opaque inside the trailing passes only
#define UWVKF_TRAIL_OPAQUE 0
#endif
#ifndef UWVKF_COLD_CALL           // 1: the classic trailing pass of the ILV kernels (refused filters only) is an out-of-line call
#define UWVKF_COLD_CALL 0
#endif
#ifndef UWVKF_ILV_UNCLAMPED       // 1: TrailStream addresses rows without the per-column clamp (no lane-dependent column offsets)
#define UWVKF_ILV_UNCLAMPED 1
#endif
#ifndef UWVKF_LANE_PINNED         // 1: the lane index cannot be re-materialised from SR_TID inside the filter loop
#define UWVKF_LANE_PINNED 0
#endif
#ifndef UWVKF_ILV_EARLY           // 1: request the first batches before the panel downdate instead of after it
#define UWVKF_ILV_EARLY 0
#endif
#ifndef UWVKF_ILV_PREDICT         // 1: the predict-only kernel interleaves its trailing pass with the predict's own phases: parity green, 4 % slower (spills), off
#define UWVKF_ILV_PREDICT 0
#endif
#ifndef UWVKF_ILV_UB
#define UWVKF_ILV_UB 8
#define UWVKF_ILV_DEPTH 2
#endif
// Per-phase latency probe (debug builds only, -DUWVKF_PHASE_TIMING; tools/phase_timing.py): warp 0 of every CTA accumulates the
// clock64() span of each phase of its filters; uwvkf_debug_phase_cycles reads the table.
#ifdef UWVKF_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[40];
#define PH_DECL long long ph_t = clock64();
#define PH(i) do { __syncwarp(); const long long t_ = clock64(); if (threadIdx.x == 0) atomicAdd(&g_phase_cycles[i], (unsigned long long)(t_ - ph_t)); ph_t = t_; } while (0)
#define PH_PARAM , long long &ph_t, int phb
#define PH_ARG(b) , ph_t, b
#define PHB(i) PH(phb + (i))
#else
#define PH_DECL
#define PH(i)
#define PH_PARAM
#define PH_ARG(b)
#define PHB(i)
#endif

// st.global.cs under a predicate the compiler cannot turn into a branch around the code that produces the value
__device__ __forceinline__ void stcs_if(bool p, double *ptr, double v) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.cs.f64 [%0], %1;\n\t}" ::"l"(ptr), "d"(v), "r"((unsigned)p) : "memory");
}

__device__ __forceinline__ double ldcs_v(const double *ptr) {   // ld.global.cs whose issue order the compiler must keep
  double v;
  asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(v) : "l"(ptr) : "memory");
  return v;
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__device__ __forceinline__ int coloff(int j) { return j * N - ((j * (j - 1)) >> 1) - j; }   // cidx(i, j) = coloff(j) + i


using ::uwvkf::wsum;   // warp-wide sum on the FP64 tensor path (dmath.cuh)

// Leading-block Cholesky factor in SQUARE column-major storage with an odd column stride (conflict-free both for
// "lane = row" and for "lane = column" accesses): L[i][j] = Lw[j * LStride<K> + i].
template <int K> constexpr int LStride = K | 1;

// Cholesky of the leading K x K block carried over the WHOLE panel: besides the factor L11 (to shared memory, for the sigma
// points) every lane ends up with the rows x0 = lane and x1 = lane + 32 of the panel factor Lp = Sigma[:, :K] L11^-T in registers.
// The rows below the block ride along in the rank-1 updates (lanes >= K were idle in them anyway; the x1 rows cost K(K+1)/2 more
// DFMA), and they replace both the triangular solve G = A L11^-1 and the product G Sigma[:K, :] of the unscented transform:
//   A L11^-1 Sigma[:K, x] = A (Lp[x, :])^T                       (Sigma[:K, x] = L11 Lp[x, :]^T),
// i.e. cross-covariances become K-term dot products of registers with broadcast rows of A: no solve, no re-read of Sigma.
struct NoHook { __device__ __forceinline__ void col(int, double) const {} };
// `hook.col(j)` runs once column J of the factor is in shared memory (the predict evaluates its transverse-Mercator projection
// there, one stage per column from column 1 on: the only sigma points whose position differs from the mean's are those of
// columns 0 and 1)
template <int K, class Hook = NoHook>
__device__ __forceinline__ bool chol_panel(const double *S, double *Lw, int lane, double (&a0)[K], double (&a1)[K], Hook &&hook = Hook{}) {
  static_assert(K <= 32, "one row per lane");
  constexpr int LS = LStride<K>;
  const int x1 = lane + 32 < N ? lane + 32 : lane;   // lanes without a second row duplicate the first (never consumed)
#pragma unroll
  for (int c = 0; c < K; ++c) { a0[c] = S[coloff(c) + lane]; a1[c] = S[coloff(c) + x1]; }
  // pivots: lane i's diagonal entry only ever receives -L[i][j]^2, products of the lane's own entries, so it is tracked in a
  // register of its own and never waits for the shared-memory round trip that carries the finished column to the other lanes
  double dg = S[cidx(lane < K ? lane : 0, lane < K ? lane : 0)];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const double d = bcast(dg, j);
    ok = ok && (d > 0.0);
    const double inv = fast_rsqrt3(d);
    a0[j] *= inv;                           // lane j: d / sqrt(d) = L[j][j]; other rows: L[i][j]
    a1[j] *= inv;
    dg = fma(-a0[j], a0[j], dg);
    if (lane < K) Lw[j * LS + lane] = a0[j];
    __syncwarp();
    hook.col(j, inv);   // (j is a constant after unrolling: the hook's switch folds away)
    {
      int c = j + 1;
      if (((j * LS + c) & 1) && c < K) {
        const double l = Lw[j * LS + c];
        a0[c] = fma(-a0[j], l, a0[c]); a1[c] = fma(-a1[j], l, a1[c]); ++c;
      }
#pragma unroll
      for (; c + 1 < K; c += 2) {
        const double2 l2 = *reinterpret_cast<const double2 *>(Lw + j * LS + c);
        a0[c] = fma(-a0[j], l2.x, a0[c]); a1[c] = fma(-a1[j], l2.x, a1[c]);
        a0[c + 1] = fma(-a0[j], l2.y, a0[c + 1]); a1[c + 1] = fma(-a1[j], l2.y, a1[c + 1]);
      }
      if (c < K) {
        const double l = Lw[j * LS + c];
        a0[c] = fma(-a0[j], l, a0[c]); a1[c] = fma(-a1[j], l, a1[c]);
      }
    }
  }
  // rows inside the block: the entries above the diagonal are zeros of L11 (the loop carried don't-care values there)
#pragma unroll
  for (int j = 1; j < K; ++j) a0[j] = (j > lane) ? 0.0 : a0[j];
  return ok;
}

// T[h][c] = sum_j A[c][j] Lp[x_h][j]: rows of the panel factor from registers, A (M x K, stored [j][4]) as broadcast loads
template <int K, int M>
__device__ __forceinline__ void a_times_panel(const double *G, const double (&a0)[K], const double (&a1)[K], double (&T)[2][M]) {
#pragma unroll
  for (int c = 0; c < M; ++c) { T[0][c] = 0.0; T[1][c] = 0.0; }
#pragma unroll
  for (int j = 0; j < K; ++j) {
    double aj[3];
    const double2 g01 = *reinterpret_cast<const double2 *>(G + 4 * j);
    aj[0] = g01.x; aj[1] = g01.y; aj[2] = 0.0;
    if (M > 2) aj[2] = G[4 * j + 2];
#pragma unroll
    for (int c = 0; c < M; ++c) { T[0][c] = fma(aj[c], a0[j], T[0][c]); T[1][c] = fma(aj[c], a1[j], T[1][c]); }
  }
}

// equal-weight iterated SO3 mean over 2N+1 logical points, one point (with multiplicity wt) per lane
__device__ __forceinline__ Quat so3_mean_lane(const Quat &qp, double wt, double tol, int max_it, unsigned &status) {
  Quat r{bcast(qp.x, 0), bcast(qp.y, 0), bcast(qp.z, 0), bcast(qp.w, 0)};
  int it = 0;
  double nrm2;
  const double tol2 = tol * tol;
  do {
    const Vec3 d = so3_boxminus(qp, r);
    constexpr double kInvPts = 1.0 / double(2 * N + 1);   // mean over the 2n+1 points (one rounding away from a division)
    const double mx = wsum(wt * d.x) * kInvPts;
    const double my = wsum(wt * d.y) * kInvPts;
    const double mz = wsum(wt * d.z) * kInvPts;
    r = so3_boxplus(r, Vec3{mx, my, mz}, 1.0);
    nrm2 = mx * mx + my * my + mz * mz;
  } while (nrm2 > tol2 && ++it < max_it);
  if (it >= max_it) status |= ST_MEAN_NOT_CONVERGED;
  return r;
}

// Unscented transform of the orientation rows (see ori_transform in pose_engine.cuh) in three parts, so that a caller can slot
// independent work between them (apply_delta interleaves batches of the trailing pass): the iterated mean, the deviations with
// S_oo and A (to shared memory), and T = A Lp^T from the panel-factor rows in registers.
__device__ __forceinline__ double ori_weight(int p, int K) { return p == 0 ? double(1 + 2 * (N - K)) : (p < 2 * K + 1 ? 1.0 : 0.0); }
template <int K>
__device__ __forceinline__ void ori_deviations(double *G, const Quat &qp, const Quat &r, int p, double soo[6]) {
  const bool on = p > 0 && p < 2 * K + 1;
  const int j = (p - 1) >> 1;
  const double wa = ori_weight(p, K);
  Vec3 a = so3_boxminus(qp, r);
  const double hw = 0.5 * wa;
  soo[0] = wsum(hw * a.x * a.x);
  soo[1] = wsum(hw * a.y * a.x);
  soo[2] = wsum(hw * a.z * a.x);
  soo[3] = wsum(hw * a.y * a.y);
  soo[4] = wsum(hw * a.z * a.y);
  soo[5] = wsum(hw * a.z * a.z);
  double nx = __shfl_down_sync(kFull, a.x, 1), ny = __shfl_down_sync(kFull, a.y, 1), nz = __shfl_down_sync(kFull, a.z, 1);
  if (on && (p & 1)) {
    *reinterpret_cast<double2 *>(G + 4 * j) = make_double2(0.5 * (a.x - nx), 0.5 * (a.y - ny));
    G[4 * j + 2] = 0.5 * (a.z - nz);
  }
  __syncwarp();
}
template <int K>
__device__ __forceinline__ Quat ori_transform_reg(double *G, const double (&a0)[K], const double (&a1)[K], const Quat &qp,
                                                  int p, const Params &P, double soo[6], double (&T)[2][3], int lane, unsigned &status PH_PARAM) {
  Quat r = so3_mean_lane(qp, ori_weight(p, K), P.mean_tol, P.mean_max_it, status);
  PHB(0);
  ori_deviations<K>(G, qp, r, p, soo);
  PHB(1);
  PHB(2);
  a_times_panel<K, 3>(G, a0, a1, T);
  PHB(3);
  return r;
}

// Panel part of Sigma <- F Sigma F^T + dt^2 Q for the affine rows, F = diag(f) + dt E, E: position <- velocity, velocity <-
// acceleration (the two couplings commute to first order: E_va E_pv = 0, so F is exactly the product the sequential couple3_f
// calls apply).  An entry (x, b) reads at most four old entries -- (x, b), (x', b), (x, b'), (x', b') with x' / b' the coupled
// partner of its row / column -- and partners only point "down" the chain p -> v -> a.  Hence: the three position columns first
// (new values into registers, one barrier, write), then the three velocity columns the same way, then every other column in
// place.  Two barriers instead of six and no read-modify-write chains through shared memory.  Orientation rows and columns are
// left alone (the caller overwrites them with the unscented transform).
template <int KP>   // KP: columns to transform (>= 9: the coupled position / velocity groups are always inside)
__device__ __forceinline__ void process_panel(double *S, const double *Qs, const double *ftab, const double (&f)[2], double dt, int lane,
                                              int colane) {
  static_assert(KP >= 9, "the coupled column groups must be transformed here");
  const double dt2 = dt * dt;
  const bool rp = lane < 3 || (lane >= T_VEL && lane < T_ACC);   // row x0 = lane has a coupled partner (rows x1 never do)
  const int px = lane + (lane < 3 ? 6 : 3);
  const int copx = coloff(px);
  auto group = [&](auto b0_, auto pbo_) {
    constexpr int b0 = decltype(b0_)::value, pbo = decltype(pbo_)::value;
    double nv[3][2];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      constexpr int dummy = 0; (void)dummy;
      const int b = b0 + i, pb = b + pbo, cb = coloff(b), cpb = coloff(pb);
      const double fb = ftab[b];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int x = lane + 32 * h;
        const int cox = h ? coloff(x < N ? x : 0) : colane;
        // every load is unconditional (inactive lanes read other entries of the warp's own shared-memory region and drop the
        // result): a branch per entry would put one shared-memory round trip per entry on the critical path
        double val = fma(f[h] * fb, S[cb + x], Qs[cb + x]);
        val = fma(dt * f[h], S[x >= pb ? cpb + x : cox + pb], val);
        if (h == 0) {
          const double t1 = S[cb + px], t3 = S[px >= pb ? cpb + px : copx + pb];
          val = rp ? fma(dt2, t3, fma(dt * fb, t1, val)) : val;
        }
        nv[i][h] = val;
      }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int b = b0 + i, x = lane + 32 * h;
        if (x >= b && x < N && !(x >= 3 && x < 6)) S[coloff(b) + x] = nv[i][h];
      }
    __syncwarp();
  };
  group(std::integral_constant<int, T_POS>{}, std::integral_constant<int, 6>{});
  group(std::integral_constant<int, T_VEL>{}, std::integral_constant<int, 3>{});
#pragma unroll
  for (int b = T_ACC; b < KP; ++b) {
    const int cb = coloff(b);
    const double fb = ftab[b];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int x = lane + 32 * h;
      if (x >= b && x < N) S[cb + x] = fma(f[h] * fb, S[cb + x], Qs[cb + x]);
    }
  }
}

// The same map OUT OF PLACE (kernels with two prefix buffers): old entries are read from SA, new ones written to S.  Nothing is
// updated in place, so there is no barrier and no staging through registers: every entry is an independent load / FMA / store
// chain that the scheduler is free to slot between the sigma-point arithmetic that follows.
template <int KP>
__device__ __forceinline__ void process_panel_out(const double *SA, double *S, const double *Qs, const double *ftab, const double (&f)[2],
                                                  double dt, int lane, int colane) {
  const double dt2 = dt * dt;
  const bool rp = lane < 3 || (lane >= T_VEL && lane < T_ACC);   // row x0 = lane has a coupled partner (rows x1 never do)
  const int px = lane + (lane < 3 ? 6 : 3);
  const int copx = coloff(px);
  auto group = [&](auto b0_, auto pbo_) {
    constexpr int b0 = decltype(b0_)::value, pbo = decltype(pbo_)::value;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int b = b0 + i, pb = b + pbo, cb = coloff(b), cpb = coloff(pb);
      const double fb = ftab[b];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int x = lane + 32 * h;
        const int cox = h ? coloff(x < N ? x : 0) : colane;
        double val = fma(f[h] * fb, SA[cb + x], Qs[cb + x]);
        val = fma(dt * f[h], SA[x >= pb ? cpb + x : cox + pb], val);
        if (h == 0) {
          const double t1 = SA[cb + px], t3 = SA[px >= pb ? cpb + px : copx + pb];
          val = rp ? fma(dt2, t3, fma(dt * fb, t1, val)) : val;
        }
        if (x >= b && x < N && !(x >= 3 && x < 6)) S[cb + x] = val;
      }
    }
  };
  group(std::integral_constant<int, T_POS>{}, std::integral_constant<int, 6>{});
  group(std::integral_constant<int, T_VEL>{}, std::integral_constant<int, 3>{});
#pragma unroll
  for (int b = T_ACC; b < KP; ++b) {
    const int cb = coloff(b);
    const double fb = ftab[b];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int x = lane + 32 * h;
      const double val = fma(f[h] * fb, SA[cb + x], Qs[cb + x]);
      if (x >= b && x < N) S[cb + x] = val;
    }
  }
}

__device__ const double kNoKs[2][3] = {{0, 0, 0}, {0, 0, 0}};

// Trailing triangle (columns >= KP of the packed covariance, a contiguous suffix of the filter block), streamed HBM -> registers
// -> HBM once:   s <- f_a f_b s + dt^2 Q_ab - sum_c (K S)_ac K_bc.
// Rows keep the owner they have everywhere else in the kernel (x0 = lane, x1 = lane + 32), so the row factors f_a, (K S)_a live in
// REGISTERS; column b contributes {K_b, f_b} through two broadcast 128-bit shared loads, and the rows b..52 of a column are one
// contiguous run in memory.  The old values come from the shared-memory staging buffer sS (same indexing as gS) that cp.async
// filled while the panel work ran, so no global-memory latency is left on this path; the new values go straight to HBM.  A column needs one warp-wide access: the "primary" row of a lane is x1 on lanes 0..20 (rows 32..52) and
// x0 on lanes 21..31 (rows 21..31); columns b <= 20 need a second, short access for the rows b..20 ("secondary" row x0 of lanes
// <= 20).  No entry -> (row, column) table, no gathers, no dependent address arithmetic: every column is independent of every
// other, so the loads of a whole batch of columns are in flight before the first is consumed.
// M2 > 0: a second rank-M2 downdate (two-update kernels; KC rows are then 8 wide, K2 in slots 4..6, ks2 its row factors).
template <bool AP, bool AK, int KP, int M, int UB, int DEPTH, int M2 = 0>
__device__ __forceinline__ void trailing_pass(double *__restrict__ gS, const double *Qs, const double *KC, const double (&f)[2],
                                              const double (&ks)[2][3], double q2add, int lane, const double (&ks2)[2][3] = kNoKs) {
  constexpr int KCW = M2 > 0 ? 8 : 4;
  const bool lo = lane <= 20;
  const int rP = lo ? lane + 32 : lane;
  const double fP = lo ? f[1] : f[0];
  const double kP0 = lo ? ks[1][0] : ks[0][0], kP1 = lo ? ks[1][1] : ks[0][1], kP2 = lo ? ks[1][2] : ks[0][2];
  auto entry = [&](int b, int row, double sv, double fr, double k0, double k1, double k2) {
    const double2 c01 = *reinterpret_cast<const double2 *>(KC + KCW * b);
    const double2 c2f = *reinterpret_cast<const double2 *>(KC + KCW * b + 2);
    if (AP) {
      sv = fma(fr * c2f.y, sv, Qs[coloff(b) + row]);
      if (b >= T_WV && b < T_BADCP && row == b) sv += q2add;
    }
    if (AK) {
      sv = fma(-k0, c01.x, sv);
      if (M > 1) sv = fma(-k1, c01.y, sv);
      if (M > 2) sv = fma(-k2, c2f.x, sv);
    }
    if constexpr (M2 > 0) {   // second downdate; the row factors are re-selected here (cold path: refused filters only)
      const bool pr = (row >= 32);
      const double2 d01 = *reinterpret_cast<const double2 *>(KC + KCW * b + 4);
      sv = fma(-(pr ? ks2[1][0] : ks2[0][0]), d01.x, sv);
      if (M2 > 1) sv = fma(-(pr ? ks2[1][1] : ks2[0][1]), d01.y, sv);
      if (M2 > 2) sv = fma(-(pr ? ks2[1][2] : ks2[0][2]), KC[KCW * b + 6], sv);
    }
    return sv;
  };
  // Inactive lanes (rows above the diagonal of column b) read the diagonal entry instead and their store is predicated off, so
  // neither loads nor arithmetic sit behind a branch.  Loads and stores are volatile asm: their program order is the issue order,
  // which lets the loop below keep the loads of DEPTH - 1 further batches of UB columns in flight while a batch is computed and
  // stored (software pipeline; the triangle was pulled into L2 at the top of the filter's iteration, and an L2 hit still costs
  // several hundred cycles).  The register-hungry predict kernels run 6 x 2, the update-only kernels 8 x 3.
  constexpr int NB = (N - KP + UB - 1) / UB;
  constexpr int NS = KP <= 20 ? 21 - KP : 0;   // columns that also have secondary rows (b .. 20 on lanes <= 20)
  double vS[NS > 0 ? NS : 1];
#pragma unroll
  for (int u = 0; u < NS; ++u) vS[u] = ldcs_v(gS + coloff(KP + u) + ((lo && lane >= KP + u) ? lane : KP + u));
  double v[DEPTH][UB];
  auto load = [&](int i, double (&dst)[UB]) {
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int b = KP + i * UB + u;
      if (b < N) dst[u] = ldcs_v(gS + coloff(b) + (rP >= b ? rP : b));
    }
  };
  auto finish = [&](int i, const double (&src)[UB]) {
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int b = KP + i * UB + u;
      if (b < N) {
        const int rowP = rP >= b ? rP : b;
        stcs_if(rP >= b, gS + coloff(b) + rowP, entry(b, rowP, src[u], fP, kP0, kP1, kP2));
      }
    }
  };
#pragma unroll
  for (int i = 0; i < DEPTH - 1; ++i)
    if (i < NB) load(i, v[i]);
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    if (i + DEPTH - 1 < NB) load(i + DEPTH - 1, v[(i + DEPTH - 1) % DEPTH]);
    if (i == 0) {
#pragma unroll
      for (int u = 0; u < NS; ++u) {
        const int b = KP + u;
        const bool actS = lo && lane >= b;
        const int rowS = actS ? lane : b;
        stcs_if(actS, gS + coloff(b) + rowS, entry(b, rowS, vS[u], f[0], ks[0][0], ks[0][1], ks[0][2]));
      }
    }
    finish(i, v[i % DEPTH]);
  }
}

// Out-of-line copy for kernels whose common path does not come through here (UWVKF_ILV: only filters whose prediction or update was
// refused do): as a real call it keeps its column offsets and predicates out of the caller's register allocation.
template <bool AP, bool AK, int KP, int M, int UB, int DEPTH>
__device__ __noinline__ void trailing_pass_cold(double *gS, const double *Qs, const double *KC, double f0, double f1, double k00, double k01,
                                                double k02, double k10, double k11, double k12, double q2add, int lane) {
  const double f[2] = {f0, f1};
  const double ks[2][3] = {{k00, k01, k02}, {k10, k11, k12}};
  trailing_pass<AP, AK, KP, M, UB, DEPTH>(gS, Qs, KC, f, ks, q2add, lane);
}

// The same pass cut into batches of UB columns that the caller INTERLEAVES with the phases of apply_delta (UWVKF_ILV): the rank-M
// downdate of the trailing triangle only needs the gain tables (KC, ks), which exist before apply_delta starts, and apply_delta is
// a chain of dependent, issue-starved phases (a 6-column factorisation, 13 sigma points, an iterated SO3 mean) that touch neither
// the trailing triangle nor the gain tables.  The loads of a batch are issued before one phase and consumed after it, so their
// L2 latency (the triangle was prefetched at the top of the filter) hides behind that phase, and the FMAs / stores of the
// previous batch fill the issue slots the phase leaves empty.  Nothing is stored before the factorisation of apply_delta has
// succeeded (a failure rolls the update back: the batches in flight are dropped and the classic pass applies the process map only).
// STAGED: the old values come from a shared-memory copy of the triangle (tr[e] = entry e of the packed array; one TMA bulk copy
// per filter) and are read where they are consumed: load() does nothing and no value is in flight across a phase.
template <bool AP, int KP, int M, int UB, int DEPTH, int M2 = 0, bool STAGED = false>
struct TrailStream {
  static constexpr int KCW = M2 > 0 ? 8 : 4;
  const double *tr = nullptr;
  static constexpr int NB = (N - KP + UB - 1) / UB;
  static constexpr int NS = KP <= 20 ? 21 - KP : 0;
  double v[DEPTH][UB];
  double vS[NS > 0 ? NS : 1];
  // Nothing but the values in flight is kept here: the row factors are re-selected from (f, ks) at every use and the column
  // offsets recomputed from an opaque copy of the lane index (a value the compiler could prove loop invariant would be hoisted
  // out of the filter loop -- some 90 lane-dependent offsets -- and spilled to local memory, which this kernel serves at L2 latency)
  __device__ __forceinline__ static double entry(const double *Qs, const double *KC, double q2add, int b, int row, double sv, double fr,
                                                 double k0, double k1, double k2, double j0 = 0.0, double j1 = 0.0, double j2 = 0.0) {
    const double2 c01 = *reinterpret_cast<const double2 *>(KC + KCW * b);
    const double2 c2f = *reinterpret_cast<const double2 *>(KC + KCW * b + 2);
    if (AP) {
      sv = fma(fr * c2f.y, sv, Qs[coloff(b) + row]);
      if (b >= T_WV && b < T_BADCP && row == b) sv += q2add;
    }
    if (M > 0) sv = fma(-k0, c01.x, sv);   // M == 0: process map only (predict-only kernel)
    if (M > 1) sv = fma(-k1, c01.y, sv);
    if (M > 2) sv = fma(-k2, c2f.x, sv);
    if constexpr (M2 > 0) {   // downdate of the second update: K2 in slots 4..6 of the column table, (K2 S2) rows j0..j2
      const double2 d01 = *reinterpret_cast<const double2 *>(KC + KCW * b + 4);
      sv = fma(-j0, d01.x, sv);
      if (M2 > 1) sv = fma(-j1, d01.y, sv);
      if (M2 > 2) sv = fma(-j2, KC[KCW * b + 6], sv);
    }
    return sv;
  }
  template <int I> __device__ __forceinline__ void load(const double *gS, int lane) {
    if constexpr (I < NB && !STAGED) {
      const int rP = lane <= 20 ? lane + 32 : lane;
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int b = KP + I * UB + u;
#if UWVKF_ILV_UNCLAMPED
        // rows above the diagonal of column b read a neighbouring entry of the previous column (in range; value dropped)
        if (b < N) v[I % DEPTH][u] = ldcs_v(gS + rP + coloff(b));
#else
        if (b < N) v[I % DEPTH][u] = ldcs_v(gS + coloff(b) + (rP >= b ? rP : b));
#endif
      }
    }
  }
  __device__ __forceinline__ void load_secondary(const double *gS, int lane) {
    if constexpr (!STAGED) {
#pragma unroll
      for (int u = 0; u < NS; ++u) vS[u] = ldcs_v(gS + coloff(KP + u) + ((lane <= 20 && lane >= KP + u) ? lane : KP + u));
    }
  }
  template <int I> __device__ __forceinline__ void finish(bool on, double *gS, const double *Qs, const double *KC, const double (&f)[2],
                                                          const double (&ks)[2][3], double q2add, int lane,
                                                          const double (&ks2)[2][3] = kNoKs) {
    if constexpr (I < NB) {
      const bool lo = lane <= 20;
      const int rP = lo ? lane + 32 : lane;
      const double fP = lo ? f[1] : f[0];
      const double kP0 = lo ? ks[1][0] : ks[0][0], kP1 = lo ? ks[1][1] : ks[0][1], kP2 = lo ? ks[1][2] : ks[0][2];
      const double jP0 = lo ? ks2[1][0] : ks2[0][0], jP1 = lo ? ks2[1][1] : ks2[0][1], jP2 = lo ? ks2[1][2] : ks2[0][2];
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        const int b = KP + I * UB + u;
        if (b < N) {
#if UWVKF_ILV_UNCLAMPED
          const int rowP = rP;
#else
          const int rowP = rP >= b ? rP : b;
#endif
#if UWVKF_EXP_NODEP      // experiment: the stored value does not depend on the loaded one (WRONG results; timing only)
          stcs_if(on && rP >= b, gS + coloff(b) + rowP, entry(Qs, KC, q2add, b, rowP, fP, fP, kP0, kP1, kP2));
#elif UWVKF_EXP_NOSTORE  // experiment: no stores (WRONG results; timing only)
          if (entry(Qs, KC, q2add, b, rowP, v[I % DEPTH][u], fP, kP0, kP1, kP2) == 1.2345e-300) stcs_if(on && rP >= b, gS + coloff(b) + rowP, 0.0);
#else
          const double old = STAGED ? tr[coloff(b) + rowP] : v[I % DEPTH][u];
          stcs_if(on && rP >= b, gS + coloff(b) + rowP, entry(Qs, KC, q2add, b, rowP, old, fP, kP0, kP1, kP2, jP0, jP1, jP2));
#endif
        }
      }
    }
  }
  __device__ __forceinline__ void finish_secondary(bool on, double *gS, const double *Qs, const double *KC, const double (&f)[2],
                                                   const double (&ks)[2][3], double q2add, int lane,
                                                   const double (&ks2)[2][3] = kNoKs) {
#pragma unroll
    for (int u = 0; u < NS; ++u) {
      const int b = KP + u;
      const bool actS = lane <= 20 && lane >= b;
      const int rowS = actS ? lane : b;
      const double old = STAGED ? tr[coloff(b) + rowS] : vS[u];
      stcs_if(on && actS, gS + coloff(b) + rowS,
              entry(Qs, KC, q2add, b, rowS, old, f[0], ks[0][0], ks[0][1], ks[0][2], ks2[0][0], ks2[0][1], ks2[0][2]));
    }
  }
};

// The same pass with the old values in the shared-memory staging copy of the triangle (tr[e] = entry e of the packed array, one
// TMA bulk copy per filter): nothing to pipeline, every column is a shared-memory read, the arithmetic and a predicated store.
template <bool AP, bool AK, int KP, int M, int M2 = 0>
__device__ __forceinline__ void trailing_pass_staged(double *__restrict__ gS, const double *tr, const double *Qs, const double *KC,
                                                     const double (&f)[2], const double (&ks)[2][3], double q2add, int lane,
                                                     const double (&ks2)[2][3] = kNoKs) {
  using TS = TrailStream<AP, KP, AK ? M : 0, 8, 1, M2, true>;
  TS ts;
  ts.tr = tr;
  ts.finish_secondary(true, gS, Qs, KC, f, ks, q2add, lane, ks2);
  auto step = [&](auto i_) { ts.template finish<decltype(i_)::value>(true, gS, Qs, KC, f, ks, q2add, lane, ks2); };
  step(std::integral_constant<int, 0>{}); step(std::integral_constant<int, 1>{}); step(std::integral_constant<int, 2>{});
  step(std::integral_constant<int, 3>{}); step(std::integral_constant<int, 4>{}); step(std::integral_constant<int, 5>{});
  step(std::integral_constant<int, 6>{}); step(std::integral_constant<int, 7>{});
  static_assert(TS::NB <= 8, "more batches");
}

// -1/tau of stored component s (0 when the process model leaves it alone), reference src/PoseUKF.cpp:37-78
__device__ inline double markov_rate(int s, const Params &P) {
  if (s < P_BG) return 0.0;
  if (s < P_BA) return -1.0 / P.gyro_bias_tau;
  if (s < P_G) return -1.0 / P.acc_bias_tau;
  if (s < P_INERTIA) return 0.0;
  if (s < P_LIN) return -1.0 / P.inertia_tau;
  if (s < P_QUAD) return -1.0 / P.lin_damping_tau;
  if (s < P_WV) return -1.0 / P.quad_damping_tau;
  if (s < P_BADCP) return -1.0 / P.wv_tau;
  if (s < P_RHO) return -1.0 / P.adcp_bias_tau;
  return -1.0 / P.rho_tau;
}

// ---------------------------------------------------------------------------------------------------------
// Latitude of the predict's sigma points (reference src/PoseUKF.cpp:29-30), ONE THREAD PER FILTER.  A sigma point's position differs
// from the mean's only for the points of columns 0 and 1 of the Cholesky factor (column j moves the coordinates >= j), so a filter
// has five distinct positions: (x, y), (x +- L00, y +- L10), (x, y +- L11).  Inside the fused kernel every lane evaluates the
// projection of its own point -- a warp instruction per operation for ONE filter, ~290 of the kernel's 5,050 instructions; here a
// warp instruction serves 32 filters and evaluates all five.  The three factor entries are computed with the very operations of
// chol_panel (same seed, same Newton step, same FMA), so the positions are the doubles the fused kernel would form itself.
// out[f][2 p], out[f][2 p + 1] = sin, cos of the latitude of sigma point p = 0..4; the fused kernel stages them with its inputs.
__global__ void __launch_bounds__(256) pose_tmerc_prepass_kernel(const double *__restrict__ state, double *__restrict__ out, long long B,
                                                                 const __grid_constant__ Params P) {
  const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= B) return;
  const double *blk = state + f * BLK;
  const double2 xy = *reinterpret_cast<const double2 *>(blk + P_POS);
  const double2 s0 = *reinterpret_cast<const double2 *>(blk + OFF_S);            // Sigma00, Sigma10
  const double s11 = blk[OFF_S + cidx(1, 1)];
  const double inv0 = fast_rsqrt3(s0.x);
  const double l00 = s0.x * inv0, l10 = s0.y * inv0;
  const double d1 = fma(-l10, l10, s11);
  const double l11 = d1 * fast_rsqrt3(d1);
  double2 *o = reinterpret_cast<double2 *>(out + f * 10);
#pragma unroll 1
  for (int p = 0; p < 5; ++p) {
    const double sg = (p & 1) ? 1.0 : -1.0;
    const double d0 = (p == 1 || p == 2) ? sg * l00 : 0.0;
    const double dy = (p == 1 || p == 2) ? sg * l10 : (p >= 3 ? sg * l11 : 0.0);
    double sl, cl;
    tmerc_nav_to_sincos_lat(P.tm, xy.x + d0, xy.y + dy, sl, cl);
    o[p] = make_double2(sl, cl);
  }
}

template <int HAS_PREDICT, int KIND, int WARPS, int KIND2 = M_NONE>
__global__ void __launch_bounds__(WARPS * 32, 1) pose_fused_kernel(Launch a, const __grid_constant__ Params P) {
  using C = FusedCfg<HAS_PREDICT, KIND, KIND2>;
  constexpr int KCW = C::KCW;
  constexpr int KP = C::KP, KT = C::KT, KE = C::KE;
  static_assert(!C::TWO || !C::SHRINK, "the two-buffer variant transforms the whole panel");
  extern __shared__ __align__(16) double smem[];
  double *Qs = smem + C::C_Q;                        // dt^2 Q, packed (only when HAS_PREDICT)
  double *ftab = smem + C::C_FTAB, *mtau = smem + C::C_MTAU, *otab = smem + C::C_OTAB;
  const int warp = threadIdx.x >> 5;
  int lane_c = threadIdx.x & 31;
#if UWVKF_LANE_PINNED
  // opaque ONCE, before the filter loop: the compiler may still hoist everything derived from the lane index out of the loop,
  // but it can no longer re-materialise the index itself from SR_TID (40 S2R + LOP3 pairs in the loop body otherwise)
  asm volatile("" : "+r"(lane_c));
#endif
  int lane = lane_c;
  double *w = smem + C::CTA + warp * C::TOTAL;
  double *mu = w + C::O_MU, *S = w + C::O_S, *Lw = w + C::O_L, *G = w + C::O_G;
  double *const muA = w + C::O_A, *const SA = muA + C::O_S;   // incoming [mu | panel] (== mu, S without a predict)
  (void)SA;
  unsigned par = 0;   // which of the two input sets belongs to the current filter
  double *KC = w + C::O_KC, *LTs = w + C::O_LT;
  (void)LTs;
  {
    const double dt_ = HAS_PREDICT ? a.ops[0].dt : 0.0;
    if (threadIdx.x < 56) {
      const int t = threadIdx.x;
      ftab[t] = (HAS_PREDICT && t < N) ? affine_factor(t, P, dt_) : 1.0;
      mtau[t] = t < DIM ? markov_rate(t, P) : 0.0;
      otab[t] = (t >= P_BG && t < P_BA) ? P.gyro_bias_offset[t - P_BG] : ((t >= P_BA && t < P_G) ? P.acc_bias_offset[t - P_BA] : 0.0);
    }
    if (HAS_PREDICT) {
      for (int e = threadIdx.x; e < NP; e += blockDim.x) Qs[e] = (dt_ * dt_) * a.Q[e];
    }
    __syncthreads();
    // per-warp column table {K_b (zero until an update fills it), f_b}
    for (int b = lane; b < 54; b += 32) {
#pragma unroll
      for (int k = 0; k < KCW; ++k) KC[KCW * b + k] = (k == 3) ? ftab[b] : 0.0;
    }
    __syncwarp();
  }
  const Op &op0 = a.ops[0];
  const Op &opu = a.ops[HAS_PREDICT ? 1 : 0];
  const Op &opu2 = a.ops[HAS_PREDICT ? 2 : 1];   // second update (two-update kernels only)
  (void)opu2;
  const double dt = HAS_PREDICT ? op0.dt : 0.0;
  const double dt2 = dt * dt;
  const int x0 = lane, x1 = lane + 32;
  const int colane = coloff(lane);
  (void)colane;
  const double frow[2] = {ftab[x0], ftab[x1 < N ? x1 : N]};   // affine process factor of the rows this lane owns (1 without predict)

  // [mu 54 | pad 2 | panel] is a prefix of the filter block.  It is copied global -> shared asynchronously (cp.async, 16 B per
  // lane and step, no registers) ONE FILTER AHEAD into the incoming buffer: kernels with a predict issue the copies for the
  // warp's next filter as soon as the predict has consumed the buffer, the others after the panel has been written back; they
  // are waited for at the top of the next iteration.
  constexpr int NV = (56 + C::PANEL_E) / 2;
  const unsigned w_saddr = (unsigned)__cvta_generic_to_shared(w);
#if UWVKF_TMA_PREFIX
  // The 6-7 KB prefix is one contiguous, 16-byte aligned range in global and in shared memory: ONE bulk copy of the TMA engine
  // (cp.async.bulk, SASS UBLKCP) issued by lane 0 brings it in and signals the warp's mbarrier with the byte count, instead of
  // a dozen 16-byte cp.async per lane.  The warp waits on the barrier's phase at the top of the next filter.
  const unsigned mb_saddr = w_saddr + 8u * C::O_MB;
  unsigned mb_phase = 0, tr_phase = 0;
  (void)tr_phase;
  if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb_saddr) : "memory");
  if (C::TRS && lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb_saddr + 8u) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
#endif
  auto fetch_prefix = [&](long long ff, unsigned pr) {
    const char *src = reinterpret_cast<const char *>(a.state + ff * BLK);
    const unsigned dsta = w_saddr + 8u * C::O_A;
#if UWVKF_TMA_PREFIX
    // the buffer was last touched through the generic proxy (write-back loads, panel stores): order those before the engine's writes
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb_saddr), "r"((unsigned)(NV * 16)) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dsta), "l"(src),
                   "r"((unsigned)(NV * 16)), "r"(mb_saddr) : "memory");
    }
#else
#pragma unroll 4
    for (int k = lane; k < NV; k += 32)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dsta + 16u * k), "l"(src + 16 * k) : "memory");
#endif
    // the filter's small inputs ride in a cp.async group, so no global-memory round trip is left inside the iteration
    const unsigned ina = w_saddr + 8u * (C::O_IN + (unsigned)C::INW * pr);
    if (HAS_PREDICT && lane < 3)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(ina + 8u * lane), "l"(a.rot_rate + ff * 3 + lane) : "memory");
    if (KIND != M_NONE && lane >= 4 && lane < 4 + C::MF::M)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(ina + 8u * lane), "l"(opu.z + ff * C::MF::M + (lane - 4)) : "memory");
    if (KIND2 != M_NONE && lane >= 8 && lane < 8 + C::MF2::M)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(ina + 8u * lane), "l"(opu2.z + ff * C::MF2::M + (lane - 8)) : "memory");
    if (C::TMP && lane >= 16 && lane < 21)   // (sin, cos) of the latitude of sigma point lane - 16, from the pre-pass kernel
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(ina + 8u * (C::INB + 2 * (lane - 16))),
                   "l"(a.tm + ff * 10 + 2 * (lane - 16)) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  {
    const long long f0 = (long long)blockIdx.x * WARPS + warp;
    if (f0 < a.B) fetch_prefix(f0, 0u);
  }

  for (long long f = (long long)blockIdx.x * WARPS + warp; f < a.B; f += (long long)gridDim.x * WARPS) {
    double *blk = a.state + f * BLK;
    const double *const inb = w + C::O_IN + C::INW * par;   // this filter's inputs
#if UWVKF_OPAQUE_LANE
    // The lane index is made opaque once per filter: everything derived from it inside the loop (row predicates, packed column
    // offsets, sigma-point signs -- well over a hundred values) would otherwise be proven loop invariant, hoisted out of the
    // filter loop and, with 168 registers per thread, spilled to local memory, which this kernel serves at L2 latency.
    // Recomputing them costs integer instructions on a pipe that is 11 % busy.
    lane = lane_c;
    asm volatile("" : "+r"(lane));
#endif
    PH_DECL
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#if UWVKF_TMA_PREFIX
    {   // the bulk copy of this filter's prefix has completed its transaction bytes when the barrier's phase flips
      unsigned done;
      do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(mb_saddr), "r"(mb_phase) : "memory");
      } while (!done);
      mb_phase ^= 1u;
    }
#endif
    __syncwarp();
    // hide HBM latency behind the panel work: pull this filter's trailing triangle (read at the very end) and the
    // prefix of the warp's next filter into L2 now; also fetch the small per-filter inputs early
    bool tr_waited = !C::TRS;
    auto tr_wait = [&]() {   // the staged triangle has landed (once per filter, on whichever path gets there first)
#if UWVKF_TMA_PREFIX
      if constexpr (C::TRS) {
        if (!tr_waited) {
          unsigned done;
          do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(mb_saddr + 8u), "r"(tr_phase) : "memory");
          } while (!done);
          tr_phase ^= 1u;
          tr_waited = true;
        }
      }
#endif
    };
    {
      const char *tr = reinterpret_cast<const char *>(blk + OFF_S + C::PANEL);   // (columns KT..KP-1 arrive with the prefix)
      constexpr int TB = (NP - C::PANEL) * 8;
#if UWVKF_TMA_PREFIX
      if constexpr (C::TRS) {
        // this filter's trailing triangle -> staging buffer, one bulk copy; it is consumed by the interleaved pass, a whole
        // predict (or, in the update-only kernels, a factorisation and a set of sigma points) later
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          constexpr unsigned bytes = (unsigned)C::TR_LEN * 8u;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb_saddr + 8u), "r"(bytes) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(w_saddr + 8u * C::O_TR),
                       "l"(reinterpret_cast<const char *>(blk + OFF_S + C::TR_START)), "r"(bytes), "r"(mb_saddr + 8u) : "memory");
        }
      } else
#endif
#if UWVKF_TRAIL_PREFETCH_L1
      for (int o = lane * 128; o < TB; o += 32 * 128) prefetch_l1(tr + o);
#else
      for (int o = lane * 128; o < TB; o += 32 * 128) prefetch_l2(tr + o);
#endif
      const long long fn = f + (long long)gridDim.x * WARPS;
      if (fn < a.B) {
        const char *nx = reinterpret_cast<const char *>(a.state + fn * BLK);
        constexpr int PB = (56 + C::PANEL_E) * 8;
        for (int o = lane * 128; o < PB; o += 32 * 128) prefetch_l2(nx + o);
      }
    }
    double off0 = 0.0, off1 = 0.0;
    if (HAS_PREDICT) {
      // Markov offsets of the stored components this lane updates (s0 = lane, s1 = lane + 32)
      const double *off = a.offsets + f * NOFF;
      const int s0 = lane, s1 = lane + 32;
      if (s0 >= P_INERTIA) off0 = off[s0 - P_INERTIA];
      if (s1 < P_WV) off1 = off[s1 - P_INERTIA];
      else if (s1 == P_RHO) off1 = off[27];
    }
    __syncwarp();
    unsigned st = 0;
    bool predicted = false, updated = false, updated2 = false, trailing_done = false;
    (void)trailing_done; (void)updated2;
    double add = 0.0;
    double ksr[2][3] = {{0, 0, 0}, {0, 0, 0}};   // (K S) rows of the accepted update for the rows this lane owns
    double ksr2[2][3] = {{0, 0, 0}, {0, 0, 0}};  // the same for the second update (two-update kernels)
    (void)ksr2;

    // =================================== predict (k = 15) ====================================================
    if constexpr (HAS_PREDICT) {
      constexpr int K = 15;
      const double *wm_g = inb;
      PH(0);
      double a0[K], a1[K];
      // a NaN rotation rate (device-resident control input the host could not scan) would poison the state: the prediction
      // is refused for this filter (status bit), like checkMeasurment() throwing in the reference (src/PoseUKF.cpp:494)
      const bool wbad = isnan(wm_g[0]) || isnan(wm_g[1]) || isnan(wm_g[2]);
      // Latitude of the sigma points (reference src/PoseUKF.cpp:29-30).  Only the points of columns 0 and 1 of the factor move the
      // position, so every lane knows its point as soon as those two columns exist.  UWVKF_TMERC_EARLY evaluates the projection
      // -- one long dependent chain per lane -- there, in stages tied to the columns of the factorisation, hoping to hide one
      // chain under the other.  Measured: the factorisation grows by exactly what the sigma-point phase loses (3,667 + 3,349 ->
      // 4,917 + 2,377 cycles per warp) and the kernel is 2 % slower: with three warps per scheduler the slots a waiting warp
      // leaves are already taken by its neighbours, only removing instructions helps.  Off.
      double tm_sl = 0.0, tm_cl = 1.0;
      bool tm_ok = false;
#if UWVKF_TMERC_EARLY
      struct TmHook {
        TMercStages &t; const TMercConst &tm; const double *Lw, *muA; int lane;
        double px, py;
        // `tie` makes a stage wait for its column (a true data dependency on the column's 1/sqrt(pivot), times zero): left alone,
        // ptxas issues the whole projection -- the longest chain in sight -- in one run as soon as column 1 exists, and the
        // factorisation stands still meanwhile
        static __device__ __forceinline__ void tie(double &v, double inv) { v = fma(inv, 0.0, v); }
        __device__ __forceinline__ void col(int j, double inv) {
          switch (j) {
            case 1: {
              const int p = lane, jj = (p - 1) >> 1;
              const bool on = p > 0 && p < 2 * K + 1;
              const double sg = (p & 1) ? 1.0 : -1.0;
              const double d0 = (on && 0 >= jj) ? sg * Lw[jj * LStride<K> + 0] : 0.0;
              const double d1 = (on && 1 >= jj) ? sg * Lw[jj * LStride<K> + 1] : 0.0;
              px = muA[P_POS] + d0; py = muA[P_POS + 1] + d1;
              t.step<0>(tm, px, py);
            } break;
            case 2: tie(t.s1, inv); t.step<1>(tm, px, py); break;
            case 3: tie(t.sj, inv); t.step<2>(tm, px, py); break;
            case 4: tie(t.sj, inv); t.step<3>(tm, px, py); break;
            case 5: tie(t.sj, inv); t.step<4>(tm, px, py); break;
            case 6: tie(t.chp, inv); t.step<5>(tm, px, py); break;
            case 7: tie(t.sch, inv); t.step<6>(tm, px, py); break;
            case 8: tie(t.sk, inv); t.step<7>(tm, px, py); break;
            case 9: tie(t.sk, inv); t.step<8>(tm, px, py); break;
            case 10: tie(t.sk, inv); t.step<9>(tm, px, py); break;
            case 11: tie(t.eps, inv); t.step<10>(tm, px, py); break;
            default: break;
          }
        }
      };
      TMercStages tms;
      tms.ok = false; tms.sin_lat = 0.0; tms.cos_lat = 1.0;
      const bool spd15 = !wbad && chol_panel<K>(SA, Lw, lane, a0, a1, TmHook{tms, P.tm, Lw, muA, lane, 0.0, 0.0});
      tm_sl = tms.sin_lat; tm_cl = tms.cos_lat; tm_ok = tms.ok;
#else
      const bool spd15 = !wbad && chol_panel<K>(SA, Lw, lane, a0, a1);
#endif
      PH(1);
      if (spd15) {
        const Quat q{muA[P_ORI], muA[P_ORI + 1], muA[P_ORI + 2], muA[P_ORI + 3]};
        const Vec3 vel{muA[P_VEL], muA[P_VEL + 1], muA[P_VEL + 2]};
        const int p = lane;
        const bool on = p > 0 && p < 2 * K + 1;
        const int j = (p - 1) >> 1;
        const double sg = (p & 1) ? 1.0 : -1.0;
        auto d = [&](int t) { return (on && t >= j) ? sg * Lw[j * LStride<K> + t] : 0.0; };
        Quat qp;
        {
          double px = muA[P_POS] + d(0), py = muA[P_POS + 1] + d(1);
          Quat qq = so3_boxplus(q, Vec3{d(3), d(4), d(5)}, 1.0);
          Vec3 wb{wm_g[0] - (muA[P_BG] + d(12)), wm_g[1] - (muA[P_BG + 1] + d(13)), wm_g[2] - (muA[P_BG + 2] + d(14))};
          double sl = tm_sl, cl = tm_cl;
#if UWVKF_TMERC_PREPASS
          {   // only the points of columns 0 and 1 of the factor (p = 1..4) sit at a position of their own
            const double2 t2 = *reinterpret_cast<const double2 *>(inb + C::INB + 2 * (p < 5 ? p : 0));
            sl = t2.x; cl = t2.y;
            (void)px; (void)py; (void)tm_ok;
          }
#elif UWVKF_TMERC_EARLY
          if (!tm_ok) tmerc_nav_to_sincos_lat(P.tm, px, py, sl, cl);   // far from the projection origin: generic routines
#elif defined(UWVKF_EXP_NOTMERC)
          sl = P.tm.sin_xi0 + 1e-7 * px; cl = P.tm.cos_xi0 + 1e-7 * py;   // experiment (WRONG results): what the projection costs
#else
          tmerc_nav_to_sincos_lat(P.tm, px, py, sl, cl);
#endif
          Vec3 wn = qrot(qq, wb);
          wn.x -= kEarthW * cl;
          wn.z -= kEarthW * sl;
          qp = so3_boxplus(qq, wn, dt);
        }
        PH(2);
        double soo[6], T[2][3];
        add = P.wv_scale * (vel.x * vel.x + vel.y * vel.y + 100.0 * vel.z * vel.z) * dt;
        // predict-only kernel: the trailing pass (process map only) needs nothing the predict computes, so its batches are
        // interleaved with the sigma-point / mean / panel phases below, like the update kernels do inside apply_delta
        constexpr bool PILV = UWVKF_ILV && UWVKF_ILV_PREDICT && KIND == M_NONE;
        TrailStream<true, KP, 0, UWVKF_ILV_UB, UWVKF_ILV_DEPTH> pts;
        double *const pgS = blk + OFF_S;
        const double pq2 = dt2 * add;
        auto pts_step = [&](auto i_) {
          constexpr int I = decltype(i_)::value;
          pts.template finish<I>(true, pgS, Qs, KC, frow, ksr, pq2, lane);
          pts.template load<I + UWVKF_ILV_DEPTH>(pgS, lane);
        };
        if constexpr (PILV) {
          pts.template load<0>(pgS, lane);
          if (UWVKF_ILV_DEPTH > 1) pts.template load<1>(pgS, lane);
          pts.load_secondary(pgS, lane);
        }
        Quat r;
        if constexpr (PILV) {
          r = so3_mean_lane(qp, ori_weight(p, K), P.mean_tol, P.mean_max_it, st);
          pts.finish_secondary(true, pgS, Qs, KC, frow, ksr, pq2, lane);
          pts_step(std::integral_constant<int, 0>{});
          ori_deviations<K>(G, qp, r, p, soo);
          pts_step(std::integral_constant<int, 1>{});
          a_times_panel<K, 3>(G, a0, a1, T);
          pts_step(std::integral_constant<int, 2>{});
        } else {
          r = ori_transform_reg<K>(G, a0, a1, qp, p, P, soo, T, lane, st PH_ARG(3));
        }
        // effective process noise pieces (reference src/PoseUKF.cpp:448-460)
        double qoo[6];
        {
          double Rm[9], Qo[9], t[9];
          qmat(q, Rm);
          // Qs holds dt^2 Q, so qoo comes out scaled as well
          Qo[0] = Qs[cidx(3, 3)]; Qo[1] = Qs[cidx(4, 3)]; Qo[2] = Qs[cidx(5, 3)];
          Qo[3] = Qo[1];          Qo[4] = Qs[cidx(4, 4)]; Qo[5] = Qs[cidx(5, 4)];
          Qo[6] = Qo[2];          Qo[7] = Qo[5];          Qo[8] = Qs[cidx(5, 5)];
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int b = 0; b < 3; ++b) t[i * 3 + b] = Rm[i * 3] * Qo[b] + Rm[i * 3 + 1] * Qo[3 + b] + Rm[i * 3 + 2] * Qo[6 + b];
          auto rq = [&](int i, int b) { return t[i * 3] * Rm[b * 3] + t[i * 3 + 1] * Rm[b * 3 + 1] + t[i * 3 + 2] * Rm[b * 3 + 2]; };
          qoo[0] = rq(0, 0); qoo[1] = rq(1, 0); qoo[2] = rq(2, 0); qoo[3] = rq(1, 1); qoo[4] = rq(2, 1); qoo[5] = rq(2, 2);
        }
        const double (&rdf)[2] = frow;
        // new mean of the affine components: x + dt * rate, rate = v | a | (-1/tau)(x - offset)   (src/PoseUKF.cpp:25-78)
        // (branch-free: the three cases sit on different lanes of the same warp)
        auto new_mean = [&](int sidx_, double offv) {
          const int sc = sidx_ < DIM ? sidx_ : 0;
          const double xv = muA[sc];
          const bool kin = sc < P_ORI || (sc >= P_VEL && sc < P_ACC);          // position <- velocity, velocity <- acceleration
          const double partner = muA[kin ? sc + (sc < P_ORI ? P_VEL : 3) : sc];
          const double o = ((sc >= P_INERTIA && sc < P_WV) || sc == P_RHO) ? offv : otab[sc];
          const double rate = kin ? partner : mtau[sc] * (xv - o);
          return fma(dt, rate, xv);
        };
        const double nm0 = new_mean(x0, off0), nm1 = new_mean(x1, off1);   // consumed only where the row is not the orientation
        __syncwarp();
        PH(7);
        // panel <- F panel F^T + dt^2 Q on the affine rows (couplings and scaling in one pass)
        PH(8);
#if !UWVKF_EXP_SKIP_PANEL   // experiment switch (WRONG results; what the phase costs)
        if constexpr (C::TWO) process_panel_out<KP>(SA, S, Qs, ftab, rdf, dt, lane, colane);
        else process_panel<C::KT>(S, Qs, ftab, rdf, dt, lane, colane);
#endif
        PH(9);
        if constexpr (PILV) pts_step(std::integral_constant<int, 3>{});
        // orientation rows/cols: T F^T (+ Q cross terms)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int x = lane + 32 * h;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            // coupled partner: T[c][x+6] for x in pos, T[c][x+3] for x in vel (both in round 0)
            double t6 = __shfl_sync(kFull, T[0][c], (lane + 6) & 31), t3 = __shfl_sync(kFull, T[0][c], (lane + 3) & 31);
            if (x < N && !(x >= 3 && x < 6)) {
              double tp = rdf[h] * T[h][c];
              if (h == 0 && x < 3) tp += dt * t6;
              else if (h == 0 && x >= T_VEL && x < T_ACC) tp += dt * t3;
              const int e = x < 3 ? (coloff(x) + 3 + c) : (coloff(3 + c) + x);
              S[e] = tp + Qs[e];
            }
          }
        }
        if (lane == 0) {   // soo / qoo are warp-uniform: one lane stores the six entries (order: (3,3) (4,3) (5,3) (4,4) (5,4) (5,5))
          S[cidx(3, 3)] = soo[0] + qoo[0]; S[cidx(4, 3)] = soo[1] + qoo[1]; S[cidx(5, 3)] = soo[2] + qoo[2];
          S[cidx(4, 4)] = soo[3] + qoo[3]; S[cidx(5, 4)] = soo[4] + qoo[4]; S[cidx(5, 5)] = soo[5] + qoo[5];
        }
        if (x0 < P_ORI || x0 >= P_VEL) mu[x0] = nm0;
        if (x1 < DIM) mu[x1] = nm1;
        if (lane == 0) { mu[P_ORI] = r.x; mu[P_ORI + 1] = r.y; mu[P_ORI + 2] = r.z; mu[P_ORI + 3] = r.w; }
        predicted = true;
        if constexpr (PILV) {
          pts_step(std::integral_constant<int, 4>{}); pts_step(std::integral_constant<int, 5>{});
          pts_step(std::integral_constant<int, 6>{}); pts_step(std::integral_constant<int, 7>{});
          static_assert(decltype(pts)::NB <= 8, "drain more batches");
          trailing_done = true;
        }
        __syncwarp();
        PH(10);
      } else {
        st |= wbad ? ST_NAN_MEAS : ST_NOT_SPD;
        if constexpr (C::TWO) {   // refused: the working buffer gets the incoming state as it is
          const double2 *src = reinterpret_cast<const double2 *>(muA);
          double2 *dst = reinterpret_cast<double2 *>(mu);
          for (int k = lane; k < NV; k += 32) dst[k] = src[k];
          __syncwarp();
        }
      }
      if constexpr (C::TWO) {   // the incoming buffer is free: the next filter's prefix lands during the update
        __syncwarp();
        const long long fn = f + (long long)gridDim.x * WARPS;
        if (fn < a.B) fetch_prefix(fn, par ^ 1u);
      }
    }

    // =================================== update (k = Meas::K) + apply_delta (k = 6) ==========================
    // One update as a generic lambda: a two-update kernel runs it twice on the panel in shared memory.  SECOND = the second
    // update of such a kernel; DEFER = its first, whose share of the trailing triangle is left to the pass of the second.
    auto do_update = [&](auto kind_, auto second_, const Op &opu, const double *zin) {
      constexpr int KINDX = decltype(kind_)::value;
      constexpr bool SECOND = decltype(second_)::value;
      constexpr bool DEFER = (KIND2 != M_NONE) && !SECOND;
      constexpr int KOFF = SECOND ? 4 : 0;      // slots of this update's gain in the column table
      double (&ksx)[2][3] = *(SECOND ? &ksr2 : &ksr);
      bool accepted = false;
      using MM = FMeas<KINDX>;
      constexpr int K = MM::K, M = MM::M, NT = MM::NT;
      constexpr int NPTS = 2 * K + 1, ROUNDS = (NPTS + 31) / 32;
      static_assert(NT == 0 || (MM::tail(0) >= K && K <= 32), "tail coordinates lie behind the prefix");
      static_assert(!SECOND || NT == 0 || MM::tail(NT - 1) < KP, "the second update must find its tail coordinates in the panel");
      // entry (ta, tb), ta >= tb >= KP, of the covariance the update works on: the trailing triangle still holds the
      // pre-predict values in HBM, so the process map is applied on the fly (same expression as the write-back pass)
      auto trailing = [&](int ta, int tb) {
        // (a second update never gets here: its tail coordinates must lie inside the panel, checked below)
        const int e = coloff(tb) + ta;
        double sv = (blk + OFF_S)[e];
        if (HAS_PREDICT && predicted) {
          sv = fma(ftab[ta] * ftab[tb], sv, Qs[HAS_PREDICT ? e : 0]);
          if (ta == tb && tb >= T_WV && tb < T_BADCP) sv += dt2 * add;
        }
        return sv;
      };
      // entry (ta, tb), ta >= tb, wherever it lives: the panel in shared memory (column tb < KP) or the trailing triangle
      auto sig = [&](int ta, int tb) { return tb < C::KP ? S[coloff(tb) + ta] : trailing(ta, tb); };
      (void)sig;
      bool go = !(opu.valid && !opu.valid[f]);
      double z[M], Rm[M * M];
      double u0[K], u1[K];   // rows x0, x1 of the panel factor (chol_panel)
      if (go) {
        bool bad = false;
#pragma unroll
        for (int k = 0; k < M; ++k) { z[k] = zin[k]; bad |= isnan(z[k]); }
#pragma unroll
        for (int k = 0; k < M * M; ++k) { Rm[k] = opu.Rdev ? opu.Rdev[f * M * M + k] : opu.R[k]; bad |= isnan(Rm[k]); }
        if (bad) { st |= ST_NAN_MEAS; go = false; }
      }
      if (go) {
        if constexpr (KINDX == M_XY) {
          if (opu.kind == M_GEOGRAPHIC) {   // reference src/PoseUKF.cpp:572-578
            double px, py;
            tmerc_world_to_nav(P.tm, z[0], z[1], px, py);
            Vec3 lever = qrot(Quat{mu[3], mu[4], mu[5], mu[6]}, Vec3{opu.aux[0], opu.aux[1], opu.aux[2]});
            z[0] = px - lever.x; z[1] = py - lever.y;
          } else if (opu.kind == M_DELAYED_XY) {   // reference src/PoseUKF.cpp:516-521
            z[0] = z[0] + (mu[0] - opu.aux[0]); z[1] = z[1] + (mu[1] - opu.aux[1]);
          }
        }
        PH(11);
        if (!chol_panel<K>(S, Lw, lane, u0, u1)) { st |= ST_NOT_SPD; go = false; }
        PH(12);
      }
      if (go) {
        double lt[NT > 0 ? NT : 1];   // lane j < K: L[T0 + r][j], rows T of the panel factor Sigma[:, :K] L11^-T
        if constexpr (NT > 0) {
          // rows T of the panel factor sit in the registers of the lanes that own those rows: transpose through shared memory
#pragma unroll
          for (int r = 0; r < NT; ++r) {
            constexpr int dummy = 0; (void)dummy;
            const int t = MM::tail(r);
            if (lane == (t & 31)) {
#pragma unroll
              for (int j = 0; j < K; ++j) LTs[r * KE + j] = (t >= 32) ? u1[j] : u0[j];
            }
          }
          __syncwarp();
#pragma unroll
          for (int r = 0; r < NT; ++r) lt[r] = (lane < K) ? LTs[r * KE + lane] : 0.0;
        }
        double Z[ROUNDS][M], wt[ROUNDS];
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
          const int p = lane + 32 * r;
          const bool act = p < NPTS;
          const bool on = act && p > 0;
          const int j = (p - 1) >> 1;
          const double sg = (p & 1) ? 1.0 : -1.0;
          auto d = [&](int t) {
            if (NT > 0 && t >= K) return on ? sg * LTs[MM::slot(t) * KE + j] : 0.0;   // t is a literal: resolved at compile time
            return (on && t >= j) ? sg * Lw[j * LStride<K> + (t < K ? t : 0)] : 0.0;
          };
          MM::eval(mu, d, opu.aux, P, Z[r]);
          wt[r] = (p == 0) ? double(1 + 2 * (N - K)) : (act ? 1.0 : 0.0);
        }
        PH(13);
        double zb[M];
#pragma unroll
        for (int k = 0; k < M; ++k) zb[k] = bcast(Z[0][k], 0);
        {
          int it = 0;
          double nrm;
          do {
            nrm = 0;
#pragma unroll
            for (int k = 0; k < M; ++k) {
              double m = 0;
#pragma unroll
              for (int r = 0; r < ROUNDS; ++r) m = fma(wt[r], Z[r][k] - zb[k], m);
              m = wsum(m) * (1.0 / double(2 * N + 1));
              zb[k] += m;
              nrm += m * m;
            }
          } while (nrm > P.mean_tol * P.mean_tol && ++it < P.mean_max_it);   // nrm holds the squared norm
          if (it >= P.mean_max_it) st |= ST_MEAN_NOT_CONVERGED;
        }
        PH(14);
        double Szz[M * M];
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
          for (int b = 0; b <= i; ++b) {
            double s = 0;
#pragma unroll
            for (int r = 0; r < ROUNDS; ++r) s = fma(0.5 * wt[r] * (Z[r][i] - zb[i]), Z[r][b] - zb[b], s);
            s = wsum(s);
            Szz[i * M + b] = s + Rm[i * M + b];
            Szz[b * M + i] = s + Rm[b * M + i];
          }
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
          const int p = lane + 32 * r;
#pragma unroll
          for (int k = 0; k < M; ++k) {
            double src = Z[r][k];
            if (r + 1 < ROUNDS) { if (lane == 0) src = Z[r + 1 < ROUNDS ? r + 1 : r][k]; }
            double nxt = __shfl_sync(kFull, src, (lane + 1) & 31);
            if ((p & 1) && p < NPTS) G[4 * ((p - 1) >> 1) + k] = 0.5 * (Z[r][k] - nxt);
          }
        }
        __syncwarp();
        double Ct[NT > 0 ? M * NT : 1];
        if constexpr (NT > 0) {
          MM::tail_jac(mu, opu.aux, P, Ct);
          // U = C L_T (lane j holds column j); S_zz += C Sigma_TT C^T - U U^T; G <- B - U
          double U[M];
#pragma unroll
          for (int c = 0; c < M; ++c) {
            U[c] = 0.0;
#pragma unroll
            for (int r = 0; r < NT; ++r) U[c] = fma(Ct[c * NT + r], lt[r], U[c]);
          }
          // C Sigma_TT C^T with ONE ROW of Sigma_TT per lane (lane r < NT): w_r = Sigma_TT[r, :] C^T, then
          // (C Sigma_TT C^T)[i][b] = sum_r C[i][r] w_r[b] rides in the same warp sum as -U U^T.  (Every lane forming all of
          // Sigma_TT and the full triple product costs NT^2 loads and 2 NT^2 FMAs per entry of S_zz.)
          const int rr = lane < NT ? lane : 0;
          int tr = MM::tail(0);
#pragma unroll
          for (int k = 1; k < NT; ++k) tr = (rr == k) ? MM::tail(k) : tr;
          double wrow[M], crow[M];
#pragma unroll
          for (int b = 0; b < M; ++b) wrow[b] = 0.0;
#pragma unroll
          for (int b2 = 0; b2 < NT; ++b2) {
            const int tb = MM::tail(b2);
            const double sv = (tr >= tb) ? sig(tr, tb) : sig(tb, tr);
#pragma unroll
            for (int b = 0; b < M; ++b) wrow[b] = fma(sv, Ct[b * NT + b2], wrow[b]);
          }
#pragma unroll
          for (int i = 0; i < M; ++i) {
            double c = Ct[i * NT];
#pragma unroll
            for (int k = 1; k < NT; ++k) c = (rr == k) ? Ct[i * NT + k] : c;
            crow[i] = lane < NT ? c : 0.0;
          }
#pragma unroll
          for (int i = 0; i < M; ++i)
#pragma unroll
            for (int b = 0; b <= i; ++b) {
              const double t2 = wsum(fma(crow[i], wrow[b], -(U[i] * U[b])));
              Szz[i * M + b] += t2;
              if (b != i) Szz[b * M + i] += t2;
            }
          if (lane < K) {
#pragma unroll
            for (int c = 0; c < M; ++c) G[4 * lane + c] -= U[c];
          }
          __syncwarp();
        }
        PH(15);
        PH(16);
        double T[2][M];
        a_times_panel<K, M>(G, u0, u1, T);   // T[h][c] = Pxz[x_h][c] = sum_j B[c][j] Lp[x_h][j]
        PH(17);
        if constexpr (NT > 0) {   // + C Sigma[T, x]: panel columns from shared memory, the rest from the trailing triangle
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int x = lane + 32 * h;
            if (x < N) {
#pragma unroll
              for (int r = 0; r < NT; ++r) {
                const int t = MM::tail(r);
                const double sx = x <= t ? sig(t, x) : sig(x, t);
#pragma unroll
                for (int c = 0; c < M; ++c) T[h][c] = fma(Ct[c * NT + r], sx, T[h][c]);
              }
            }
          }
        }
        double Si[M * M], nu[M];
        sym_inverse<M>(Szz, Si);
#pragma unroll
        for (int k = 0; k < M; ++k) nu[k] = z[k] - zb[k];
        double d2 = 0;
#pragma unroll
        for (int i = 0; i < M; ++i)
#pragma unroll
          for (int b = 0; b < M; ++b) d2 += nu[i] * Si[i * M + b] * nu[b];
        // d2p95: GeographicPosition (reference src/PoseUKF.cpp:578) and the ADCP update (:610)
        const bool gate95 = ((KINDX == M_XY) && opu.kind == M_GEOGRAPHIC) || KINDX == M_WATER_VELOCITY;
        if (gate95 && d2 > 5.991) {
          st |= ST_GATE_REJECTED;
        } else {
          // gain rows, (K S) rows and delta for the rows this lane owns; K and K S go straight to the row tables the
          // trailing pass reads (and are re-read from there in the rare rollback) so that they do not stay in registers
          double dl[2], kk[2][3] = {{0, 0, 0}, {0, 0, 0}};
          double (&ks)[2][3] = ksx;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            double dx = 0;
#pragma unroll
            for (int i = 0; i < M; ++i) {
              double s = 0;
#pragma unroll
              for (int c = 0; c < M; ++c) s += T[h][c] * Si[c * M + i];
              kk[h][i] = s;
              dx += s * nu[i];
            }
#pragma unroll
            for (int i = 0; i < M; ++i) {
              double s = 0;
#pragma unroll
              for (int c = 0; c < M; ++c) s += kk[h][c] * Szz[c * M + i];
              ks[h][i] = s;
            }
            dl[h] = dx;
          }
          // column-side broadcast copy of K (the f_b slot of the table is never rewritten)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int x = lane + 32 * h;
            if (x < N) {
              *reinterpret_cast<double2 *>(KC + KCW * x + KOFF) = make_double2(kk[h][0], kk[h][1]);
              KC[KCW * x + KOFF + 2] = kk[h][2];
            }
          }
          __syncwarp();
          PH(18);
          // interleaved trailing pass (TrailStream): the first batches are requested now and fly during the panel downdate and
          // the factorisation of apply_delta
          constexpr bool ILV = UWVKF_ILV && (HAS_PREDICT || UWVKF_ILV_UPDATE_ONLY) && !DEFER;
          constexpr int M1_ = SECOND ? C::MF::M : M, M2_ = SECOND ? M : 0;   // the pass of a second update carries both downdates
          TrailStream<HAS_PREDICT != 0, KT, M1_, UWVKF_ILV_UB, UWVKF_ILV_DEPTH, M2_, C::TRS> ts;
          if constexpr (C::TRS) { tr_wait(); ts.tr = w + C::O_TR - C::TR_START; }
          // Filters whose prediction was refused (rare) still run the batches -- no branch around loads whose values are consumed
          // several phases later -- but with every store predicated off; the classic pass then handles them
          const bool ilv = ILV && (HAS_PREDICT ? predicted : true);
          double *const tgS = blk + OFF_S;
          const double tq2 = dt2 * add;
          int tl = lane;
#if UWVKF_OPAQUE_LANE || UWVKF_TRAIL_OPAQUE
          asm volatile("" : "+r"(tl));   // opaque per filter (see TrailStream)
#endif
          auto ts_step = [&](auto i_) {   // batch I out, batch I + DEPTH in
            constexpr int I = decltype(i_)::value;
            ts.template finish<I>(ilv, tgS, Qs, KC, frow, ksr, tq2, tl, ksr2);
            ts.template load<I + UWVKF_ILV_DEPTH>(tgS, tl);
          };
#if UWVKF_ILV_EARLY
          if constexpr (ILV) {
            ts.template load<0>(tgS, tl);
            if (UWVKF_ILV_DEPTH > 1) ts.template load<1>(tgS, tl);
            if (UWVKF_ILV_DEPTH > 2) ts.template load<2>(tgS, tl);
            ts.load_secondary(tgS, tl);
          }
#endif
          // panel downdate: Sigma[x][b] -= (KS)_x . K_b
#pragma unroll
          for (int b = 0; b < KT; ++b) {
            const int cb = coloff(b);
            const double2 k01 = *reinterpret_cast<const double2 *>(KC + KCW * b + KOFF);
            const double k0 = k01.x, k1 = k01.y, k2 = KC[KCW * b + KOFF + 2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int x = lane + 32 * h;
              if (x >= b && x < N) {
                double s = S[cb + x];
                s = fma(-ks[h][0], k0, s);
                if (M > 1) s = fma(-ks[h][1], k1, s);
                if (M > 2) s = fma(-ks[h][2], k2, s);
                S[cb + x] = s;
              }
            }
          }
          __syncwarp();
          PH(19);
#if !UWVKF_ILV_EARLY
          if constexpr (ILV) {   // the first batches fly during the factorisation of apply_delta
            ts.template load<0>(tgS, tl);
            if (UWVKF_ILV_DEPTH > 1) ts.template load<1>(tgS, tl);
            if (UWVKF_ILV_DEPTH > 2) ts.template load<2>(tgS, tl);
            ts.load_secondary(tgS, tl);
          }
#endif
          // ---- apply_delta(K nu): only the orientation (prefix 6) is a manifold component -----------------
          constexpr int KA = 6;
          double c0[KA], c1[KA];
#if UWVKF_EXP_SKIP_APPLY    // experiment switch (WRONG results; what apply_delta costs)
          const bool spd6 = false;
#else
          const bool spd6 = chol_panel<KA>(S, Lw, lane, c0, c1);
#endif
          PH(20);
          if (spd6) {
            const Quat q{mu[P_ORI], mu[P_ORI + 1], mu[P_ORI + 2], mu[P_ORI + 3]};
            const double d3 = bcast(dl[0], 3), d4 = bcast(dl[0], 4), d5 = bcast(dl[0], 5);
            const int p = lane;
            const bool on = p > 0 && p < 2 * KA + 1;
            const int j = (p - 1) >> 1;
            const double sg = (p & 1) ? 1.0 : -1.0;
            auto d = [&](int t) { return (on && t >= j) ? sg * Lw[j * LStride<KA> + t] : 0.0; };
            if constexpr (ILV) { ts.finish_secondary(ilv, tgS, Qs, KC, frow, ksr, tq2, tl, ksr2); ts_step(std::integral_constant<int, 0>{}); }
            Quat qp = so3_boxplus(q, Vec3{d3 + d(3), d4 + d(4), d5 + d(5)}, 1.0);
            PH(21);
            if constexpr (ILV) ts_step(std::integral_constant<int, 1>{});
            double soo[6], TA[2][3];
            Quat r = so3_mean_lane(qp, ori_weight(p, KA), P.mean_tol, P.mean_max_it, st);
            PH(22);
            if constexpr (ILV) ts_step(std::integral_constant<int, 2>{});
            ori_deviations<KA>(G, qp, r, p, soo);
            PH(23);
            if constexpr (ILV) ts_step(std::integral_constant<int, 3>{});
            PH(24);
            a_times_panel<KA, 3>(G, c0, c1, TA);
            PH(25);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int x = lane + 32 * h;
              if (x < N && !(x >= 3 && x < 6)) {
#pragma unroll
                for (int c = 0; c < 3; ++c) S[x < 3 ? (coloff(x) + 3 + c) : (coloff(3 + c) + x)] = TA[h][c];
              }
            }
            if (lane == 0) {
              S[cidx(3, 3)] = soo[0]; S[cidx(4, 3)] = soo[1]; S[cidx(5, 3)] = soo[2];
              S[cidx(4, 4)] = soo[3]; S[cidx(5, 4)] = soo[4]; S[cidx(5, 5)] = soo[5];
            }
            // mu (+)= delta on the vector components: tangent x -> stored x (x < 3) or x + 1 (x >= 6)
            if (x0 < 3) mu[x0] += dl[0];
            else if (x0 >= 6) mu[x0 + 1] += dl[0];
            if (x1 < N) mu[x1 + 1] += dl[1];
            if (lane == 0) { mu[P_ORI] = r.x; mu[P_ORI + 1] = r.y; mu[P_ORI + 2] = r.z; mu[P_ORI + 3] = r.w; }
            updated = true;
            accepted = true;
            if constexpr (SECOND) updated2 = true;
            if constexpr (ILV) {
              ts_step(std::integral_constant<int, 4>{}); ts_step(std::integral_constant<int, 5>{});
              ts_step(std::integral_constant<int, 6>{}); ts_step(std::integral_constant<int, 7>{});
              static_assert(decltype(ts)::NB <= 8, "drain more batches");
              trailing_done = ilv;
            }
            __syncwarp();
            PH(26);
          } else {
            // the downdated covariance lost definiteness in its leading block: roll the panel back, drop the update
            st |= ST_NOT_SPD;
#pragma unroll 1
            for (int b = 0; b < KT; ++b) {
              const int cb = coloff(b);
              const double k0 = KC[KCW * b + KOFF], k1 = KC[KCW * b + KOFF + 1], k2 = KC[KCW * b + KOFF + 2];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int x = lane + 32 * h;
                if (x >= b && x < N) {
                  double s = S[cb + x];
                  s = fma(ksx[h][0], k0, s);
                  if (M > 1) s = fma(ksx[h][1], k1, s);
                  if (M > 2) s = fma(ksx[h][2], k2, s);
                  S[cb + x] = s;
                }
              }
            }
            __syncwarp();
            if constexpr (KIND2 != M_NONE) {   // a dropped update contributes nothing to the (shared) trailing pass
#pragma unroll
              for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int c = 0; c < 3; ++c) ksx[h][c] = 0.0;
            }
          }
        }
      }
      if constexpr (KIND2 != M_NONE) {
        // The trailing pass of a two-update kernel always applies both downdates; an update that was refused (dropout mask, NaN,
        // gate, lost definiteness) has zero (K S) rows, and its slots of the column table are cleared as well, so that a
        // non-finite gain left there by an earlier, degenerate filter of this warp cannot leak in as 0 x NaN
        if (!accepted) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int x = lane + 32 * h;
            if (x < 54) {
              *reinterpret_cast<double2 *>(KC + KCW * x + KOFF) = make_double2(0.0, 0.0);
              KC[KCW * x + KOFF + 2] = 0.0;
            }
          }
          __syncwarp();
        }
      }
    };
    if constexpr (KIND != M_NONE) do_update(std::integral_constant<int, KIND>{}, std::false_type{}, opu, inb + 4);
    if constexpr (KIND2 != M_NONE) do_update(std::integral_constant<int, KIND2>{}, std::true_type{}, opu2, inb + 8);

    // =================================== write back ===========================================================
    PH(27);
    // navigation solution of the pipelined host step: the first ten stored doubles of mu, whatever happened to the filter
    if constexpr (HAS_PREDICT) {   // (the host step always predicts; the update-only kernels stay as they were: 4 % each, register allocation)
      if (a.nav != nullptr && lane < 10) __stcs(a.nav + f * 10 + lane, mu[lane]);
    }
    if (predicted || updated) {
      {
        double2 *dst = reinterpret_cast<double2 *>(blk);
        const double2 *src = reinterpret_cast<const double2 *>(mu);
        // if PANEL is odd the last double2 would also carry the first trailing entry, which the interleaved trailing pass
        // has already rewritten in HBM: the last panel entry goes out on its own
        constexpr int NVT = (56 + C::PANEL_T + 1) / 2;                 // [mu | the KT columns that are still panel]
        constexpr int NVW = (C::PANEL_T & 1) ? NVT - 1 : NVT;
#pragma unroll 4
        for (int k = lane; k < NVW; k += 32) __stcs(dst + k, src[k]);
        if ((C::PANEL_T & 1) && lane == 0) __stcs(blk + 2 * NVW, mu[2 * NVW]);
      }
      __syncwarp();
      if constexpr (!C::TWO) {   // the panel buffer is free now: start the next filter's prefix, it lands during the trailing pass
        const long long fn = f + (long long)gridDim.x * WARPS;
        if (fn < a.B) fetch_prefix(fn, par ^ 1u);
      }
      // trailing triangle (columns >= KP), see trailing_pass: process map and / or rank-M downdate, whichever happened
      PH(28);
      double *gS = blk + OFF_S;
      const double q2add = dt2 * add;
      constexpr int MM_ = (KIND == M_NONE) ? 1 : C::MF::M;
      constexpr int TUB = HAS_PREDICT ? UWVKF_TUB : 8, TDEPTH = HAS_PREDICT ? UWVKF_TDEPTH : 3;
      // opaque per filter: the column offsets of the pass are recomputed instead of being hoisted out of the filter loop and
      // spilled (see TrailStream)
      int tlane = lane;
#if UWVKF_OPAQUE_LANE || UWVKF_TRAIL_OPAQUE
      if (HAS_PREDICT && KIND != M_NONE) asm volatile("" : "+r"(tlane));
#endif
#ifndef UWVKF_EXP_SKIPTRAIL
#define UWVKF_EXP_SKIPTRAIL 0      // experiment: no trailing pass at all (WRONG results; upper bound of what the pass can cost)
#endif
      if (trailing_done || UWVKF_EXP_SKIPTRAIL) {
        // the interleaved pass (TrailStream) has already rewritten the trailing triangle
      } else if constexpr (C::TRS && KIND == M_NONE) {
        // predict-only kernel: this IS the hot path -- old values from the staging copy (one bulk copy per filter), nothing to
        // pipeline.  (The kernels with an update come here for refused filters only and keep the pipelined pass from HBM: giving
        // them the staged form as well costs the fused predict + DVL kernel 5 % -- register allocation again.)
        tr_wait();
        trailing_pass_staged<true, false, KT, MM_>(gS, w + C::O_TR - C::TR_START, Qs, KC, frow, ksr, q2add, tlane);
      } else if constexpr (KIND2 != M_NONE) {   // both downdates in one pass; refused updates carry zero factors
        if (HAS_PREDICT && predicted) trailing_pass<true, true, KT, MM_, TUB, TDEPTH, C::MF2::M>(gS, Qs, KC, frow, ksr, q2add, tlane, ksr2);
        else trailing_pass<false, true, KT, MM_, TUB, TDEPTH, C::MF2::M>(gS, Qs, KC, frow, ksr, q2add, tlane, ksr2);
      } else if constexpr (KIND == M_NONE) {
        trailing_pass<true, false, KT, MM_, TUB, TDEPTH>(gS, Qs, KC, frow, ksr, q2add, tlane);
      } else if constexpr (!HAS_PREDICT) {
        trailing_pass<false, true, KT, MM_, TUB, TDEPTH>(gS, Qs, KC, frow, ksr, q2add, tlane);
      } else if constexpr (UWVKF_ILV && UWVKF_COLD_CALL) {   // rare: the prediction or the update of this filter was refused
        if (predicted && updated)
          trailing_pass_cold<true, true, KT, MM_, TUB, TDEPTH>(gS, Qs, KC, frow[0], frow[1], ksr[0][0], ksr[0][1], ksr[0][2], ksr[1][0], ksr[1][1], ksr[1][2], q2add, tlane);
        else if (predicted)
          trailing_pass_cold<true, false, KT, MM_, TUB, TDEPTH>(gS, Qs, KC, frow[0], frow[1], ksr[0][0], ksr[0][1], ksr[0][2], ksr[1][0], ksr[1][1], ksr[1][2], q2add, tlane);
        else
          trailing_pass_cold<false, true, KT, MM_, TUB, TDEPTH>(gS, Qs, KC, frow[0], frow[1], ksr[0][0], ksr[0][1], ksr[0][2], ksr[1][0], ksr[1][1], ksr[1][2], q2add, tlane);
      } else {
        if (predicted && updated) trailing_pass<true, true, KT, MM_, TUB, TDEPTH>(gS, Qs, KC, frow, ksr, q2add, tlane);
        else if (predicted) trailing_pass<true, false, KT, MM_, TUB, TDEPTH>(gS, Qs, KC, frow, ksr, q2add, tlane);
        else trailing_pass<false, true, KT, MM_, TUB, TDEPTH>(gS, Qs, KC, frow, ksr, q2add, tlane);
      }
    }
    else if constexpr (!C::TWO) {
      __syncwarp();
      const long long fn = f + (long long)gridDim.x * WARPS;
      if (fn < a.B) fetch_prefix(fn, par ^ 1u);
    }
    par ^= 1u;
    tr_wait();   // (filters that never reached the interleaved pass: the barrier's phase still has to be consumed)
    if (st && lane == 0) a.status[f] |= st;
    __syncwarp();
    PH(29);
#ifdef UWVKF_PHASE_TIMING
    if (threadIdx.x == 0) atomicAdd(&g_phase_cycles[39], 1ull);
#endif
  }
}

}  // namespace pose
}  // namespace uwvkf
