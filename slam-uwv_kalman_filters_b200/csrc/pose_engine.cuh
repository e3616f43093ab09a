// PoseUKF engine: one warp per filter, filter block resident in shared memory, prefix-structured
// unscented transform.
//
// What it computes is the ukfom recipe the reference drives (reference call
// sites src/PoseUKF.cpp:462-465 predict, :476-611 updates): sigma points mu (+) (+-L.col(j)) with the
// unscaled lower Cholesky factor, equal-weight iterated manifold mean, 1/2 sum d d^T covariances over all
// 2n+1 points, S/K/gate/downdate, apply_delta re-draw.  HOW it computes it: because L is lower
// triangular and the tangent order is fixed, a model that reads only the first k tangent coordinates
// sees identical sigma points for all columns j >= k, so those 2(n-k) points are carried as a multiplicity
// of the centre point and every covariance block becomes an O(n^2) expression in Sigma and the k x k
// leading Cholesky factor.  tests/test_structured_equivalence.py proves the identity on the CPU.
#pragma once
#include <type_traits>
#include "dmath.cuh"

namespace uwvkf {
namespace pose {

constexpr int N = 53;        // tangent DOF
constexpr int DIM = 54;      // stored doubles of mu
constexpr int NP = 1431;     // packed lower triangle
constexpr int BLK = 1488;    // doubles per filter block in HBM: [mu 54 | pad 2 | Sigma packed 1431 | pad 1]
constexpr int OFF_MU = 0;
constexpr int OFF_S = 56;
constexpr int NOFF = 28;     // Markov offsets per filter

// stored indices of mu
enum { P_POS = 0, P_ORI = 3, P_VEL = 7, P_ACC = 10, P_BG = 13, P_BA = 16, P_G = 19, P_INERTIA = 20, P_LIN = 29,
       P_QUAD = 38, P_WV = 47, P_WVB = 49, P_BADCP = 51, P_RHO = 53 };
// tangent indices
enum { T_POS = 0, T_ORI = 3, T_VEL = 6, T_ACC = 9, T_BG = 12, T_BA = 15, T_G = 18, T_INERTIA = 19, T_LIN = 28,
       T_QUAD = 37, T_WV = 46, T_WVB = 48, T_BADCP = 50, T_RHO = 52 };

enum { ST_NOT_SPD = 1u, ST_NAN_MEAS = 2u, ST_GATE_REJECTED = 4u, ST_MEAN_NOT_CONVERGED = 8u };

enum MeasKind { M_XY = 0, M_Z = 1, M_PRESSURE = 2, M_ACCELERATION = 3, M_VELOCITY = 4, M_WATER_VELOCITY = 5,
                M_GEOGRAPHIC = 6, M_DELAYED_XY = 7, M_BODY_EFFORTS = 8, M_BODY_EFFORTS_VELOCITY_ONLY = 9 };
constexpr int MAXM = 6;      // largest measurement dimension (body efforts)
constexpr int NUWV = 116;    // flat uwv_dynamic_model::UWVParameters: inertia, lin damping, quad damping (6x6 row-major), cog, cob, W, B

struct Params {
  TMercConst tm;
  double gyro_bias_tau, gyro_bias_offset[3];
  double acc_bias_tau, acc_bias_offset[3];
  double inertia_tau, lin_damping_tau, quad_damping_tau;
  double wv_tau, wv_limits, wv_scale, adcp_bias_tau, atm_pressure, rho_tau;
  double imu_in_body[3];
  double mean_tol;
  int mean_max_it;
};

struct Op {
  int type;            // 0 = predict, 1 = update
  int kind;            // MeasKind for updates
  double dt;           // predict
  const double *z;     // device [B][m]
  const double *Rdev;  // device [B][m*m] or nullptr -> R below
  const uint8_t *valid;  // device [B] or nullptr
  double R[MAXM * MAXM];
  double aux[4];
};

struct Launch {
  double *state;           // [B][BLK]
  const double *offsets;   // [B][NOFF]
  const double *rot_rate;  // [B][3]
  uint32_t *status;        // [B]
  const double *Q;         // packed lower, column-major, NP doubles
  const double *uwv;       // NUWV doubles (motion-model parameters of the constructor, src/PoseUKF.cpp:353-354) or nullptr
  double *nav;             // [B][10] or nullptr: the fused kernel also writes p(3) q(4) v(3) of every filter after the step (the
                           // navigation read-back of uwvkf_pose_step_host without a second kernel and a second pass over mu)
  const double *tm;        // [B][10] or nullptr: sin / cos of the latitude of the five distinct sigma-point positions of the predict,
                           // written by pose_tmerc_prepass_kernel right before this launch (pose_fused.cuh, UWVKF_TMERC_PREPASS)
  long long B;
  int nops;
  Op ops[3];
};

// Sigma is stored as the lower triangle packed by columns: column j holds rows j..N-1.
// The first k columns (the "panel" every prefix-k model needs) are therefore a prefix of the array.
__host__ __device__ __forceinline__ int cidx(int i, int j) { return j * N - ((j * (j - 1)) >> 1) + (i - j); }
template <int K>
__device__ __forceinline__ int lidx(int i, int j) { return j * K - ((j * (j - 1)) >> 1) + (i - j); }
__device__ __forceinline__ int sidx(int a, int b) { return a >= b ? cidx(a, b) : cidx(b, a); }

// per-warp shared-memory layout (doubles)
template <int KMAX>
struct WL {
  // rows of G / T / KS (max measurement dim): six for the dense instantiation and for KMAX == 20, the prefix-19 engine with
  // room for the six-dimensional velocity-only body-efforts update (prefix 9)
  static constexpr int MG = (KMAX >= 48 || KMAX == 20) ? MAXM : 3;
  static constexpr int LSZ = ((KMAX * (KMAX + 1) / 2 + 1) / 2) * 2;
  static constexpr int KE = ((KMAX + 1) / 2) * 2;
  static constexpr int L = BLK;             // leading-block Cholesky factor, packed by columns (dimension K)
  static constexpr int INVD = L + LSZ;      // 1 / L_jj
  static constexpr int G = INVD + KE;       // MG x KE : A (or B) then G = A L^-1
  static constexpr int T = G + MG * KE;     // MG x 56 : G Sigma[0:k,:]  (= Pxz^T, then K^T)
  static constexpr int KS = T + MG * 56;    // MG x 56 : (K S)^T
  static constexpr int DL = KS + MG * 56;   // 56 : delta = K nu
  static constexpr int F = DL + 56;         // 56 : diagonal of the affine process map
  static constexpr int TOTAL = F + 56;
};

// Layout of the STREAMED dense update (pose_efforts_kernel): the covariance is not resident.  The factorisation reads it from
// global memory, the rank-6 downdate streams it HBM -> registers -> HBM, and only the six leading columns -- what apply_delta
// works on, a prefix of the packed array -- live in shared memory.  23.6 KB per warp instead of 32.6: 9 warps per SM instead of 6.
struct WLS {
  static constexpr int MG = MAXM;
  static constexpr int S6 = 304;            // 6 leading columns: 6 * 53 - 15 = 303 entries
  static constexpr int LSZ = 1432;          // 53 x 48 panel factor, column-packed with N rows per column
  static constexpr int KE = 54;
  static constexpr int L = OFF_S + S6;
  static constexpr int INVD = L + LSZ;
  static constexpr int G = INVD + KE;
  static constexpr int T = G + MG * KE;
  static constexpr int KS = T + MG * 56;
  static constexpr int DL = KS + MG * 56;
  static constexpr int F = DL + 56;
  static constexpr int TOTAL = F + 56;
};

// ------------------------------------------------------------------------------------------------------
// entry (i, j), i >= j, of a lower triangle packed by columns with R rows per column
template <int R>
__device__ __forceinline__ int lidxr(int i, int j) { return j * R - ((j * (j - 1)) >> 1) + (i - j); }
// BLOCKED left-looking Cholesky of the leading K columns of a column-packed symmetric matrix (SN rows per column), carried over
// R >= K rows (R > K: the panel factor of the leading R x K block); lanes own rows (i = lane + 32 h).  Four columns jb..jb+3 are
// brought up to date together,
//   acc[c] = S[i][jb+c] - sum_{k<jb} L[i][k] L[jb+c][k],
// so one load of L[i][k] and four broadcast loads of L[jb..jb+3][k] (consecutive in the column-packed factor) feed eight FMAs on
// independent accumulators -- the unblocked form spent two loads and three integer instructions per FMA on one dependent chain
// (47 % of all instructions of the body-efforts update, 70 % of the visual-marker update).  The 4 x 4 diagonal block is then
// factorised column by column with the pivots and multipliers travelling by shuffle.  The summation order per entry is the
// sequential one (k ascending).  The factor is stored column-packed with R rows per column (lidxr<R>).
// SRCPIPE: the source lives in global memory (streamed dense update): the entries of block jb + 4 are requested before the
// k-loop of block jb runs, so their latency hides behind it.
template <int K, int R, int SN, bool FAST, bool SRCPIPE = false>
__device__ __forceinline__ bool chol_blocked(const double *S, double *Lw, double *invd, int lane) {
  bool ok = true;
  double nxt[2][4];      // SRCPIPE: source entries of the next block (rows jb' + lane, jb' + lane + 32)
  (void)nxt;
  auto fetch = [&](int jb, double (&dst)[2][4]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = jb + lane + 32 * h;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int cc = jb + c < K ? jb + c : K - 1;                       // (clamped: the tail block is narrower)
        const int ii = (i >= cc && i < R) ? i : cc;                       // in-range address for lanes without an entry
        dst[h][c] = S[lidxr<SN>(ii, cc)];
      }
    }
  };
  if constexpr (SRCPIPE) fetch(0, nxt);
  // Rows are dealt out RELATIVE to the block: lane l works on rows jb + l (and jb + l + 32 while the block still has more than 32
  // rows below its top), so no lane idles on rows above the block, the second round disappears for the later blocks, and the
  // pivot row jb + c always sits on lane c.
  auto block = [&](int jb, auto bw_, auto rnd_) {
    constexpr int BW = decltype(bw_)::value, RND = decltype(rnd_)::value;
    double acc[RND][BW];
#pragma unroll
    for (int h = 0; h < RND; ++h) {
      const int i = jb + lane + 32 * h;
#pragma unroll
      for (int c = 0; c < BW; ++c) {
        if constexpr (SRCPIPE) acc[h][c] = (i >= jb + c && i < R) ? nxt[h][c] : 0.0;
        else acc[h][c] = (i >= jb + c && i < R) ? S[lidxr<SN>(i, jb + c)] : 0.0;
      }
    }
    if constexpr (SRCPIPE) {
      if (jb + 4 < K) fetch(jb + 4, nxt);
    }
    int cb = 0;                       // lidxr<R>(0, k) - k: Lw[cb + i] = L[i][k]
#pragma unroll 2
    for (int k = 0; k < jb; ++k) {
      double lj[BW];
#pragma unroll
      for (int c = 0; c < BW; ++c) lj[c] = Lw[cb + jb + c];
#pragma unroll
      for (int h = 0; h < RND; ++h) {
        const int i = jb + lane + 32 * h;
        const double li = Lw[cb + (i < R ? i : jb)];     // lanes past the last row run on a clamped (in-range) address
#pragma unroll
        for (int c = 0; c < BW; ++c) acc[h][c] = fma(-li, lj[c], acc[h][c]);
      }
      cb += R - k - 1;
    }
#pragma unroll
    for (int c = 0; c < BW; ++c) {
      const int r = jb + c;                                  // pivot row / column: lane c, round 0
      const double d = __shfl_sync(kFull, acc[0][c], c);
      ok = ok && (d > 0.0);
      const double inv = FAST ? fast_rsqrt3(d) : rsqrt(d);
      double lcol[RND];
#pragma unroll
      for (int h = 0; h < RND; ++h) {
        const int i = jb + lane + 32 * h;
        lcol[h] = acc[h][c] * inv;
        if (i >= r && i < R) Lw[lidxr<R>(i, r)] = lcol[h];
      }
      if (invd != nullptr && lane == 0) invd[r] = inv;
#pragma unroll
      for (int c2 = c + 1; c2 < BW; ++c2) {
        const double l2 = __shfl_sync(kFull, lcol[0], c2);   // L[jb + c2][r]
#pragma unroll
        for (int h = 0; h < RND; ++h) acc[h][c2] = fma(-lcol[h], l2, acc[h][c2]);
      }
    }
    __syncwarp();
  };
  constexpr int NBLK = K / 4, TAIL = K % 4;
  using I1 = std::integral_constant<int, 1>;
  using I2 = std::integral_constant<int, 2>;
  using I4 = std::integral_constant<int, 4>;
  static_assert(R <= 64, "two rounds of rows");
  for (int b = 0; b < NBLK; ++b) {
    if (4 * b + 32 < R) block(4 * b, I4{}, I2{});
    else block(4 * b, I4{}, I1{});
  }
  if constexpr (TAIL > 0) block(4 * NBLK, std::integral_constant<int, (TAIL > 0 ? TAIL : 1)>{}, I1{});
  return ok;
}

// Cholesky of the leading K x K block of Sigma.  K <= 32: lane i owns row i in registers, pivots and
// column entries travel by shuffles (no shared-memory round trips inside the recurrence).
// R > K (general path only): the factorisation is carried over R rows -- the first K columns of the factor of the leading
// R x R block, i.e. the PANEL FACTOR Sigma[0:R, 0:K] L11^-T -- stored column-packed with R rows per column (lidxr<R>).
template <int K, int R = K, bool SRCPIPE = false>
__device__ __forceinline__ bool chol_leading(const double *S, double *Lw, double *invd, int lane) {
  static_assert(R == K || K > 32, "extra rows only in the general path");
  if constexpr (K <= 32) {
    double a[K];
    const int i = lane;
#pragma unroll
    for (int c = 0; c < K; ++c) a[c] = (i < K && c <= i) ? S[cidx(i, c)] : 0.0;
    bool ok = true;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      double d = bcast(a[j], j);
      ok = ok && (d > 0.0);
      double inv = fast_rsqrt(d);
      a[j] = (i == j) ? d * inv : a[j] * inv;
      if (i == j) invd[j] = inv;
#pragma unroll
      for (int c = j + 1; c < K; ++c) {
        double lcj = bcast(a[j], c);
        a[c] = fma(-a[j], lcj, a[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < K; ++c)
      if (i < K && c <= i) Lw[lidx<K>(i, c)] = a[c];
    __syncwarp();
    return ok;
  } else {
    // general path (K up to N): blocked left-looking factorisation, lanes own rows (chol_blocked)
    return chol_blocked<K, R, N, false, SRCPIPE>(S, Lw, invd, lane);
  }
}

// G <- G L^-1 for the K x K lower factor L (G is M x K, row stride gs).
template <int K, int M>
__device__ __forceinline__ void solve_right(double *G, int gs, const double *Lw, const double *invd, int lane) {
  if constexpr (K <= 32) {
    double acc[M];
#pragma unroll
    for (int r = 0; r < M; ++r) acc[r] = (lane < K) ? G[r * gs + lane] : 0.0;
    for (int c = K - 1; c >= 0; --c) {
      double ic = invd[c];
      double g[M];
#pragma unroll
      for (int r = 0; r < M; ++r) g[r] = bcast(acc[r] * ic, c);
      if (lane < c) {
        double l = Lw[lidx<K>(c, lane)];
#pragma unroll
        for (int r = 0; r < M; ++r) acc[r] = fma(-g[r], l, acc[r]);
      } else if (lane == c) {
#pragma unroll
        for (int r = 0; r < M; ++r) acc[r] = g[r];
      }
    }
    if (lane < K) {
#pragma unroll
      for (int r = 0; r < M; ++r) G[r * gs + lane] = acc[r];
    }
    __syncwarp();
  } else {
    for (int c = K - 1; c >= 0; --c) {
      double ic = invd[c];
      double g[M];
#pragma unroll
      for (int r = 0; r < M; ++r) g[r] = G[r * gs + c] * ic;
      __syncwarp();
      for (int j = lane; j < c; j += 32) {
        double l = Lw[lidx<K>(c, j)];
#pragma unroll
        for (int r = 0; r < M; ++r) G[r * gs + j] = fma(-g[r], l, G[r * gs + j]);
      }
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < M; ++r) G[r * gs + c] = g[r];
      }
      __syncwarp();
    }
  }
}

// T[r][x] = sum_{b<K} G[r][b] * Sigma[b][x]   for x = 0..N-1
template <int K, int M>
__device__ __forceinline__ void g_times_sigma(const double *G, int gs, const double *S, double *T, int lane) {
  for (int x = lane; x < N; x += 32) {
    double acc[M];
#pragma unroll
    for (int r = 0; r < M; ++r) acc[r] = 0.0;
#pragma unroll 4
    for (int b = 0; b < K; ++b) {
      double s = S[sidx(b, x)];
#pragma unroll
      for (int r = 0; r < M; ++r) acc[r] = fma(G[r * gs + b], s, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < M; ++r) T[r * 56 + x] = acc[r];
  }
  __syncwarp();
}

// T[r][x] = sum_{j<K} Lp[x][j] * B[r][j] for x = 0..N-1, with Lp the panel factor over all N rows (chol_leading<K, N>):
// Pxz = Sigma[:, :K] L11^-T B^T without the triangular solve G = B L11^-1 and without re-reading Sigma[:K, :].
template <int K, int M>
__device__ __forceinline__ void panel_times_b(const double *G, int gs, const double *Lw, double *T, int lane) {
  double acc[2][M];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int r = 0; r < M; ++r) acc[h][r] = 0.0;
  const int x1 = lane + 32 < N ? lane + 32 : N - 1;
  int cb = 0;   // lidxr<N>(0, j) - j
#pragma unroll 4
  for (int j = 0; j < K; ++j) {
    const double l0 = lane >= j ? Lw[cb + lane] : 0.0;     // rows above the diagonal are zeros of the factor
    const double l1 = x1 >= j ? Lw[cb + x1] : 0.0;
#pragma unroll
    for (int r = 0; r < M; ++r) {
      const double b = G[r * gs + j];
      acc[0][r] = fma(b, l0, acc[0][r]);
      acc[1][r] = fma(b, l1, acc[1][r]);
    }
    cb += N - j - 1;
  }
#pragma unroll
  for (int r = 0; r < M; ++r) {
    T[r * 56 + lane] = acc[0][r];
    if (lane + 32 < N) T[r * 56 + lane + 32] = acc[1][r];
  }
  __syncwarp();
}

// equal-weight iterated SO3 mean over 2N+1 logical points; each lane holds ROUNDS points with
// multiplicities wt (ukfom mean_sigma_points restricted to the one non-vector component).
template <int ROUNDS>
__device__ __forceinline__ Quat so3_mean(const Quat (&qp)[ROUNDS], const double (&wt)[ROUNDS], double tol, int max_it,
                                         unsigned &status) {
  Quat r{bcast(qp[0].x, 0), bcast(qp[0].y, 0), bcast(qp[0].z, 0), bcast(qp[0].w, 0)};
  int it = 0;
  double nrm;
  do {
    double mx = 0, my = 0, mz = 0;
#pragma unroll
    for (int k = 0; k < ROUNDS; ++k) {
      Vec3 d = so3_boxminus(qp[k], r);
      mx = fma(wt[k], d.x, mx); my = fma(wt[k], d.y, my); mz = fma(wt[k], d.z, mz);
    }
    mx = wsum(mx) / double(2 * N + 1);
    my = wsum(my) / double(2 * N + 1);
    mz = wsum(mz) / double(2 * N + 1);
    r = so3_boxplus(r, Vec3{mx, my, mz}, 1.0);
    nrm = sqrt(mx * mx + my * my + mz * mz);
  } while (nrm > tol && ++it < max_it);
  if (it >= max_it) status |= ST_MEAN_NOT_CONVERGED;
  return r;
}

// Unscented transform of the orientation rows with the leading-K sigma points qp (one round, K <= 15):
// returns the new mean quaternion, writes Sigma_oo (6 values) and T = (A L^-1) Sigma[0:K,:] (3 x N).
template <int K, int KMAX, class W = WL<KMAX>>
__device__ __forceinline__ Quat ori_transform(double *w, const Quat &qp, int p, const Params &P, double soo[6], int lane,
                                              unsigned &status) {
  double *S = w + OFF_S, *Lw = w + W::L, *invd = w + W::INVD, *G = w + W::G, *T = w + W::T;
  const bool act = p < 2 * K + 1;
  const bool on = act && p > 0;
  const int j = (p - 1) >> 1;
  Quat qa[1] = {qp};
  double wa[1] = {p == 0 ? double(1 + 2 * (N - K)) : (act ? 1.0 : 0.0)};
  Quat r = so3_mean<1>(qa, wa, P.mean_tol, P.mean_max_it, status);
  Vec3 a = so3_boxminus(qp, r);
  const double hw = 0.5 * wa[0];
  soo[0] = wsum(hw * a.x * a.x);
  soo[1] = wsum(hw * a.y * a.x);
  soo[2] = wsum(hw * a.z * a.x);
  soo[3] = wsum(hw * a.y * a.y);
  soo[4] = wsum(hw * a.z * a.y);
  soo[5] = wsum(hw * a.z * a.z);
  // A_j = 1/2 (a_{2j+1} - a_{2j+2}): odd lanes pair with lane + 1
  double nx = __shfl_down_sync(kFull, a.x, 1), ny = __shfl_down_sync(kFull, a.y, 1), nz = __shfl_down_sync(kFull, a.z, 1);
  if (on && (p & 1)) {
    G[0 * W::KE + j] = 0.5 * (a.x - nx);
    G[1 * W::KE + j] = 0.5 * (a.y - ny);
    G[2 * W::KE + j] = 0.5 * (a.z - nz);
  }
  __syncwarp();
  solve_right<K, 3>(G, W::KE, Lw, invd, lane);
  g_times_sigma<K, 3>(G, W::KE, S, T, lane);
  return r;
}

// In-place congruence Sigma <- C Sigma C^T with C = I + dt * E(dst <- src) on 3 consecutive rows.
__device__ __forceinline__ void couple3(double *S, int d0, int s0, double dt, int lane) {
  double nb = 0.0;
  int bi = 0, bj = 0;
  if (lane < 6) {
    bi = (lane < 3) ? lane : (lane < 5 ? lane - 2 : 2);
    bj = (lane < 3) ? 0 : (lane < 5 ? 1 : 2);
    nb = S[sidx(d0 + bi, d0 + bj)] + dt * (S[sidx(s0 + bi, d0 + bj)] + S[sidx(d0 + bi, s0 + bj)]) +
         dt * dt * S[sidx(s0 + bi, s0 + bj)];
  }
  __syncwarp();
  for (int b = lane; b < N; b += 32) {
    if (b < d0 || b >= d0 + 3) {
#pragma unroll
      for (int i = 0; i < 3; ++i) S[sidx(d0 + i, b)] += dt * S[sidx(s0 + i, b)];
    }
  }
  __syncwarp();
  if (lane < 6) S[cidx(d0 + bi, d0 + bj)] = nb;
  __syncwarp();
}

// diagonal of the affine part of the process model (reference src/PoseUKF.cpp:37-78)
__device__ __forceinline__ double affine_factor(int t, const Params &P, double dt) {
  if (t < T_BG) return 1.0;
  if (t < T_BA) return 1.0 - dt / P.gyro_bias_tau;
  if (t < T_G) return 1.0 - dt / P.acc_bias_tau;
  if (t < T_INERTIA) return 1.0;
  if (t < T_LIN) return 1.0 - dt / P.inertia_tau;
  if (t < T_QUAD) return 1.0 - dt / P.lin_damping_tau;
  if (t < T_WV) return 1.0 - dt / P.quad_damping_tau;
  if (t < T_BADCP) return 1.0 - dt / P.wv_tau;
  if (t < T_RHO) return 1.0 - dt / P.adcp_bias_tau;
  return 1.0 - dt / P.rho_tau;
}

// new value of stored component s (not orientation) under the process model, src/PoseUKF.cpp:25-78
__device__ __forceinline__ double affine_mean(int s, const double *mu, const double *off, const Params &P, double dt) {
  double x = mu[s];
  if (s < P_ORI) return x + dt * mu[P_VEL + s];
  if (s < P_ACC) return x + dt * mu[s + 3];   // velocity += acceleration * dt
  if (s < P_BG) return x;
  if (s < P_BA) return x + dt * ((-1.0 / P.gyro_bias_tau) * (x - P.gyro_bias_offset[s - P_BG]));
  if (s < P_G) return x + dt * ((-1.0 / P.acc_bias_tau) * (x - P.acc_bias_offset[s - P_BA]));
  if (s < P_INERTIA) return x;
  if (s < P_LIN) return x + dt * ((-1.0 / P.inertia_tau) * (x - off[s - P_INERTIA]));
  if (s < P_QUAD) return x + dt * ((-1.0 / P.lin_damping_tau) * (x - off[9 + s - P_LIN]));
  if (s < P_WV) return x + dt * ((-1.0 / P.quad_damping_tau) * (x - off[18 + s - P_QUAD]));
  if (s < P_BADCP) return x + dt * ((-1.0 / P.wv_tau) * x);
  if (s < P_RHO) return x + dt * ((-1.0 / P.adcp_bias_tau) * x);
  return x + dt * ((-1.0 / P.rho_tau) * (x - off[27]));
}

// ------------------------------------------------------------------------------------------------------
// predictionStepImpl (reference src/PoseUKF.cpp:446-466) with processModel (src/PoseUKF.cpp:12-84).
// The orientation is the only non-affine output and reads (p, q, b_gyro) = tangent prefix 15.
template <int KMAX>
__device__ __forceinline__ bool predict(double *w, const double *Qs, const uint16_t *lut, const double *off,
                                        const double *wm_g, const Params &P, double dt, int lane, unsigned &status) {
  using W = WL<KMAX>;
  constexpr int K = 15;
  double *mu = w + OFF_MU, *S = w + OFF_S, *Lw = w + W::L, *invd = w + W::INVD, *T = w + W::T, *fv = w + W::F;
  // a NaN rotation rate (device-resident control input the host could not scan) would poison the state: refuse the prediction
  if (isnan(wm_g[0]) || isnan(wm_g[1]) || isnan(wm_g[2])) { status |= ST_NAN_MEAS; return false; }
  if (!chol_leading<K>(S, Lw, invd, lane)) { status |= ST_NOT_SPD; return false; }
  const Quat q{mu[P_ORI], mu[P_ORI + 1], mu[P_ORI + 2], mu[P_ORI + 3]};
  const Vec3 vel{mu[P_VEL], mu[P_VEL + 1], mu[P_VEL + 2]};
  const Vec3 wm{wm_g[0], wm_g[1], wm_g[2]};
  const int p = lane;
  const bool on = p > 0 && p < 2 * K + 1;
  const int j = (p - 1) >> 1;
  const double sg = (p & 1) ? 1.0 : -1.0;
  auto d = [&](int t) { return (on && t >= j) ? sg * Lw[lidx<K>(t, j)] : 0.0; };
  Quat qp;
  {
    double px = mu[P_POS] + d(0), py = mu[P_POS + 1] + d(1);
    Quat qq = so3_boxplus(q, Vec3{d(3), d(4), d(5)}, 1.0);
    Vec3 wb{wm.x - (mu[P_BG] + d(12)), wm.y - (mu[P_BG + 1] + d(13)), wm.z - (mu[P_BG + 2] + d(14))};
    double lat = tmerc_nav_to_lat(P.tm, px, py);
    double sl, cl;
    sincos(lat, &sl, &cl);
    Vec3 wn = qrot(qq, wb);
    wn.x -= kEarthW * cl;
    wn.z -= kEarthW * sl;
    qp = so3_boxplus(qq, wn, dt);
  }
  double soo[6];
  Quat r = ori_transform<K, KMAX>(w, qp, p, P, soo, lane, status);

  // effective process noise pieces (src/PoseUKF.cpp:448-460)
  const double dt2 = dt * dt;
  double qoo[6];
  {
    double Rm[9], Qo[9], t[9];
    qmat(q, Rm);
    Qo[0] = Qs[cidx(3, 3)]; Qo[1] = Qs[cidx(4, 3)]; Qo[2] = Qs[cidx(5, 3)];
    Qo[3] = Qo[1];          Qo[4] = Qs[cidx(4, 4)]; Qo[5] = Qs[cidx(5, 4)];
    Qo[6] = Qo[2];          Qo[7] = Qo[5];          Qo[8] = Qs[cidx(5, 5)];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) t[a * 3 + b] = Rm[a * 3] * Qo[b] + Rm[a * 3 + 1] * Qo[3 + b] + Rm[a * 3 + 2] * Qo[6 + b];
    // (i, j) order matches soo: (0,0) (1,0) (2,0) (1,1) (2,1) (2,2)
    auto rq = [&](int a, int b) { return t[a * 3] * Rm[b * 3] + t[a * 3 + 1] * Rm[b * 3 + 1] + t[a * 3 + 2] * Rm[b * 3 + 2]; };
    qoo[0] = rq(0, 0); qoo[1] = rq(1, 0); qoo[2] = rq(2, 0); qoo[3] = rq(1, 1); qoo[4] = rq(2, 1); qoo[5] = rq(2, 2);
  }
  const double add = P.wv_scale * (vel.x * vel.x + vel.y * vel.y + 100.0 * vel.z * vel.z) * dt;

  // affine map: diagonal factors and new mean (registers until Sigma is done)
  for (int t = lane; t < N; t += 32) fv[t] = affine_factor(t, P, dt);
  double nm0 = 0, nm1 = 0;
  {
    int s0 = lane, s1 = lane + 32;
    if (s0 < P_ORI || s0 >= P_VEL) nm0 = affine_mean(s0, mu, off, P, dt);
    if (s1 < DIM) nm1 = affine_mean(s1, mu, off, P, dt);
  }
  __syncwarp();
  // Sigma <- F Sigma F^T on the affine rows: couplings p <- v, then v <- a, then the diagonal scaling + Q
  couple3(S, T_POS, T_VEL, dt, lane);
  couple3(S, T_VEL, T_ACC, dt, lane);
  for (int e = lane; e < NP; e += 32) {
    unsigned ij = lut[e];
    int i = ij & 0xff, jj = ij >> 8;
    if ((i >= 3 && i < 6) || (jj >= 3 && jj < 6)) continue;
    double v = fv[i] * fv[jj] * S[e] + dt2 * Qs[e];
    if (i == jj && i >= T_WV && i < T_BADCP) v += dt2 * add;
    S[e] = v;
  }
  // orientation rows: T F^T  (+ Q cross terms, zero in the reference's configuration but part of the API)
  for (int x = lane; x < N; x += 32) {
    if (x >= 3 && x < 6) continue;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double tp = fv[x] * T[c * 56 + x];
      if (x < 3) tp += dt * T[c * 56 + x + 6];
      else if (x >= T_VEL && x < T_ACC) tp += dt * T[c * 56 + x + 3];
      int e = sidx(3 + c, x);
      S[e] = tp + dt2 * Qs[e];
    }
  }
  if (lane < 6) {
    int bi = (lane < 3) ? lane : (lane < 5 ? lane - 2 : 2);
    int bj = (lane < 3) ? 0 : (lane < 5 ? 1 : 2);
    S[cidx(3 + bi, 3 + bj)] = soo[lane] + dt2 * qoo[lane];
  }
  // mean
  {
    int s0 = lane, s1 = lane + 32;
    if (s0 < P_ORI || s0 >= P_VEL) mu[s0] = nm0;
    if (s1 < DIM) mu[s1] = nm1;
    if (lane == 0) { mu[P_ORI] = r.x; mu[P_ORI + 1] = r.y; mu[P_ORI + 2] = r.z; mu[P_ORI + 3] = r.w; }
  }
  __syncwarp();
  return true;
}

// ukfom apply_delta(delta): only the orientation (tangent 3..5, prefix 6) is a non-vector component.
template <int KMAX, class W = WL<KMAX>>
__device__ __forceinline__ bool apply_delta(double *w, const Params &P, int lane, unsigned &status) {
  constexpr int K = 6;
  double *mu = w + OFF_MU, *S = w + OFF_S, *Lw = w + W::L, *invd = w + W::INVD, *T = w + W::T, *dl = w + W::DL;
  if (!chol_leading<K>(S, Lw, invd, lane)) { status |= ST_NOT_SPD; return false; }
  const Quat q{mu[P_ORI], mu[P_ORI + 1], mu[P_ORI + 2], mu[P_ORI + 3]};
  const int p = lane;
  const bool on = p > 0 && p < 2 * K + 1;
  const int j = (p - 1) >> 1;
  const double sg = (p & 1) ? 1.0 : -1.0;
  auto d = [&](int t) { return (on && t >= j) ? sg * Lw[lidx<K>(t, j)] : 0.0; };
  Quat qp = so3_boxplus(q, Vec3{dl[3] + d(3), dl[4] + d(4), dl[5] + d(5)}, 1.0);
  double soo[6];
  Quat r = ori_transform<K, KMAX, W>(w, qp, p, P, soo, lane, status);
  for (int x = lane; x < N; x += 32) {
    if (x >= 3 && x < 6) continue;
#pragma unroll
    for (int c = 0; c < 3; ++c) S[sidx(3 + c, x)] = T[c * 56 + x];
  }
  if (lane < 6) {
    int bi = (lane < 3) ? lane : (lane < 5 ? lane - 2 : 2);
    int bj = (lane < 3) ? 0 : (lane < 5 ? 1 : 2);
    S[cidx(3 + bi, 3 + bj)] = soo[lane];
  }
  for (int s = lane; s < DIM; s += 32) {
    if (s < P_ORI) mu[s] += dl[s];
    else if (s >= P_VEL) mu[s] += dl[s - 1];
  }
  if (lane == 0) { mu[P_ORI] = r.x; mu[P_ORI + 1] = r.y; mu[P_ORI + 2] = r.z; mu[P_ORI + 3] = r.w; }
  __syncwarp();
  return true;
}

// ---- measurement models (reference src/PoseUKF.cpp:87-151): K = tangent prefix read, M = dimension ----
template <int KIND> struct Meas;
template <> struct Meas<M_XY> {   // src/PoseUKF.cpp:87-92
  static constexpr int K = 2, M = 2;
  template <class D> __device__ static void eval(const double *mu, D d, const double *, const Params &, double *o, const double * = nullptr) {
    o[0] = mu[P_POS] + d(0); o[1] = mu[P_POS + 1] + d(1);
  }
};
template <> struct Meas<M_Z> {    // src/PoseUKF.cpp:100-105
  static constexpr int K = 3, M = 1;
  template <class D> __device__ static void eval(const double *mu, D d, const double *, const Params &, double *o, const double * = nullptr) {
    o[0] = mu[P_POS + 2] + d(2);
  }
};
template <> struct Meas<M_VELOCITY> {   // src/PoseUKF.cpp:117-123
  static constexpr int K = 9, M = 3;
  template <class D> __device__ static void eval(const double *mu, D d, const double *, const Params &, double *o, const double * = nullptr) {
    Quat q = so3_boxplus(Quat{mu[P_ORI], mu[P_ORI + 1], mu[P_ORI + 2], mu[P_ORI + 3]}, Vec3{d(3), d(4), d(5)}, 1.0);
    Vec3 v = qrot_inv(q, Vec3{mu[P_VEL] + d(6), mu[P_VEL + 1] + d(7), mu[P_VEL + 2] + d(8)});
    o[0] = v.x; o[1] = v.y; o[2] = v.z;
  }
};
template <> struct Meas<M_ACCELERATION> {   // src/PoseUKF.cpp:125-131
  static constexpr int K = 19, M = 3;
  template <class D> __device__ static void eval(const double *mu, D d, const double *, const Params &, double *o, const double * = nullptr) {
    Quat q = so3_boxplus(Quat{mu[P_ORI], mu[P_ORI + 1], mu[P_ORI + 2], mu[P_ORI + 3]}, Vec3{d(3), d(4), d(5)}, 1.0);
    Vec3 a{mu[P_ACC] + d(9), mu[P_ACC + 1] + d(10), (mu[P_ACC + 2] + d(11)) + (mu[P_G] + d(18))};
    Vec3 v = qrot_inv(q, a);
    o[0] = v.x + (mu[P_BA] + d(15)); o[1] = v.y + (mu[P_BA + 1] + d(16)); o[2] = v.z + (mu[P_BA + 2] + d(17));
  }
};
template <> struct Meas<M_PRESSURE> {   // src/PoseUKF.cpp:107-115 ; aux = pressure_sensor_in_imu
  static constexpr int K = 53, M = 1;
  template <class D> __device__ static void eval(const double *mu, D d, const double *aux, const Params &P, double *o, const double * = nullptr) {
    Quat q = so3_boxplus(Quat{mu[P_ORI], mu[P_ORI + 1], mu[P_ORI + 2], mu[P_ORI + 3]}, Vec3{d(3), d(4), d(5)}, 1.0);
    Vec3 rn = qrot(q, Vec3{aux[0], aux[1], aux[2]});
    double zn = (mu[P_POS + 2] + d(2)) + rn.z;
    o[0] = P.atm_pressure - zn * (mu[P_G] + d(18)) * (mu[P_RHO] + d(52));
  }
};
template <> struct Meas<M_WATER_VELOCITY> {   // src/PoseUKF.cpp:133-151 ; aux[0] = cell_weighting
  static constexpr int K = 52, M = 2;
  template <class D> __device__ static void eval(const double *mu, D d, const double *aux, const Params &, double *o, const double * = nullptr) {
    Quat q = so3_boxplus(Quat{mu[P_ORI], mu[P_ORI + 1], mu[P_ORI + 2], mu[P_ORI + 3]}, Vec3{d(3), d(4), d(5)}, 1.0);
    double vx = mu[P_VEL] + d(6), vy = mu[P_VEL + 1] + d(7), vz = mu[P_VEL + 2] + d(8);
    Vec3 b = qrot_inv(q, Vec3{vx - (mu[P_WVB] + d(48)), vy - (mu[P_WVB + 1] + d(49)), vz - 0.0});
    Vec3 a = qrot_inv(q, Vec3{vx - (mu[P_WV] + d(46)), vy - (mu[P_WV + 1] + d(47)), vz - 0.0});
    double wgt = aux[0];
    o[0] = wgt * b.x + (1 - wgt) * a.x + (mu[P_BADCP] + d(50));
    o[1] = wgt * b.y + (1 - wgt) * a.y + (mu[P_BADCP + 1] + d(51));
  }
};


// ---- uwv_dynamic_model::DynamicModel::calcEfforts with the inertia / damping blocks of a sigma point ----------------
// tau = M nu_dot + C(M, nu) + D_lin nu + D_quad (|nu| o nu) + g(q)   (SURVEY.md Appendix D).
// `base` = flat UWVParameters; blk(which, i, j) returns entry (i, j) of the state's 3x3 inertia (0) / linear (1) /
// quadratic (2) damping block, which replaces rows / columns {0, 1, 5} of the 6x6 matrix (reference src/PoseUKF.cpp:160-171).
template <bool SUBST, class Blk>
__device__ __forceinline__ void calc_efforts_dev(const double *base, Blk blk, const double *acc6, const double *vel6, const Quat &q,
                                                 double *out) {
  auto el = [&](int which, int r, int c) {
    const bool sr = (r < 2 || r == 5), sc = (c < 2 || c == 5);
    if (SUBST && sr && sc) return blk(which, r == 5 ? 2 : r, c == 5 ? 2 : c);
    return base[which * 36 + r * 6 + c];
  };
  double mv[6], ma[6], dl[6], dq[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const double m = el(0, r, c);
      s0 += m * vel6[c];
      s1 += m * acc6[c];
      s2 += el(1, r, c) * vel6[c];
      s3 += el(2, r, c) * (fabs(vel6[c]) * vel6[c]);
    }
    mv[r] = s0; ma[r] = s1; dl[r] = s2; dq[r] = s3;
  }
  auto cross = [](const double *a, const double *b, double *o) {
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
  };
  double c0[3], c1[3], c2[3];
  cross(mv, vel6 + 3, c0);
  cross(mv, vel6, c1);
  cross(mv + 3, vel6 + 3, c2);
  const double W = base[114], Bu = base[115];
  const Vec3 gf = qrot_inv(q, Vec3{0, 0, W - Bu});
  const Vec3 gu = qrot_inv(q, Vec3{0, 0, 1});
  const double lever[3] = {base[108] * W - base[111] * Bu, base[109] * W - base[112] * Bu, base[110] * W - base[113] * Bu};
  const double guv[3] = {gu.x, gu.y, gu.z}, gfv[3] = {gf.x, gf.y, gf.z};
  double gt[3];
  cross(lever, guv, gt);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    out[i] = ma[i] + (c0[i] + dl[i] + dq[i] + gfv[i]);
    out[3 + i] = ma[3 + i] + (c1[i] + c2[i] + dl[3 + i] + dq[3 + i] + gt[i]);
  }
}

// per-filter context of the body-efforts models, evaluated at the MEAN before the update (reference src/PoseUKF.cpp:585-598):
// ctx[0..2] = rotation_rate_body = getRotationRate(), [3..5] = w x imu_in_body, [6..8] = w x (w x imu_in_body),
// velocity-only variant: [9..11] = q^-1 water_velocity, [12..14] = acceleration_body, [15..18] = q
__device__ __forceinline__ void efforts_context(const double *mu, const double *wm, const Params &P, double *ctx) {
  const Quat q{mu[P_ORI], mu[P_ORI + 1], mu[P_ORI + 2], mu[P_ORI + 3]};
  const double lat = tmerc_nav_to_lat(P.tm, mu[P_POS], mu[P_POS + 1]);
  const Vec3 eb = qrot_inv(q, Vec3{kEarthW * cos(lat), 0.0, kEarthW * sin(lat)});
  const double w[3] = {wm[0] - mu[P_BG] - eb.x, wm[1] - mu[P_BG + 1] - eb.y, wm[2] - mu[P_BG + 2] - eb.z};
  const double *r = P.imu_in_body;
  const double c[3] = {w[1] * r[2] - w[2] * r[1], w[2] * r[0] - w[0] * r[2], w[0] * r[1] - w[1] * r[0]};
  const double cc[3] = {w[1] * c[2] - w[2] * c[1], w[2] * c[0] - w[0] * c[2], w[0] * c[1] - w[1] * c[0]};
  const Vec3 wvb = qrot_inv(q, Vec3{mu[P_WV], mu[P_WV + 1], 0.0});
  const Vec3 ab = qrot_inv(q, Vec3{mu[P_ACC], mu[P_ACC + 1], mu[P_ACC + 2]});
#pragma unroll
  for (int k = 0; k < 3; ++k) { ctx[k] = w[k]; ctx[3 + k] = c[k]; ctx[6 + k] = cc[k]; }
  ctx[9] = wvb.x; ctx[10] = wvb.y; ctx[11] = wvb.z;
  ctx[12] = ab.x - cc[0]; ctx[13] = ab.y - cc[1]; ctx[14] = ab.z - cc[2];
  ctx[15] = q.x; ctx[16] = q.y; ctx[17] = q.z; ctx[18] = q.w;
}

template <> struct Meas<M_BODY_EFFORTS> {   // measurementEfforts, src/PoseUKF.cpp:153-196; aux = ctx, model = UWVParameters
  static constexpr int K = 48, M = 6;       // reads q, v, a, the 27 model states and the water velocity (tangent 3..47)
  template <class D> __device__ static void eval(const double *mu, D d, const double *ctx, const Params &, double *o, const double *model) {
    const Quat q = so3_boxplus(Quat{mu[P_ORI], mu[P_ORI + 1], mu[P_ORI + 2], mu[P_ORI + 3]}, Vec3{d(3), d(4), d(5)}, 1.0);
    const Vec3 vb = qrot_inv(q, Vec3{mu[P_VEL] + d(6), mu[P_VEL + 1] + d(7), mu[P_VEL + 2] + d(8)});
    const Vec3 wvb = qrot_inv(q, Vec3{mu[P_WV] + d(T_WV), mu[P_WV + 1] + d(T_WV + 1), 0.0});
    const Vec3 ab = qrot_inv(q, Vec3{mu[P_ACC] + d(9), mu[P_ACC + 1] + d(10), mu[P_ACC + 2] + d(11)});
    const double vel6[6] = {vb.x - ctx[3] - wvb.x, vb.y - ctx[4] - wvb.y, vb.z - ctx[5] - wvb.z, ctx[0], ctx[1], ctx[2]};
    const double acc6[6] = {ab.x - ctx[6], ab.y - ctx[7], ab.z - ctx[8], 0.0, 0.0, 0.0};
    // state blocks are MTK::matrix<3,3>, vectorised column-major (src/PoseState.hpp:37)
    auto blk = [&](int which, int i, int j) { const int k = which * 9 + j * 3 + i; return mu[P_INERTIA + k] + d(T_INERTIA + k); };
    calc_efforts_dev<true>(model, blk, acc6, vel6, q, o);
  }
};
template <> struct Meas<M_BODY_EFFORTS_VELOCITY_ONLY> {   // constrainVelocity, src/PoseUKF.cpp:198-219: reads the velocity only
  static constexpr int K = 9, M = 6;
  template <class D> __device__ static void eval(const double *mu, D d, const double *ctx, const Params &, double *o, const double *model) {
    const Quat q{ctx[15], ctx[16], ctx[17], ctx[18]};
    const Vec3 vb = qrot_inv(q, Vec3{mu[P_VEL] + d(6), mu[P_VEL + 1] + d(7), mu[P_VEL + 2] + d(8)});
    const double vel6[6] = {vb.x - ctx[3] - ctx[9], vb.y - ctx[4] - ctx[10], vb.z - ctx[5] - ctx[11], ctx[0], ctx[1], ctx[2]};
    const double acc6[6] = {ctx[12], ctx[13], ctx[14], 0.0, 0.0, 0.0};
    auto blk = [&](int, int, int) { return 0.0; };
    calc_efforts_dev<false>(model, blk, acc6, vel6, q, o);
  }
};

template <int M>
__device__ __forceinline__ void sym_inverse(const double *S, double *Si) {
  if constexpr (M == 1) {
    Si[0] = 1.0 / S[0];
  } else if constexpr (M == 2) {
    double id = 1.0 / (S[0] * S[3] - S[1] * S[2]);
    Si[0] = S[3] * id; Si[1] = -S[1] * id; Si[2] = -S[2] * id; Si[3] = S[0] * id;
  } else if constexpr (M > 3) {
    // Gauss-Jordan with partial pivoting (what Eigen's fixed-size inverse does above 4x4, as SURVEY.md Appendix B.2 describes),
    // one COLUMN of the augmented matrix [S | I] per lane (lanes 0 .. 2M-1): row operations act on every column independently,
    // the pivot column travels by shuffle, and nothing is indexed at run time.  (Every lane eliminating its own copy of the
    // 2 x M x M entries needs 144 registers for M = 6 and a run-time pivot index: the arrays went to local memory -- a quarter
    // of the stall samples of the body-efforts kernels.)  Same operations in the same order per entry as the sequential form.
    const int lane_ = threadIdx.x & 31;
    double col[M];
#pragma unroll
    for (int r = 0; r < M; ++r) {
      double v = 0.0;
#pragma unroll
      for (int k = 0; k < M; ++k) v = (lane_ == k) ? S[r * M + k] : v;
      col[r] = (lane_ == M + r) ? 1.0 : v;
    }
#pragma unroll
    for (int c = 0; c < M; ++c) {
      double cc[M];
#pragma unroll
      for (int r = 0; r < M; ++r) cc[r] = __shfl_sync(kFull, col[r], c);
      int p = c;
      double best = fabs(cc[c]);
#pragma unroll
      for (int r = c + 1; r < M; ++r) {
        const double v = fabs(cc[r]);
        if (v > best) { best = v; p = r; }
      }
#pragma unroll
      for (int r = c + 1; r < M; ++r) {
        const bool sw = (r == p);
        const double t = col[r], u = cc[r];
        col[r] = sw ? col[c] : col[r]; col[c] = sw ? t : col[c];
        cc[r] = sw ? cc[c] : cc[r]; cc[c] = sw ? u : cc[c];
      }
      const double ip = 1.0 / cc[c];
      col[c] *= ip;
#pragma unroll
      for (int r = 0; r < M; ++r) {
        if (r != c) {
          const double f = cc[r];
          col[r] -= f * col[c];
        }
      }
    }
    // lanes M .. 2M-1 hold the columns of the inverse: hand every lane the whole matrix
#pragma unroll
    for (int r = 0; r < M; ++r)
#pragma unroll
      for (int k = 0; k < M; ++k) Si[r * M + k] = __shfl_sync(kFull, col[r], M + k);
  } else {
    double c00 = S[4] * S[8] - S[5] * S[7], c01 = S[5] * S[6] - S[3] * S[8], c02 = S[3] * S[7] - S[4] * S[6];
    double id = 1.0 / (S[0] * c00 + S[1] * c01 + S[2] * c02);
    Si[0] = c00 * id; Si[1] = (S[2] * S[7] - S[1] * S[8]) * id; Si[2] = (S[1] * S[5] - S[2] * S[4]) * id;
    Si[3] = c01 * id; Si[4] = (S[0] * S[8] - S[2] * S[6]) * id; Si[5] = (S[2] * S[3] - S[0] * S[5]) * id;
    Si[6] = c02 * id; Si[7] = (S[1] * S[6] - S[0] * S[7]) * id; Si[8] = (S[0] * S[4] - S[1] * S[3]) * id;
  }
}

// ukfom update(z, h, R, gate) for a vector measurement (reference call sites src/PoseUKF.cpp:479-610).
// gate95: the reference's d2p95 (src/PoseUKF.cpp:275-286), otherwise accept_any.
// gSig != nullptr (STREAMED form, layout WLS): the covariance lives in global memory at gSig (packed); the shared-memory S
// holds its six leading columns only, filled -- downdated -- by the streaming pass before apply_delta runs.
template <int KIND, int KMAX, class W = WL<KMAX>>
__device__ __forceinline__ bool update(double *w, const uint16_t *lut, const double *z, const double *Rm, const double *aux,
                                       bool gate95, const Params &P, int lane, unsigned &status, const double *model = nullptr,
                                       double *gSig = nullptr) {
  constexpr bool STREAM = std::is_same<W, WLS>::value;
  using MM = Meas<KIND>;
  constexpr int K = MM::K, M = MM::M;
  static_assert(K <= KMAX, "kernel instantiated with too small a panel");
  constexpr int NPTS = 2 * K + 1, ROUNDS = (NPTS + 31) / 32;
  double *mu = w + OFF_MU, *S = w + OFF_S, *Lw = w + W::L, *invd = w + W::INVD, *G = w + W::G, *T = w + W::T,
         *KS = w + W::KS, *dl = w + W::DL;
  // general path with room for N rows (dense instantiation): carry the factorisation over the whole panel, so that the
  // cross-covariance is a product with the panel factor (no triangular solve, no G Sigma)
  constexpr bool PANELF = (K > 32) && (KMAX >= N);
  constexpr int R = PANELF ? N : K;
  static_assert(!STREAM || PANELF, "the streamed form carries the panel factor");
  if (!chol_leading<K, R, STREAM>(STREAM ? gSig : S, Lw, invd, lane)) { status |= ST_NOT_SPD; return false; }
  // sigma points through the measurement model
  double Z[ROUNDS][M], wt[ROUNDS];
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const int p = lane + 32 * r;
    const bool act = p < NPTS;
    const bool on = act && p > 0;
    const int j = (p - 1) >> 1;
    const double sg = (p & 1) ? 1.0 : -1.0;
    auto d = [&](int t) { return (on && t >= j) ? sg * Lw[lidxr<R>(t, j)] : 0.0; };
    MM::eval(mu, d, aux, P, Z[r], model);
    wt[r] = (p == 0) ? double(1 + 2 * (N - K)) : (act ? 1.0 : 0.0);
  }
  // mean (ukfom mean_sigma_points on a vector manifold)
  double zb[M];
#pragma unroll
  for (int a = 0; a < M; ++a) zb[a] = bcast(Z[0][a], 0);
  {
    int it = 0;
    double nrm;
    do {
      nrm = 0;
#pragma unroll
      for (int a = 0; a < M; ++a) {
        double m = 0;
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) m = fma(wt[r], Z[r][a] - zb[a], m);
        m = wsum(m) / double(2 * N + 1);
        zb[a] += m;
        nrm += m * m;
      }
      nrm = sqrt(nrm);
    } while (nrm > P.mean_tol && ++it < P.mean_max_it);
    if (it >= P.mean_max_it) status |= ST_MEAN_NOT_CONVERGED;
  }
  // S = 1/2 sum dz dz^T + R
  double Szz[M * M];
#pragma unroll
  for (int a = 0; a < M; ++a)
#pragma unroll
    for (int b = 0; b <= a; ++b) {
      double s = 0;
#pragma unroll
      for (int r = 0; r < ROUNDS; ++r) s = fma(0.5 * wt[r] * (Z[r][a] - zb[a]), Z[r][b] - zb[b], s);
      s = wsum(s);
      Szz[a * M + b] = s + Rm[a * M + b];
      Szz[b * M + a] = s + Rm[b * M + a];
    }
  // B_j = 1/2 (Z_{2j+1} - Z_{2j+2}) -> G, then G <- B L^-1, T = G Sigma[0:K,:] = Pxz^T
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const int p = lane + 32 * r;
#pragma unroll
    for (int a = 0; a < M; ++a) {
      double src = Z[r][a];
      if (r + 1 < ROUNDS) { if (lane == 0) src = Z[r + 1 < ROUNDS ? r + 1 : r][a]; }
      double nxt = __shfl_sync(kFull, src, (lane + 1) & 31);
      if ((p & 1) && p < NPTS) G[a * W::KE + ((p - 1) >> 1)] = 0.5 * (Z[r][a] - nxt);
    }
  }
  __syncwarp();
  if constexpr (PANELF) {
    panel_times_b<K, M>(G, W::KE, Lw, T, lane);
  } else {
    solve_right<K, M>(G, W::KE, Lw, invd, lane);
    g_times_sigma<K, M>(G, W::KE, S, T, lane);
  }
  // gain, innovation, gate
  double Si[M * M];
  sym_inverse<M>(Szz, Si);
  double nu[M];
#pragma unroll
  for (int a = 0; a < M; ++a) nu[a] = z[a] - zb[a];
  double d2 = 0;
#pragma unroll
  for (int a = 0; a < M; ++a)
#pragma unroll
    for (int b = 0; b < M; ++b) d2 += nu[a] * Si[a * M + b] * nu[b];
  if (gate95 && d2 > 5.991) { status |= ST_GATE_REJECTED; return false; }
  for (int x = lane; x < N; x += 32) {
    double pxz[M], kx[M];
#pragma unroll
    for (int a = 0; a < M; ++a) pxz[a] = T[a * 56 + x];
    double dx = 0;
#pragma unroll
    for (int a = 0; a < M; ++a) {
      double s = 0;
#pragma unroll
      for (int c = 0; c < M; ++c) s += pxz[c] * Si[c * M + a];
      kx[a] = s;
      dx += s * nu[a];
    }
#pragma unroll
    for (int a = 0; a < M; ++a) {
      double s = 0;
#pragma unroll
      for (int c = 0; c < M; ++c) s += kx[c] * Szz[c * M + a];
      KS[a * 56 + x] = s;
      T[a * 56 + x] = kx[a];
    }
    dl[x] = dx;
  }
  __syncwarp();
  // Sigma -= (K S) K^T
  // STREAMED form: columns [b0, b1) of the packed covariance are read from HBM; rows x0 = lane, x1 = lane + 32 keep their (K S)
  // factors in registers and column b contributes K_b by M broadcast loads.  Columns 0..5 go to the shared-memory panel (all
  // that apply_delta works on), the others straight back to HBM -- but only once apply_delta has succeeded, so that a refused
  // update leaves the filter bit-identical in memory.
  auto stream_downdate = [&](int b0, int b1) {
    double ksx[2][M];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int x = lane + 32 * h;
#pragma unroll
      for (int c = 0; c < M; ++c) ksx[h][c] = x < N ? KS[c * 56 + x] : 0.0;
    }
#pragma unroll 8
    for (int b = b0; b < b1; ++b) {
      double kb[M];
#pragma unroll
      for (int c = 0; c < M; ++c) kb[c] = T[c * 56 + b];
      const int cb = cidx(b, b) - b;   // cidx(x, b) = cb + x
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int x = lane + 32 * h;
        if (x >= b && x < N) {
          double sv = gSig[cb + x];
#pragma unroll
          for (int c = 0; c < M; ++c) sv = fma(-ksx[h][c], kb[c], sv);
          if (b < 6) S[cb + x] = sv; else gSig[cb + x] = sv;
        }
      }
    }
    __syncwarp();
  };
  if constexpr (STREAM) {
    stream_downdate(0, 6);
  } else {
    for (int e = lane; e < NP; e += 32) {
      unsigned ij = lut[e];
      int i = ij & 0xff, jj = ij >> 8;
      double s = S[e];
#pragma unroll
      for (int c = 0; c < M; ++c) s = fma(-KS[c * 56 + i], T[c * 56 + jj], s);
      S[e] = s;
    }
    __syncwarp();
  }
  // apply_delta reuses the first three rows of T (its 3 x N cross-covariance); those rows of K are kept in registers for the
  // pass that follows (rollback, or the streamed rest of the downdate).  K S and the other rows of K stay where they are.
  constexpr int MK = M < 3 ? M : 3;
  double kr[2][MK];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    int x = lane + 32 * h;
#pragma unroll
    for (int a = 0; a < MK; ++a) kr[h][a] = x < N ? T[a * 56 + x] : 0.0;
  }
  __syncwarp();
  if (!apply_delta<KMAX, W>(w, P, lane, status)) {
    // downdated Sigma lost definiteness in its leading block: roll the downdate back, leave the filter as it was
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int x = lane + 32 * h;
      if (x < N) {
#pragma unroll
        for (int a = 0; a < MK; ++a) T[a * 56 + x] = kr[h][a];
      }
    }
    __syncwarp();
    if constexpr (!STREAM) {   // (streamed form: nothing has been written to HBM yet, the panel copy is simply dropped)
      for (int e = lane; e < NP; e += 32) {
        unsigned ij = lut[e];
        int i = ij & 0xff, jj = ij >> 8;
        double s = S[e];
#pragma unroll
        for (int c = 0; c < M; ++c) s = fma(KS[c * 56 + i], T[c * 56 + jj], s);
        S[e] = s;
      }
      __syncwarp();
    }
    return false;
  }
  if constexpr (STREAM) {   // accepted: the gain tables back from the registers (apply_delta reused T), then the rest of Sigma
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int x = lane + 32 * h;
      if (x < N) {
#pragma unroll
        for (int a = 0; a < MK; ++a) T[a * 56 + x] = kr[h][a];
      }
    }
    __syncwarp();
    stream_downdate(6, N);
  }
  return true;
}

}  // namespace pose
}  // namespace uwvkf
