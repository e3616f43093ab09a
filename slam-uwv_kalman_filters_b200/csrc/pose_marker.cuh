// PoseUKF visual-marker update on the marker-augmented state (reference src/PoseUKF.cpp:221-243, 613-654).
//
// The reference appends the marker pose (position 3 + SO3 3) BEHIND the 53 filter coordinates, so every column of the
// Cholesky factor moves the marker orientation and the measurement model (pose + marker pose -> bearing on S2) sees 119
// distinct sigma points: there is no prefix structure to exploit and the ukfom recipe is executed densely, one warp per
// filter with the 59 x 59 augmented covariance and its factor in shared memory:
//   per feature:  L = chol(Sigma_a);  Z_i = h(mu (+) +-L.col(j)) in S2;  zbar = iterated S2 mean;  S = 1/2 sum dz dz^T + R;
//                 Pxz = L B^T (B_j = 1/2 (dz_j+ - dz_j-));  K = Pxz S^-1;  Sigma_a -= (K S) K^T;  apply_delta(K (z (-) zbar))
//   apply_delta:  L = chol(Sigma_a); the two SO3 components are re-drawn through all 119 points (joint iterated mean), their
//                 rows of Sigma_a become A L^T (A_j = 1/2 (a_j+ - a_j-)) and 1/2 sum a a^T; vector rows are unchanged.
// Lanes run over sigma points (4 rounds of 32) or over covariance rows (2 rounds); reductions use the DMMA warp sum.
#pragma once
#include "pose_fused.cuh"
#include "small_engine.cuh"

namespace uwvkf {
namespace pose {

constexpr int NA = N + 6;                  // augmented tangent dimension (59)
constexpr int DA = DIM + 7;                // augmented stored dimension (61)
constexpr int NPA = NA * (NA + 1) / 2;     // packed lower triangle (1770)
constexpr int TA_MP = N, TA_MQ = N + 3;    // tangent offsets of marker position / orientation
constexpr int PA_MP = DIM, PA_MQ = DIM + 3;
constexpr int MAXFEAT = 8;

__host__ __device__ __forceinline__ int aidx(int i, int j) { return j * NA - ((j * (j - 1)) >> 1) + (i - j); }   // i >= j

struct MarkerArgs {
  int nfeat;
  double pxcov[MAXFEAT][4];      // pixel covariances
  double feat[MAXFEAT][3];       // feature positions in the marker frame
  double marker_pose[7];         // t(3), q(x,y,z,w): marker in the navigation frame
  double cov_marker[36];
  double camera[4];              // fx, fy, cx, cy
  double cam_in_imu[7];          // t(3), q(x,y,z,w)
};

struct MarkerLayout {              // per-warp shared memory (doubles)
  static constexpr int O_MU = 0;                 // 62
  static constexpr int O_A = 62;                 // NPA
  static constexpr int O_L = O_A + NPA;          // NPA
  static constexpr int O_PB = O_L + NPA;         // 6 x 60: B (2 rows) / A_q, A_m (3 + 3 rows)
  static constexpr int O_K = O_PB + 6 * 60;      // 2 x 60
  static constexpr int O_KS = O_K + 2 * 60;      // 2 x 60
  static constexpr int O_DL = O_KS + 2 * 60;     // 60
  static constexpr int TOTAL = O_DL + 60;
};

// Cholesky of the packed NA x NA matrix A into L (both in shared memory): blocked left-looking (chol_blocked, pose_engine.cuh)
__device__ __forceinline__ bool chol_aug(const double *A, double *L, int lane) {
  return chol_blocked<NA, NA, NA, true>(A, L, nullptr, lane);
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) pose_marker_kernel(double *state, uint32_t *status, const double *px, const uint8_t *valid,
                                                                    long long B, const __grid_constant__ MarkerArgs M,
                                                                    const __grid_constant__ Params P) {
  using W = MarkerLayout;
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double *w = smem + warp * W::TOTAL;
  double *mu = w + W::O_MU, *A = w + W::O_A, *L = w + W::O_L, *PB = w + W::O_PB, *KG = w + W::O_K, *KSm = w + W::O_KS, *DL = w + W::O_DL;
  const double fx = M.camera[0], fy = M.camera[1], cx = M.camera[2], cy = M.camera[3];
  const Quat camq{M.cam_in_imu[3], M.cam_in_imu[4], M.cam_in_imu[5], M.cam_in_imu[6]};
  const Vec3 camt{M.cam_in_imu[0], M.cam_in_imu[1], M.cam_in_imu[2]};
  constexpr double kInvPts = 1.0 / double(2 * NA + 1);
  const double tol2 = P.mean_tol * P.mean_tol;

  for (long long f = (long long)blockIdx.x * WARPS + warp; f < B; f += (long long)gridDim.x * WARPS) {
    if (valid && !valid[f]) continue;
    double *blk = state + f * BLK;
    const double *pxf = px + f * M.nfeat * 2;
    unsigned st = 0;
    {   // checkMeasurment() on every corner before anything is touched (reference src/PoseUKF.cpp:635)
      bool bad = false;
      for (int k = lane; k < M.nfeat * 2; k += 32) bad |= isnan(pxf[k]);
      for (int k = lane; k < M.nfeat * 4; k += 32) bad |= isnan(M.pxcov[k >> 2][k & 3]);
      if (__any_sync(kFull, bad)) {
        if (lane == 0) status[f] |= ST_NAN_MEAS;
        continue;
      }
    }
    // ---- augmented state and covariance (:619-626) ------------------------------------------------------------
    for (int k = lane; k < DIM; k += 32) mu[k] = blk[k];
    if (lane < 7) mu[DIM + lane] = M.marker_pose[lane];
    for (int j = 0; j < NA; ++j) {
      for (int i = j + lane; i < NA; i += 32) {
        double v = 0.0;
        if (i < N) v = blk[OFF_S + cidx(i, j)];
        else if (j >= N) v = M.cov_marker[(i - N) * 6 + (j - N)];
        A[aidx(i, j)] = v;
      }
    }
    __syncwarp();
    bool alive = true;

    for (int ft = 0; ft < M.nfeat && alive; ++ft) {
      // measurement on S2 and its covariance (:638-644)
      double zm[3] = {(pxf[2 * ft] - cx) / fx, (pxf[2 * ft + 1] - cy) / fy, 1.0};
      {
        const double n = sqrt(zm[0] * zm[0] + zm[1] * zm[1] + zm[2] * zm[2]);
        zm[0] /= n; zm[1] /= n; zm[2] /= n;
      }
      const double R00 = M.pxcov[ft][0] / (fx * fx), R01 = M.pxcov[ft][1] / (fx * fy), R10 = M.pxcov[ft][2] / (fx * fy),
                   R11 = M.pxcov[ft][3] / (fy * fy);
      const Vec3 feat{M.feat[ft][0], M.feat[ft][1], M.feat[ft][2]};
      if (!chol_aug(A, L, lane)) { st |= ST_NOT_SPD; alive = false; break; }
      // ---- sigma points through measurementVisualLandmark (:233-243) -----------------------------------------
      const Quat q0{mu[P_ORI], mu[P_ORI + 1], mu[P_ORI + 2], mu[P_ORI + 3]};
      const Quat mq0{mu[PA_MQ], mu[PA_MQ + 1], mu[PA_MQ + 2], mu[PA_MQ + 3]};
      double Z[4][3], wt[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int p = lane + 32 * r;
        const bool act = p < 2 * NA + 1, on = act && p > 0;
        const int j = on ? (p - 1) >> 1 : 0;
        const double sg = (p & 1) ? 1.0 : -1.0;
        auto d = [&](int t) { return (on && t >= j) ? sg * L[aidx(t, j)] : 0.0; };
        const Quat qi = so3_boxplus(q0, Vec3{d(3), d(4), d(5)}, 1.0);
        const Quat mqi = so3_boxplus(mq0, Vec3{d(TA_MQ), d(TA_MQ + 1), d(TA_MQ + 2)}, 1.0);
        Vec3 ct = qrot(qi, camt);                                  // imu_in_nav * cam_in_imu
        ct.x += mu[0] + d(0); ct.y += mu[1] + d(1); ct.z += mu[2] + d(2);
        const Quat cq = qmul(qi, camq);
        Vec3 fp = qrot(mqi, feat);                                 // marker_orientation * feature + marker_position
        fp.x += mu[PA_MP] + d(TA_MP) - ct.x; fp.y += mu[PA_MP + 1] + d(TA_MP + 1) - ct.y; fp.z += mu[PA_MP + 2] + d(TA_MP + 2) - ct.z;
        const Vec3 zc = qrot_inv(cq, fp);
        const double n = sqrt(zc.x * zc.x + zc.y * zc.y + zc.z * zc.z);
        Z[r][0] = zc.x / n; Z[r][1] = zc.y / n; Z[r][2] = zc.z / n;
        wt[r] = act ? 1.0 : 0.0;
      }
      // ---- iterated S2 mean, deviations, innovation covariance ------------------------------------------------
      double zb[3] = {bcast(Z[0][0], 0), bcast(Z[0][1], 0), bcast(Z[0][2], 0)};
      double dz[4][2];
      {
        int it = 0;
        double n2;
        do {
          double m0 = 0, m1 = 0;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            small::s2_boxminus(Z[r], zb, dz[r]);
            m0 = fma(wt[r], dz[r][0], m0); m1 = fma(wt[r], dz[r][1], m1);
          }
          const double mm[2] = {wsum(m0) * kInvPts, wsum(m1) * kInvPts};
          double nz[3];
          small::s2_boxplus(zb, mm, nz);
          zb[0] = nz[0]; zb[1] = nz[1]; zb[2] = nz[2];
          n2 = mm[0] * mm[0] + mm[1] * mm[1];
        } while (n2 > tol2 && ++it < P.mean_max_it);
        if (it >= P.mean_max_it) st |= ST_MEAN_NOT_CONVERGED;
      }
      double s00 = 0, s10 = 0, s11 = 0;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        small::s2_boxminus(Z[r], zb, dz[r]);
        dz[r][0] *= wt[r]; dz[r][1] *= wt[r];
        s00 = fma(0.5 * dz[r][0], dz[r][0], s00); s10 = fma(0.5 * dz[r][1], dz[r][0], s10); s11 = fma(0.5 * dz[r][1], dz[r][1], s11);
      }
      s00 = wsum(s00); s10 = wsum(s10); s11 = wsum(s11);
      const double Szz[4] = {s00 + R00, s10 + R01, s10 + R10, s11 + R11};
      // B_j = 1/2 (dz_{2j+1} - dz_{2j+2}): odd points pair with the next point (next lane, or lane 0 of the next round)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int p = lane + 32 * r;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          double src = dz[r][c];
          if (lane == 0 && r + 1 < 4) src = dz[r + 1 < 4 ? r + 1 : r][c];
          const double nxt = __shfl_sync(kFull, src, (lane + 1) & 31);
          if ((p & 1) && p < 2 * NA + 1) PB[c * 60 + ((p - 1) >> 1)] = 0.5 * (dz[r][c] - nxt);
        }
      }
      __syncwarp();
      // ---- gain, innovation, downdate ------------------------------------------------------------------------------
      double Si[4];
      {
        const double id = 1.0 / (Szz[0] * Szz[3] - Szz[1] * Szz[2]);
        Si[0] = Szz[3] * id; Si[1] = -Szz[1] * id; Si[2] = -Szz[2] * id; Si[3] = Szz[0] * id;
      }
      double nu[2];
      small::s2_boxminus(zm, zb, nu);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int x = lane + 32 * h;
        if (x < NA) {
          double p0 = 0, p1 = 0;
          int e = x;                                       // aidx(x, 0)
          for (int j = 0; j <= x; ++j) {
            const double l = L[e];
            p0 = fma(l, PB[j], p0); p1 = fma(l, PB[60 + j], p1);
            e += NA - j - 1;
          }
          const double k0 = p0 * Si[0] + p1 * Si[2], k1 = p0 * Si[1] + p1 * Si[3];
          KG[x] = k0; KG[60 + x] = k1;
          KSm[x] = k0 * Szz[0] + k1 * Szz[2]; KSm[60 + x] = k0 * Szz[1] + k1 * Szz[3];
          DL[x] = k0 * nu[0] + k1 * nu[1];
        }
      }
      __syncwarp();
      for (int b = 0; b < NA; ++b) {
        const double k0 = KG[b], k1 = KG[60 + b];
        for (int x = b + lane; x < NA; x += 32) A[aidx(x, b)] = fma(-KSm[60 + x], k1, fma(-KSm[x], k0, A[aidx(x, b)]));
      }
      __syncwarp();
      // ---- apply_delta ------------------------------------------------------------------------------------------
      if (!chol_aug(A, L, lane)) { st |= ST_NOT_SPD; alive = false; break; }
      const Vec3 dq{DL[3], DL[4], DL[5]}, dm{DL[TA_MQ], DL[TA_MQ + 1], DL[TA_MQ + 2]};
      Quat qp[4], mp[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int p = lane + 32 * r;
        const bool on = p > 0 && p < 2 * NA + 1;
        const int j = on ? (p - 1) >> 1 : 0;
        const double sg = (p & 1) ? 1.0 : -1.0;
        auto d = [&](int t) { return (on && t >= j) ? sg * L[aidx(t, j)] : 0.0; };
        qp[r] = so3_boxplus(q0, Vec3{dq.x + d(3), dq.y + d(4), dq.z + d(5)}, 1.0);
        mp[r] = so3_boxplus(mq0, Vec3{dm.x + d(TA_MQ), dm.y + d(TA_MQ + 1), dm.z + d(TA_MQ + 2)}, 1.0);
      }
      Quat rq{bcast(qp[0].x, 0), bcast(qp[0].y, 0), bcast(qp[0].z, 0), bcast(qp[0].w, 0)};
      Quat rm{bcast(mp[0].x, 0), bcast(mp[0].y, 0), bcast(mp[0].z, 0), bcast(mp[0].w, 0)};
      {
        int it = 0;
        double n2;
        do {
          double a[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const Vec3 u = so3_boxminus(qp[r], rq), v = so3_boxminus(mp[r], rm);
            a[0] = fma(wt[r], u.x, a[0]); a[1] = fma(wt[r], u.y, a[1]); a[2] = fma(wt[r], u.z, a[2]);
            a[3] = fma(wt[r], v.x, a[3]); a[4] = fma(wt[r], v.y, a[4]); a[5] = fma(wt[r], v.z, a[5]);
          }
          n2 = 0;
#pragma unroll
          for (int k = 0; k < 6; ++k) { a[k] = wsum(a[k]) * kInvPts; n2 = fma(a[k], a[k], n2); }
          rq = so3_boxplus(rq, Vec3{a[0], a[1], a[2]}, 1.0);
          rm = so3_boxplus(rm, Vec3{a[3], a[4], a[5]}, 1.0);
        } while (n2 > tol2 && ++it < P.mean_max_it);
        if (it >= P.mean_max_it) st |= ST_MEAN_NOT_CONVERGED;
      }
      // deviations of the two SO3 components, their covariance blocks and half-differences
      double dv[4][6];
      double cq[21];
#pragma unroll
      for (int k = 0; k < 21; ++k) cq[k] = 0.0;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const Vec3 u = so3_boxminus(qp[r], rq), v = so3_boxminus(mp[r], rm);
        dv[r][0] = wt[r] * u.x; dv[r][1] = wt[r] * u.y; dv[r][2] = wt[r] * u.z;
        dv[r][3] = wt[r] * v.x; dv[r][4] = wt[r] * v.y; dv[r][5] = wt[r] * v.z;
        int k = 0;
#pragma unroll
        for (int a = 0; a < 6; ++a)
#pragma unroll
          for (int b = 0; b <= a; ++b) { cq[k] = fma(0.5 * dv[r][a], dv[r][b], cq[k]); ++k; }
      }
#pragma unroll
      for (int k = 0; k < 21; ++k) cq[k] = wsum(cq[k]);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int p = lane + 32 * r;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          double src = dv[r][c];
          if (lane == 0 && r + 1 < 4) src = dv[r + 1 < 4 ? r + 1 : r][c];
          const double nxt = __shfl_sync(kFull, src, (lane + 1) & 31);
          if ((p & 1) && p < 2 * NA + 1) PB[c * 60 + ((p - 1) >> 1)] = 0.5 * (dv[r][c] - nxt);
        }
      }
      __syncwarp();
      // rows of the two SO3 components against the vector coordinates: A L^T
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int x = lane + 32 * h;
        const bool isvec = x < NA && !(x >= 3 && x < 6) && x < TA_MQ;
        if (isvec) {
          double acc[6] = {0, 0, 0, 0, 0, 0};
          int e = x;
          for (int j = 0; j <= x; e += NA - j - 1, ++j) {
            const double l = L[e];
#pragma unroll
            for (int c = 0; c < 6; ++c) acc[c] = fma(l, PB[c * 60 + j], acc[c]);
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            A[x < 3 ? aidx(3 + c, x) : aidx(x, 3 + c)] = acc[c];
            A[aidx(TA_MQ + c, x)] = acc[3 + c];
          }
        }
      }
      if (lane == 0) {
        int k = 0;
        for (int a = 0; a < 6; ++a)
          for (int b = 0; b <= a; ++b) {
            const int ia = a < 3 ? 3 + a : TA_MQ + a - 3, ib = b < 3 ? 3 + b : TA_MQ + b - 3;
            A[aidx(ia, ib)] = cq[k++];
          }
      }
      // mean: vector components move by delta, the SO3 components take their iterated means
      for (int x = lane; x < NA; x += 32) {
        if (x < 3) mu[x] += DL[x];
        else if (x >= 6 && x < TA_MQ) mu[x + 1] += DL[x];
      }
      if (lane == 0) {
        mu[P_ORI] = rq.x; mu[P_ORI + 1] = rq.y; mu[P_ORI + 2] = rq.z; mu[P_ORI + 3] = rq.w;
        mu[PA_MQ] = rm.x; mu[PA_MQ + 1] = rm.y; mu[PA_MQ + 2] = rm.z; mu[PA_MQ + 3] = rm.w;
      }
      __syncwarp();
    }
    // ---- keep the leading 53 x 53 block and the filter state (:653); a failed factorisation leaves the filter untouched
    if (alive) {
      for (int k = lane; k < DIM; k += 32) blk[k] = mu[k];
      for (int j = 0; j < N; ++j)
        for (int i = j + lane; i < N; i += 32) blk[OFF_S + cidx(i, j)] = A[aidx(i, j)];
    }
    if (st && lane == 0) status[f] |= st;
    __syncwarp();
  }
}

}  // namespace pose
}  // namespace uwvkf
