// Batched estimators, one env per thread, state in registers (kernels K2 and K3).
//   K3  9-state position/velocity/accel-bias Kalman filter   CPU twin: oracle/pv_filter.py
//       reference: isaacgymenvs/PVFilter.py:25-64 (prediction_step), :67-110 (correction_step), :113-142
//   K2  4-state attitude EKF, float64                          CPU twin: oracle/ahrs_ekf.py
//       reference: isaacgymenvs/ahrs_ekf.py:1072-1158, 1280-1337 (the `ang` branch)
// The reference runs both as Python loops over envs (tasks/ekf_lee_landed.py:378-391, 417-444: N iterations of ~15 numpy
// calls / ~40 tiny CUDA launches per step).  Here one launch handles all envs; HBM traffic is the filter state once in,
// once out (SoA planes [k][N], coalesced across envs).
#pragma once
#include <math.h>

namespace ozl {

// ------------------------------------------------------------------------------------------------ K3: PV filter
struct PV {
    float x[9];
    float P[9][9];
};

// quaternion_to_matrix of PVFilter.py:113-142 (normalises first), wxyz in
__device__ __forceinline__ void pv_quat_to_R(float r, float i, float j, float k, float R[3][3]) {
    const float n = sqrtf(((r * r + i * i) + j * j) + k * k);
    r /= n; i /= n; j /= n; k /= n;
    const float two_s = 2.0f / (((r * r + i * i) + j * j) + k * k);
    R[0][0] = 1.f - two_s * (j * j + k * k); R[0][1] = two_s * (i * j - k * r); R[0][2] = two_s * (i * k + j * r);
    R[1][0] = two_s * (i * j + k * r); R[1][1] = 1.f - two_s * (i * i + k * k); R[1][2] = two_s * (j * k - i * r);
    R[2][0] = two_s * (i * k - j * r); R[2][1] = two_s * (j * k + i * r); R[2][2] = 1.f - two_s * (i * i + j * j);
}

// prediction_step: x <- F x + G (a - b_a) ; P <- F P F^T + G diag(acc_var) G^T, with the reference's F and G:
//   F = [[I, R dt, R dt^2/2], [0, R, R dt], [0, 0, I]],  G = [R dt^2/2; R dt; 0],  R = quaternion_to_matrix(q)^T
__device__ __forceinline__ void pv_predict(PV& s, const float acc[3], const float q_wxyz[4], float dt, float dt2,
                                           const float acc_var[3]) {
    float Rq[3][3], R[3][3], A[3][3], B[3][3];
    pv_quat_to_R(q_wxyz[0], q_wxyz[1], q_wxyz[2], q_wxyz[3], Rq);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            R[i][j] = Rq[j][i];                    // .T   (PVFilter.py:33-35)
            A[i][j] = R[i][j] * dt;                // F[0:3,3:6], F[3:6,6:9]
            B[i][j] = (R[i][j] * dt2) * 0.5f;      // F[0:3,6:9]
        }
    // ---- state
    float u[3], np_[3], nv[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) u[i] = acc[i] - s.x[6 + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float fp = s.x[i], fv = 0.f, gp = 0.f, gv = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            fp += A[i][j] * s.x[3 + j];
            fv += R[i][j] * s.x[3 + j];
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            fp += B[i][j] * s.x[6 + j];
            fv += A[i][j] * s.x[6 + j];
            gp += B[i][j] * u[j];
            gv += A[i][j] * u[j];
        }
        np_[i] = fp + gp;
        nv[i] = fv + gv;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) { s.x[i] = np_[i]; s.x[3 + i] = nv[i]; }
    // ---- covariance: M = F P (rows), in place
#pragma unroll
    for (int c = 0; c < 9; ++c) {
        float pv[3], pb[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) { pv[j] = s.P[3 + j][c]; pb[j] = s.P[6 + j][c]; }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float mp = s.P[i][c], mv = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) { mp += A[i][j] * pv[j]; mv += R[i][j] * pv[j]; }
#pragma unroll
            for (int j = 0; j < 3; ++j) { mp += B[i][j] * pb[j]; mv += A[i][j] * pb[j]; }
            s.P[i][c] = mp;
            s.P[3 + i][c] = mv;
        }
    }
    // N = M F^T (columns), in place
#pragma unroll
    for (int r = 0; r < 9; ++r) {
        float mv[3], mb[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) { mv[j] = s.P[r][3 + j]; mb[j] = s.P[r][6 + j]; }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float np2 = s.P[r][i], nv2 = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) { np2 += mv[j] * A[i][j]; nv2 += mv[j] * R[i][j]; }
#pragma unroll
            for (int j = 0; j < 3; ++j) { np2 += mb[j] * B[i][j]; nv2 += mb[j] * A[i][j]; }
            s.P[r][i] = np2;
            s.P[r][3 + i] = nv2;
        }
    }
    // + G Q G^T   (only the p/v blocks are non-zero)
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float bb = 0.f, ba = 0.f, ab = 0.f, aa = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                bb += (B[i][k] * acc_var[k]) * B[j][k];
                ba += (B[i][k] * acc_var[k]) * A[j][k];
                ab += (A[i][k] * acc_var[k]) * B[j][k];
                aa += (A[i][k] * acc_var[k]) * A[j][k];
            }
            s.P[i][j] += bb; s.P[i][3 + j] += ba; s.P[3 + i][j] += ab; s.P[3 + i][3 + j] += aa;
        }
}

// correction_step on block H = [lo, lo+3):  K = P[:,H] inv(P[H,H] + diag(rvar)) ; x += K (z - x[H]) ; P = (I - K H) P
template <int LO>
__device__ __forceinline__ void pv_correct(PV& s, const float z[3], const float rvar[3]) {
    float S[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) S[i][j] = s.P[LO + i][LO + j] + (i == j ? rvar[i] : 0.f);
    // 3x3 inverse by the adjugate
    const float c00 = S[1][1] * S[2][2] - S[1][2] * S[2][1];
    const float c01 = S[1][2] * S[2][0] - S[1][0] * S[2][2];
    const float c02 = S[1][0] * S[2][1] - S[1][1] * S[2][0];
    const float det = (S[0][0] * c00 + S[0][1] * c01) + S[0][2] * c02;
    const float id = 1.0f / det;
    float Si[3][3];
    Si[0][0] = c00 * id; Si[1][0] = c01 * id; Si[2][0] = c02 * id;
    Si[0][1] = (S[0][2] * S[2][1] - S[0][1] * S[2][2]) * id;
    Si[1][1] = (S[0][0] * S[2][2] - S[0][2] * S[2][0]) * id;
    Si[2][1] = (S[0][1] * S[2][0] - S[0][0] * S[2][1]) * id;
    Si[0][2] = (S[0][1] * S[1][2] - S[0][2] * S[1][1]) * id;
    Si[1][2] = (S[0][2] * S[1][0] - S[0][0] * S[1][2]) * id;
    Si[2][2] = (S[0][0] * S[1][1] - S[0][1] * S[1][0]) * id;
    float K[9][3];
#pragma unroll
    for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) K[i][j] = (s.P[i][LO] * Si[0][j] + s.P[i][LO + 1] * Si[1][j]) + s.P[i][LO + 2] * Si[2][j];
    const float inn[3] = {z[0] - s.x[LO], z[1] - s.x[LO + 1], z[2] - s.x[LO + 2]};
#pragma unroll
    for (int i = 0; i < 9; ++i) s.x[i] = s.x[i] + ((K[i][0] * inn[0] + K[i][1] * inn[1]) + K[i][2] * inn[2]);
    // P <- IKH @ P with IKH = I, IKH[:,H] -= K  (PVFilter.py:88-89 / 108-109)
    float PH[3][9];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int j = 0; j < 9; ++j) PH[k][j] = s.P[LO + k][j];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        float w[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) w[k] = ((i == LO + k) ? 1.0f : 0.0f) - K[i][k];
        const bool inH = (i >= LO) && (i < LO + 3);
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const float acc3 = (w[0] * PH[0][j] + w[1] * PH[1][j]) + w[2] * PH[2][j];
            s.P[i][j] = inH ? acc3 : (s.P[i][j] + acc3);
        }
    }
}

// ------------------------------------------------------------------------------------------------ K3, covariance in shared memory
// Same arithmetic as pv_predict / pv_correct (same operations in the same order), but the 9x9 covariance of the thread's env
// lives in SHARED memory -- element (r,c) of thread t at P[(r*9 + c) * STRIDE + t], conflict-free -- and the loops over
// columns / rows stay ROLLED.  The register variants above unroll into ~2000 straight-line instructions holding 81 + ~50 live
// floats per thread (168 registers, 3-4 CTAs/SM, instruction-cache misses: "no_instruction" was 19 % of the stall cycles of
// the fused kernel); this form needs ~60 registers for the filter and ~250 instructions of loop body.
template <int STRIDE>
struct PVShared {
    float x[9];
    float* P;                              // this thread's column of the block's [81][STRIDE] covariance tile
    __device__ __forceinline__ float& at(int r, int c) { return P[(r * 9 + c) * STRIDE]; }
};

template <int STRIDE>
__device__ __forceinline__ void pv_predict(PVShared<STRIDE>& s, const float acc[3], const float q_wxyz[4], float dt, float dt2,
                                           const float acc_var[3]) {
    float Rq[3][3], R[3][3], A[3][3], B[3][3];
    pv_quat_to_R(q_wxyz[0], q_wxyz[1], q_wxyz[2], q_wxyz[3], Rq);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            R[i][j] = Rq[j][i];
            A[i][j] = R[i][j] * dt;
            B[i][j] = (R[i][j] * dt2) * 0.5f;
        }
    float u[3], np_[3], nv[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) u[i] = acc[i] - s.x[6 + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float fp = s.x[i], fv = 0.f, gp = 0.f, gv = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            fp += A[i][j] * s.x[3 + j];
            fv += R[i][j] * s.x[3 + j];
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            fp += B[i][j] * s.x[6 + j];
            fv += A[i][j] * s.x[6 + j];
            gp += B[i][j] * u[j];
            gv += A[i][j] * u[j];
        }
        np_[i] = fp + gp;
        nv[i] = fv + gv;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) { s.x[i] = np_[i]; s.x[3 + i] = nv[i]; }
    // M = F P, one column per iteration
#pragma unroll 1
    for (int c = 0; c < 9; ++c) {
        float pp[3], pv[3], pb[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) { pp[j] = s.at(j, c); pv[j] = s.at(3 + j, c); pb[j] = s.at(6 + j, c); }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float mp = pp[i], mv = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) { mp += A[i][j] * pv[j]; mv += R[i][j] * pv[j]; }
#pragma unroll
            for (int j = 0; j < 3; ++j) { mp += B[i][j] * pb[j]; mv += A[i][j] * pb[j]; }
            s.at(i, c) = mp;
            s.at(3 + i, c) = mv;
        }
    }
    // N = M F^T, one row per iteration
#pragma unroll 1
    for (int r = 0; r < 9; ++r) {
        float mp3[3], mv[3], mb[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) { mp3[j] = s.at(r, j); mv[j] = s.at(r, 3 + j); mb[j] = s.at(r, 6 + j); }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float np2 = mp3[i], nv2 = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) { np2 += mv[j] * A[i][j]; nv2 += mv[j] * R[i][j]; }
#pragma unroll
            for (int j = 0; j < 3; ++j) { np2 += mb[j] * B[i][j]; nv2 += mb[j] * A[i][j]; }
            s.at(r, i) = np2;
            s.at(r, 3 + i) = nv2;
        }
    }
    // + G Q G^T
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float bb = 0.f, ba = 0.f, ab = 0.f, aa = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                bb += (B[i][k] * acc_var[k]) * B[j][k];
                ba += (B[i][k] * acc_var[k]) * A[j][k];
                ab += (A[i][k] * acc_var[k]) * B[j][k];
                aa += (A[i][k] * acc_var[k]) * A[j][k];
            }
            s.at(i, j) += bb; s.at(i, 3 + j) += ba; s.at(3 + i, j) += ab; s.at(3 + i, 3 + j) += aa;
        }
}

template <int LO, int STRIDE>
__device__ __forceinline__ void pv_correct(PVShared<STRIDE>& s, const float z[3], const float rvar[3]) {
    // rows H of the OLD covariance (also S = P[H,H] + diag(rvar))
    float PH[3][9];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int j = 0; j < 9; ++j) PH[k][j] = s.at(LO + k, j);
    float S[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) S[i][j] = PH[i][LO + j] + (i == j ? rvar[i] : 0.f);
    const float c00 = S[1][1] * S[2][2] - S[1][2] * S[2][1];
    const float c01 = S[1][2] * S[2][0] - S[1][0] * S[2][2];
    const float c02 = S[1][0] * S[2][1] - S[1][1] * S[2][0];
    const float det = (S[0][0] * c00 + S[0][1] * c01) + S[0][2] * c02;
    const float id = 1.0f / det;
    float Si[3][3];
    Si[0][0] = c00 * id; Si[1][0] = c01 * id; Si[2][0] = c02 * id;
    Si[0][1] = (S[0][2] * S[2][1] - S[0][1] * S[2][2]) * id;
    Si[1][1] = (S[0][0] * S[2][2] - S[0][2] * S[2][0]) * id;
    Si[2][1] = (S[0][1] * S[2][0] - S[0][0] * S[2][1]) * id;
    Si[0][2] = (S[0][1] * S[1][2] - S[0][2] * S[1][1]) * id;
    Si[1][2] = (S[0][2] * S[1][0] - S[0][0] * S[1][2]) * id;
    Si[2][2] = (S[0][0] * S[1][1] - S[0][1] * S[1][0]) * id;
    const float inn[3] = {z[0] - s.x[LO], z[1] - s.x[LO + 1], z[2] - s.x[LO + 2]};
    // one row of K, x and P per iteration: row i of K needs only row i of the old P, read before the row is overwritten
    float xn[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) xn[i] = s.x[i];
#pragma unroll 1
    for (int i = 0; i < 9; ++i) {
        float row[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) row[j] = s.at(i, j);
        float K[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) K[j] = (row[LO] * Si[0][j] + row[LO + 1] * Si[1][j]) + row[LO + 2] * Si[2][j];
        const float dx = (K[0] * inn[0] + K[1] * inn[1]) + K[2] * inn[2];
#pragma unroll
        for (int ii = 0; ii < 9; ++ii) xn[ii] = (ii == i) ? xn[ii] + dx : xn[ii];      // select chain: no dynamic register index
        float w[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) w[k] = ((i == LO + k) ? 1.0f : 0.0f) - K[k];
        const bool inH = (i >= LO) && (i < LO + 3);
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const float acc3 = (w[0] * PH[0][j] + w[1] * PH[1][j]) + w[2] * PH[2][j];
            s.at(i, j) = inH ? acc3 : (row[j] + acc3);
        }
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) s.x[i] = xn[i];
}

// ------------------------------------------------------------------------------------------------ K2: attitude EKF
struct EKF4 {
    double q[4];        // wxyz
    double P[4][4];
};

// EKF.update(q/|q|, gyr, ang) -- ahrs_ekf.py:1301-1337, `ang` branch.  Dt = 1/frequency, g_noise = 0.3^2.
__device__ __forceinline__ void ekf_update(EKF4& s, const double g[3], const double ang[4], double Dt, double g_noise,
                                           double s_eps) {
    // caller-side normalisation (tasks/ekf_lee_landed.py:386: q=self.Q_state[idx]/np.linalg.norm(...))
    double nq = sqrt(((s.q[0] * s.q[0] + s.q[1] * s.q[1]) + s.q[2] * s.q[2]) + s.q[3] * s.q[3]);
    double q[4] = {s.q[0] / nq, s.q[1] / nq, s.q[2] / nq, s.q[3] / nq};
    const double hd = 0.5 * Dt;
    // Omega(x) rows: [0,-x0,-x1,-x2],[x0,0,x2,-x1],[x1,-x2,0,x0],[x2,x1,-x0,0]          (:1100-1106)
    double Om[4][4] = {{0.0, -g[0], -g[1], -g[2]}, {g[0], 0.0, g[2], -g[1]}, {g[1], -g[2], 0.0, g[0]}, {g[2], g[1], -g[0], 0.0}};
    double qt[4], F[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double a = (i == j ? 1.0 : 0.0) + hd * Om[i][j];       // (I + 0.5 Dt Omega) q        (:1132-1133)
            acc += a * q[j];
        }
        qt[i] = acc;
    }
    const double x[3] = {hd * g[0], hd * g[1], hd * g[2]};                // F = I + Omega(0.5 Dt g)      (:1157-1158)
    double Ox[4][4] = {{0.0, -x[0], -x[1], -x[2]}, {x[0], 0.0, x[2], -x[1]}, {x[1], -x[2], 0.0, x[0]}, {x[2], x[1], -x[0], 0.0}};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) F[i][j] = (i == j ? 1.0 : 0.0) + Ox[i][j];
    // W = 0.5 Dt [ -q_v ; q_w I + skew(q_v) ]                                                         (:1320)
    double W[4][3] = {{-q[1], -q[2], -q[3]}, {q[0], -q[3], q[2]}, {q[3], q[0], -q[1]}, {-q[2], q[1], q[0]}};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) W[i][j] = hd * W[i][j];
    // P_t = F P F^T + 0.5 Dt g_noise W W^T                                                            (:1321-1322)
    double FP[4][4], Pt[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) a += F[i][k] * s.P[k][j];
            FP[i][j] = a;
        }
    const double qs = hd * g_noise;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double a = 0.0, ww = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) a += FP[i][k] * F[j][k];
#pragma unroll
            for (int k = 0; k < 3; ++k) ww += W[i][k] * W[j][k];
            Pt[i][j] = a + qs * ww;
        }
    // S = P_t + eps I ; K = P_t S^-1  (Gauss-Jordan on the SPD 4x4, no pivoting)                       (:1332-1333)
    double S[4][4], Si[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { S[i][j] = Pt[i][j] + (i == j ? s_eps : 0.0); Si[i][j] = (i == j ? 1.0 : 0.0); }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const double ip = 1.0 / S[c][c];
#pragma unroll
        for (int j = 0; j < 4; ++j) { S[c][j] *= ip; Si[c][j] *= ip; }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (r == c) continue;
            const double f = S[r][c];
#pragma unroll
            for (int j = 0; j < 4; ++j) { S[r][j] -= f * S[c][j]; Si[r][j] -= f * Si[c][j]; }
        }
    }
    double K[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) a += Pt[i][k] * Si[k][j];
            K[i][j] = a;
        }
    // P = (I - K) P_t ; q = normalize(q_t + K (ang - q_t))                                             (:1334-1336)
    double v[4], qn[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = ang[i] - qt[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) a += K[i][k] * v[k];
        qn[i] = qt[i] + a;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double b = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) b += ((i == k ? 1.0 : 0.0) - K[i][k]) * Pt[k][j];
            s.P[i][j] = b;
        }
    }
    const double nn = sqrt(((qn[0] * qn[0] + qn[1] * qn[1]) + qn[2] * qn[2]) + qn[3] * qn[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) s.q[i] = qn[i] / nn;
}

}  // namespace ozl
